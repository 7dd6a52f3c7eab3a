python -m pytest tests -m gpu -x -q -k "fp32 or golden or cpp_api or reorth or full_size" 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_c.log 2>gpurun_out/bench_r02_c.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r02_c.log") if l.startswith("{")][-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"spmv_ms",d["detail"]["spmv_ms_avg"],"parity",d["parity"]["ok"],"reorth",d["reorth_variant"]["value"])
print(json.dumps(d["basis_f32"]))
PY
tail -3 gpurun_out/bench_r02_c.err
