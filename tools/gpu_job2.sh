python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -6
python -m pytest tests -m gpu -x -q -k "fp32 or reorth or multout" 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_d.log 2>gpurun_out/bench_r02_d.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r02_d.log") if l.startswith("{")][-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"spmv_ms",d["detail"]["spmv_ms_avg"],"parity",d["parity"]["ok"],"reorth",d["reorth_variant"]["value"], "graph_build_s", d["detail"]["graph_build_s"])
print(json.dumps(d["basis_f32"]))
PY
tail -3 gpurun_out/bench_r02_d.err
bash tools/run_n.sh 2 bench.py --gpus 2 --steps 5 --warmup 3 --no-reorth-detail > gpurun_out/bench_r02_d2.log 2>gpurun_out/bench_r02_d2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r02_d2.log") if l.startswith("{")][-1])
print("N=2 value",d["value"],"e2e",d["e2e"]["value"],"parity",d["parity"]["ok"], "graph_build_s", d["detail"]["graph_build_s"])
PY
tail -3 gpurun_out/bench_r02_d2.err
