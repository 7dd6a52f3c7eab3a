python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_r02_final.log 2>gpurun_out/bench_r02_final.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r02_final.log") if l.startswith("{")][-1])
print("C3 value",d["value"],"e2e",d["e2e"]["value"],"e2e_rank",d["e2e_rank"]["value"],"spmv_ms",d["detail"]["spmv_ms_avg"],"parity",d["parity"]["ok"],"reorth",d["reorth_variant"]["value"], d["reorth_variant"]["roofline"]["frac"], "cpu", d["cpu_baseline"])
print(json.dumps(d["basis_f32"]))
PY
tail -3 gpurun_out/bench_r02_final.err
python -c "
import __graft_entry__ as g; g.smoke()"
