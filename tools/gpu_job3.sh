python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_e.log 2>gpurun_out/bench_r02_e.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r02_e.log") if l.startswith("{")][-1])
print("C3 value",d["value"],"e2e",d["e2e"]["value"],"spmv_ms",d["detail"]["spmv_ms_avg"],"parity",d["parity"]["ok"],"reorth",d["reorth_variant"]["value"], d["reorth_variant"]["roofline"]["frac"])
print(json.dumps(d["basis_f32"]))
PY
tail -3 gpurun_out/bench_r02_e.err
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-f32-detail --save-summary gpurun_out/own_rmat_s27_k50_summary.npz > gpurun_out/bench_r02_c4_n1.log 2>gpurun_out/bench_r02_c4_n1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r02_c4_n1.log") if l.startswith("{")][-1])
print("C4 N=1 value",d["value"],"e2e",d["e2e"]["value"],"spmv_ms",d["detail"]["spmv_ms_avg"],"finite",d["result_finite"],"parity",d["parity"], "graph_build_s", d["detail"]["graph_build_s"])
PY
tail -3 gpurun_out/bench_r02_c4_n1.err; ls -la gpurun_out/*.npz
