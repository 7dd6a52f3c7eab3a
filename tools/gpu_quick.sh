python -m pytest tests -m gpu -x -q -k "spmv or quad or golden or summary" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f32-detail --no-reorth-detail > gpurun_out/bench_q.log 2>gpurun_out/bench_q.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_q.log") if l.startswith("{")][-1])
print("C3 value",d["value"],"e2e",d["e2e"]["value"],"e2e_rank",d["e2e_rank"]["value"],"spmv_ms",d["detail"]["spmv_ms_avg"],"parity",d["parity"]["ok"], d["parity"]["rel_2norm"])
PY
tail -3 gpurun_out/bench_q.err
