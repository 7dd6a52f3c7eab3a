bash tools/run_n.sh 2 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_n2_final.log 2>gpurun_out/bench_r02_n2_final.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r02_n2_final.log") if l.startswith("{")][-1])
print("N=2 value",d["value"],"e2e",d["e2e"]["value"],"rank",d["e2e_rank"]["value"],"parity",d["parity"]["ok"],"reorth",d["reorth_variant"]["value"])
print(json.dumps(d["detail"]["step_breakdown_us"]))
PY
tail -3 gpurun_out/bench_r02_n2_final.err
