python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4 | tee gpurun_out/r02_multi_tests_8gpu.log
bash tools/gpu_c4.sh 8
for wl in c3 c5; do
bash tools/run_n.sh 8 bench.py --gpus 8 --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_r02_${wl}_n8.log 2>gpurun_out/bench_r02_${wl}_n8.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_r02_${wl}_n8.log") if l.startswith("{")][-1])
    p=d["parity"]
    print("$wl N=8 value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"rank",round(d["e2e_rank"]["value"],1),"spmv_ms",round(d["detail"]["spmv_ms_avg"],3),"upd_ms",round(d["detail"]["update_scale_ms_per_iter"],3),"exchange",d["impl_config"]["exchange"],"parity",p.get("ok"),p.get("rel_2norm"),p.get("top_k_api_identical_to_host_argsort"), "reorth", (d.get("reorth_variant") or {}).get("value"), "finite", d["result_finite"])
except Exception as e:
    print("$wl N=8 FAILED", e); print(open("gpurun_out/bench_r02_${wl}_n8.err").read()[-2500:])
PY
done
