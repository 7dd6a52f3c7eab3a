# usage: bash tools/gpu_c4.sh N  — C4 (R-MAT 2^27, k=50) on N GPUs, parity against the committed 1-GPU summary
n=$1
bash tools/run_n.sh $n bench.py --gpus $n --workload c4 --steps 3 --warmup 3 > gpurun_out/bench_r02_c4_n$n.log 2>gpurun_out/bench_r02_c4_n$n.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_r02_c4_n$n.log") if l.startswith("{")][-1])
    p=d["parity"]
    print("C4 N=$n value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"rank",round(d["e2e_rank"]["value"],1),"spmv_ms",round(d["detail"]["spmv_ms_avg"],3),"upd_ms",round(d["detail"]["update_scale_ms_per_iter"],3),"graph_build_s",round(d["detail"]["graph_build_s"],2),"parity ok",p.get("ok"),p.get("rel_2norm"),p.get("top100_identical"),p.get("top_k_api_identical"))
except Exception as e:
    print("C4 N=$n FAILED", e); print(open("gpurun_out/bench_r02_c4_n$n.err").read()[-2500:])
PY
