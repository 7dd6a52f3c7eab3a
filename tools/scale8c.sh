run() { n=$1; tag=$2; shift 2; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) bench.py --gpus $n --steps 10 --warmup 3 --no-reorth-detail > gpurun_out/s8_$tag.log 2>gpurun_out/s8_$tag.err; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/s8_$tag.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$tag", "N=",d["n_gpus"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "rank", round(d["e2e_rank"]["value"],1), "spmv_ms", round(d["detail"]["spmv_ms_avg"],4), "upd_ms", round(d["detail"]["update_scale_ms_per_iter"],4), "parity_ok", d["parity"]["ok"], d["parity"]["rel_2norm"])
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/s8_$tag.err").read()[-1500:])
PY
}
run 8 n8 A=1
run 8 n8_p37 LZ_PUSH_CTAS=37
run 4 n4 A=1
bash tools/run_n.sh 8 tools/trace_step.py 2>&1 | grep "rank 0" -A14
