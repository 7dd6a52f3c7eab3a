"""Runs the reference's OWN CUDA path (parallel-final cu_decompose built for sm_100, oracle/_ref/ref_final --cuda) on the
workload graphs, next to ours, and checks the two answers against each other. Evidence for 'the kernels to beat'."""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import oracle as orc
lz = g.load_package()
for name, scale, k in (("c2", 20, 30), ("c3", 24, 50)):
    with lz.Context(0) as ctx:
        ctx.graph_generate(lz.GraphSpec.rmat(scale, 8, 1))
        ro, ci = ctx.csr_download()
        t0 = time.perf_counter(); y = ctx.expv_host(None, k); t_ours = time.perf_counter() - t0
        t0 = time.perf_counter(); y = ctx.expv_host(None, k); t_ours = min(t_ours, time.perf_counter() - t0)
    p = os.path.join(tempfile.gettempdir(), f"ref_{name}.bin")
    lz.write_bin(p, ro, ci)
    r = orc.run_ref_final(None, None, k, cuda=True, reps=2, csr_path=p)
    os.unlink(p)
    rel = np.linalg.norm(y - r["ans"]) / np.linalg.norm(r["ans"])
    same = np.array_equal(orc.top_k(y), orc.top_k(r["ans"]))
    print(json.dumps({"workload": name, "n": len(ro) - 1, "nnz": int(ro[-1]), "k": k, "reference_cuda_runs": r["timings"],
                      "ours_expv_host_s": t_ours, "rel_2norm_ours_vs_reference_cuda": rel, "top100_identical": bool(same)}), flush=True)
