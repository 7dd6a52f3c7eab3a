"""Small driver for ncu captures: C3 graph, one plain run (k=3) and one fp32-basis reorthogonalised run (k=10).
Kernel order per plain step: k_spmv_sell pass 0, pass 1, k_update_lagged."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
import bench
lz = g.load_package()
with lz.Context(0) as ctx:
    ctx.graph_generate(bench.make_spec(lz, bench.WORKLOADS["c3"], None))
    ctx.set_start_vector(None)
    ctx.lanczos_run(3); ctx.tridiag_expv(); ctx.multout(); ctx.top_k(100)
    ctx.set_basis_precision(lz.BASIS_F32)
    ctx.set_start_vector(None)
    ctx.lanczos_run(10, lz.REORTH_FULL); ctx.tridiag_expv(); ctx.multout()
    ctx.sync()
print("ncu target done")
