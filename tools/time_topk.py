"""Times lz_top_k(100) on the C3 answer (device time via the ctx stopwatch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
import bench
lz = g.load_package()
with lz.Context(0) as ctx:
    ctx.graph_generate(bench.make_spec(lz, bench.WORKLOADS["c3"], None))
    ctx.expv_host(None, 10)
    for m in (100, 1024):
        ctx.top_k(m)
        ctx.timer_start()
        for _ in range(10):
            ctx.top_k(m)
        print(f"lz_top_k({m}) on n = 2^24: {ctx.timer_stop() / 10:.3f} ms per call (12 radix-select passes + collect + sort + D2H of {m} pairs)")
