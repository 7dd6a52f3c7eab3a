"""Detailed device trace of ONE SpMV pass 0 at N GPUs: end time of every gatherer / sender CTA with its SM id (rank 0 prints)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
import bench
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
lz = g.load_package()
uid = None
if world > 1:
    dist.init_process_group("gloo")
    box = [lz.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
ctx = lz.Context(local, rank, world, uid)
ctx.graph_generate(bench.make_spec(lz, bench.WORKLOADS["c3"], None))
ctx.set_start_vector(None)
for _ in range(3):
    ctx.lanczos_run(6)
ctx.sync()
if world > 1:
    dist.barrier()
CAP = 65536
ctx.trace_on(CAP)
ctx.lanczos_run(6)
ev = ctx.trace_read(CAP)
if rank == 0:
    tag, t = ev[:, 0].astype(np.uint64), ev[:, 1].astype(np.int64)
    kern, ph = (tag >> np.uint64(8)) & np.uint64(255), tag & np.uint64(255)
    smid, cta = (tag >> np.uint64(16)) & np.uint64(255), tag >> np.uint64(24)
    starts = t[(kern == 0x10) & (ph == 1)]
    for step in (3, 4):
        t0 = np.sort(starts)[step]
        sel = (kern == 0x10) & (ph == 7) & (t > t0) & (t < t0 + 400000)
        sel_s = (kern == 0x10) & (ph == 8) & (t > t0 - 20000) & (t < t0 + 400000)
        e = (t[sel] - t0) / 1e3
        es = (t[sel_s] - t0) / 1e3
        print(f"step {step}: {sel.sum()} gatherer CTAs end at min {e.min():.1f} p10 {np.percentile(e,10):.1f} median {np.median(e):.1f} p90 {np.percentile(e,90):.1f} max {e.max():.1f} us; "
              f"{sel_s.sum()} sender CTAs end at min {es.min():.1f} median {np.median(es):.1f} max {es.max():.1f} us")
        sender_sms = set(int(x) for x in smid[sel_s])
        on = np.array([int(x) in sender_sms for x in smid[sel]])
        print(f"   gatherers on SMs that host a sender: {on.sum()} median end {np.median(e[on]) if on.any() else 0:.1f} max {e[on].max() if on.any() else 0:.1f}; "
              f"on other SMs: {(~on).sum()} median {np.median(e[~on]) if (~on).any() else 0:.1f} max {e[~on].max() if (~on).any() else 0:.1f}")
        # per-SM gatherer count
        cnt = np.bincount(smid[sel].astype(np.int64), minlength=148)
        print("   gatherer CTAs per SM: min", cnt.min(), "max", cnt.max(), " senders on", len(sender_sms), "SMs")
        late = e > np.percentile(e, 90)
        print("   the latest 10%: cta ids", np.sort(cta[sel][late].astype(np.int64))[:12], "... SMs", np.unique(smid[sel][late].astype(np.int64))[:20])
ctx.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
