"""Per-step timeline of the multi-GPU Lanczos loop from the device-side trace (lz_debug_trace): how long each kernel waits for
its peers (chunk arrival, alpha) and how long it works. Run under torchrun:
    python -m torch.distributed.run --nproc-per-node N tools/trace_step.py [--workload c3] [--k 20]
Prints, per rank, the mean over the traced steps of every phase in microseconds (globaltimer is per GPU: durations only)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402

PH = {1: "start", 2: "waited", 3: "pushed", 4: "end", 5: "send_start", 6: "send_end"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--scale", type=int, default=None)
    ap.add_argument("--k", type=int, default=20)
    a = ap.parse_args()
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
    ctx.graph_generate(bench.make_spec(lz, bench.WORKLOADS[a.workload], a.scale))
    ctx.set_start_vector(None)
    for _ in range(3):
        ctx.lanczos_run(a.k)
    ctx.sync()
    if world > 1:
        dist.barrier()
    ctx.trace_on(8192)
    ctx.timer_start()
    ctx.lanczos_run(a.k)
    ms = ctx.timer_stop()
    ev = ctx.trace_read(8192)
    tags, t = ev[:, 0].astype(np.int64), ev[:, 1].astype(np.int64)
    order = np.argsort(t, kind="stable")
    tags, t = tags[order], t[order]
    # split into steps at every update-kernel end (or last SpMV pass end on one GPU)
    lines = [f"rank {rank}: {a.k} steps in {ms:.3f} ms = {1e3 * ms / a.k:.1f} us/step, {len(ev)} events"]
    acc = {}
    prev_end = None
    for tg, ts in zip(tags, t):
        kern, ph = tg >> 8, tg & 255
        name = "update" if kern == 1 else f"spmv{kern - 0x10}"
        key = (name, PH.get(ph, str(ph)))
        acc.setdefault(key, []).append(ts)
    def dur(a_, b_):
        x, y = np.array(acc.get(a_, [])), np.array(acc.get(b_, []))
        m = min(len(x), len(y))
        if m < 3:
            return None
        d = (y[-m:] - x[-m:])[2:]          # skip the first steps
        return float(np.mean(d)) / 1e3, float(np.max(d)) / 1e3
    names = sorted({k[0] for k in acc})
    for nm in names:
        for a_, b_, what in (("start", "waited", "wait for peers"), ("waited", "end", "work after the wait"), ("waited", "pushed", "until chunk 0 pushed"),
                             ("send_start", "send_end", "sender CTAs (next chunk)"), ("start", "end", "whole kernel")):
            d = dur((nm, a_), (nm, b_))
            if d:
                lines.append(f"   {nm:8s} {what:28s} mean {d[0]:8.1f} us   max {d[1]:8.1f} us")
    # gaps between consecutive kernels
    seq = [(ts, tg) for tg, ts in zip(tags, t) if (tg & 255) in (1, 4)]
    gaps = {}
    for (t0, g0), (t1, g1) in zip(seq[:-1], seq[1:]):
        if (g0 & 255) == 4 and (g1 & 255) == 1:
            gaps.setdefault((g0 >> 8, g1 >> 8), []).append((t1 - t0) / 1e3)
    for (k0, k1), v in sorted(gaps.items()):
        n0 = "update" if k0 == 1 else f"spmv{k0 - 0x10}"
        n1 = "update" if k1 == 1 else f"spmv{k1 - 0x10}"
        lines.append(f"   gap {n0} end -> {n1} start: mean {np.mean(v[2:]):6.1f} us")
    out = "\n".join(lines)
    if world > 1:
        allo = [None] * world
        dist.all_gather_object(allo, out)
        if rank == 0:
            print("\n".join(allo[:2] + allo[-1:]), flush=True)
        dist.destroy_process_group()
    else:
        print(out, flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
