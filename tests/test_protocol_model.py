"""Randomised model check of the multi-GPU exchange protocol — no GPU, no library code.

compute-sanitizer's racecheck is closed on this pool, and the hand-rolled NVLink protocol (chunk pushes with release/acquire
arrival counters, two-parity scalar slots, sender CTAs fused into the SpMV passes, the lagged update that publishes ||u||^2
one step ahead) is argued correct in comments. This test states that argument as an executable model: every rank runs the
exact sequence of operations `enqueue_steps` issues for the lagged multi-GPU loop, a random scheduler interleaves the ranks at
the granularity of single peer stores / flag writes / slot reads, and the model asserts the two properties the argument claims:

  (1) a chunk of a rank's gathered vector is never overwritten while that rank still gathers the previous version from it,
      and a gather only ever sees the version it waited for (no write-after-read, no torn read);
  (2) a scalar slot (kind, parity) is never overwritten before every rank consumed the value it held (two parities suffice),
      and the schedule never deadlocks.

It also shows the model has teeth: dropping the alpha wait from the update kernel (so a rank can run ahead) makes (1) fail.
"""
import random

import pytest


class Violation(AssertionError):
    pass


class Rank:
    def __init__(self, r, world, chunks):
        self.r, self.world = r, world
        self.data = [[0] * world for _ in range(chunks)]      # data[c][src] = version of chunk c written by src
        self.flag = [[0] * world for _ in range(chunks)]      # arrival counters
        self.slot = [[[0] * world for _ in range(2)] for _ in range(2)]   # slot[kind][parity][src] = seq
        self.consumed = [[[0] * world for _ in range(2)] for _ in range(2)]   # last seq this rank consumed from that slot
        self.reading = {}                                      # chunk -> version being gathered right now
        self.done_reading = [0] * chunks                       # highest version whose gather of chunk c has completed
        self.pc = 0
        self.prog = []
        self.sub = None                                        # micro-state of the current operation


def build_program(k, chunks, fused_push, wait_alpha=True):
    """Operations of one rank for a k-step lagged run; versions: the vector of step j has version j + 1; reductions seq = j + 1."""
    prog = [("push", list(range(1 if fused_push else chunks)), 1)]          # q_0: chunk 0 (fused) or all chunks
    for j in range(k):
        v, s = j + 1, j + 1
        for b in range(chunks):
            send = [b + 1] if (fused_push and b + 1 < chunks) else []
            prog.append(("pass", b, v, send))                               # gatherers wait for chunk b; senders push chunk b + 1
        prog.append(("publish", 0, s))                                      # alpha partial (last pass's last CTA)
        cons = ([(1, s - 1)] if j else []) + ([(0, s)] if wait_alpha else [])
        if j + 1 == k:
            prog.append(("consume", cons))                                  # k_lagged_finish
            break
        prog.append(("consume", cons))                                      # k_update_lagged_push: ||u_j||^2, then alpha
        prog.append(("push", list(range(1 if fused_push else chunks)), v + 1))
        prog.append(("publish", 1, s))                                      # partial of ||u_{j+1}||^2
    return prog


def store_chunk(ranks, src, dst, c, v):
    d = ranks[dst]
    if c in d.reading:
        raise Violation(f"rank {src} overwrites chunk {c} (v{v}) of rank {dst} while it gathers v{d.reading[c]}")
    if d.done_reading[c] < v - 1:
        raise Violation(f"rank {src} writes chunk {c} v{v} into rank {dst} before it finished gathering v{v - 1}")
    d.data[c][src] = v


def step(ranks, me, rng):
    """Advance rank `me` by one micro-operation. Returns False if it is blocked."""
    op = me.prog[me.pc]
    kind = op[0]
    W = me.world
    if kind == "push":
        _, cs, v = op
        if me.sub is None:
            me.sub = {"stores": [(c, d) for c in cs for d in range(W)], "flags": [(c, d) for c in cs for d in range(W)]}
            rng.shuffle(me.sub["stores"])
        if me.sub["stores"]:
            c, d = me.sub["stores"].pop()
            store_chunk(ranks, me.r, d, c, v)
            return True
        if me.sub["flags"]:                                    # all stores fenced, then the release stores of the counters
            c, d = me.sub["flags"].pop()
            ranks[d].flag[c][me.r] = v
            return True
        me.sub = None
        me.pc += 1
        return True
    if kind == "pass":
        _, b, v, send = op
        if me.sub is None:
            me.sub = {"stores": [(c, d) for c in send for d in range(W)], "flags": [(c, d) for c in send for d in range(W)],
                      "gather": "wait"}
            rng.shuffle(me.sub["stores"])
        choices = []
        if me.sub["stores"] or me.sub["flags"]:
            choices.append("send")
        if me.sub["gather"] == "wait" and all(f >= v for f in me.flag[b]):
            choices.append("begin")
        if me.sub["gather"] == "reading":
            choices.append("end")
        if not choices:
            if me.sub["gather"] == "done":
                me.sub = None
                me.pc += 1
                return True
            return False                                        # gatherers spin on the arrival counters, senders are done
        what = rng.choice(choices)
        if what == "send":
            if me.sub["stores"]:
                c, d = me.sub["stores"].pop()
                store_chunk(ranks, me.r, d, c, v)
            else:
                c, d = me.sub["flags"].pop()
                ranks[d].flag[c][me.r] = v
        elif what == "begin":
            if any(x != v for x in me.data[b]):
                raise Violation(f"rank {me.r} gathers chunk {b}: expected v{v}, sees {me.data[b]}")
            me.reading[b] = v
            me.sub["gather"] = "reading"
        else:
            if any(x != v for x in me.data[b]):
                raise Violation(f"rank {me.r}: chunk {b} changed under the gather of v{v}: {me.data[b]}")
            del me.reading[b]
            me.done_reading[b] = v
            me.sub["gather"] = "done"
        return True
    if kind == "publish":
        _, kd, s = op
        if me.sub is None:
            me.sub = list(range(W))
            rng.shuffle(me.sub)
        if me.sub:
            d = me.sub.pop()
            old = ranks[d].slot[kd][s & 1][me.r]
            if old and ranks[d].consumed[kd][s & 1][me.r] < old:
                raise Violation(f"rank {me.r} overwrites slot kind {kd} parity {s & 1} on rank {d} (seq {old}) before it was consumed")
            ranks[d].slot[kd][s & 1][me.r] = s
            return True
        me.sub = None
        me.pc += 1
        return True
    if kind == "consume":
        # every CTA of the consuming kernel reads the slots on its own, at different times: model the first and the last reader
        if me.sub is None:
            for kd, s in op[1]:
                if any(x < s for x in me.slot[kd][s & 1]):
                    return False
            me.sub = "first_read_done"
        for kd, s in op[1]:
            if any(x != s for x in me.slot[kd][s & 1]):
                raise Violation(f"rank {me.r} consumes kind {kd} seq {s} but the slot holds {me.slot[kd][s & 1]} ({me.sub})")
        if me.sub == "first_read_done":
            me.sub = "last_read"
            return True
        for kd, s in op[1]:
            for src in range(W):
                me.consumed[kd][s & 1][src] = s
        me.sub = None
        me.pc += 1
        return True
    raise AssertionError(kind)


def run_model(world, chunks, k, fused_push, seed, wait_alpha=True, bias=None):
    rng = random.Random(seed)
    ranks = [Rank(r, world, chunks) for r in range(world)]
    for rk in ranks:
        rk.prog = build_program(k, chunks, fused_push, wait_alpha)
    weights = [1.0] * world
    if bias is not None:                                        # one rank much faster / slower than the others
        weights[bias[0]] = bias[1]
    while True:
        live = [rk for rk in ranks if rk.pc < len(rk.prog)]
        if not live:
            return
        order = sorted(live, key=lambda rk: rng.random() / weights[rk.r])
        for rk in order:
            if step(ranks, rk, rng):
                break
        else:
            raise Violation("deadlock: " + ", ".join(f"rank {rk.r} at {rk.prog[rk.pc][:3]}" for rk in live))


@pytest.mark.parametrize("world,chunks,fused", [(2, 1, False), (2, 2, True), (4, 2, True), (8, 2, True), (4, 3, True), (3, 4, False), (8, 16, True)])
def test_exchange_protocol_has_no_hazard_under_random_schedules(world, chunks, fused):
    for seed in range(12):
        run_model(world, chunks, k=5, fused_push=fused, seed=seed)
    for fast in (0, world - 1):                                 # a rank that runs far ahead, and one that lags
        run_model(world, chunks, k=5, fused_push=fused, seed=99, bias=(fast, 50.0))
        run_model(world, chunks, k=5, fused_push=fused, seed=98, bias=(fast, 0.02))


def test_model_detects_a_broken_protocol():
    """Without the alpha wait a rank can push the next vector while a peer still gathers the current one: the model must see it."""
    caught = 0
    for seed in range(40):
        try:
            run_model(4, 2, k=5, fused_push=True, seed=seed, wait_alpha=False, bias=(0, 50.0))
        except Violation:
            caught += 1
    assert caught > 0
