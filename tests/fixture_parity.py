"""Parity of a full-size e^A·x result against the committed summary fixtures of the UNMODIFIED reference
(tests/golden/rmat_s{scale}_k{k}_summary.npz, written by tests/golden/make_golden_c3.py from oracle/_ref/ref_final).

Pure numpy on host data — a checker for tests/ and for bench.py's `parity` object; nothing here is on the product path and
nothing here calls the oracle. The fixture holds the reference's alpha/beta, ||y||, its 256 largest entries, 4096 seeded
sample entries and 1024 block sums of y, so every entry of the reference answer is represented."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9   # BASELINE.json north_star: relative 2-norm in fp64


def fixture_path(kind, scale, k, seed=1, ef=8):
    if kind != "rmat" or seed != 1 or ef != 8:
        return None
    p = os.path.join(GOLDEN, f"rmat_s{scale}_k{k}_summary.npz")
    return p if os.path.exists(p) else None


def top_order(y, m):
    """argsort(-y) with ties -> lower index, first m (host-side restatement of the ranking definition)."""
    y = np.asarray(y)
    m = min(m, len(y))
    part = np.argpartition(-y, m - 1)[:m] if m < len(y) else np.arange(len(y))
    thr = y[part].min()
    cand = np.flatnonzero(y >= thr)                      # everything tied with the m-th value takes part in the tie-break
    return cand[np.lexsort((cand, -y[cand]))][:m].astype(np.uint32)


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def compare(y, path, alpha=None, beta=None, top_idx_product=None, top_val_product=None):
    """-> dict for bench.py / asserts in tests. `y` = full answer on the host (original vertex order)."""
    g = np.load(path)
    n = int(g["n"])
    assert len(y) == n, (len(y), n)
    meta = json.loads(str(g["meta"]))
    sidx, tidx = g["sample_idx"].astype(np.int64), g["top_idx"].astype(np.int64)
    edges = np.linspace(0, n, len(g["block_sums"]) + 1).astype(np.int64)
    entries = np.concatenate([sidx, tidx])
    ref_entries = np.concatenate([g["sample_val"], g["top_val"]])
    out = {
        "fixture": os.path.basename(path), "source": meta["source"], "entries_compared": int(len(entries)),
        "rel_2norm": _rel(y[entries], ref_entries),                                  # sampled + top entries
        "rel_2norm_block_sums": _rel(np.add.reduceat(y, edges[:-1]), g["block_sums"]),   # every entry contributes
        "rel_norm2": float(abs(np.linalg.norm(y) - float(g["norm2"])) / float(g["norm2"])),
        "top100_identical": bool(np.array_equal(top_order(y, 100), g["top_idx"][:100])),
        "top256_identical": bool(np.array_equal(top_order(y, 256), g["top_idx"])),
        "top_gap": float(g["top_gap"]),
    }
    if top_idx_product is not None:
        m = len(top_idx_product)
        out["top_k_api_identical"] = bool(np.array_equal(np.asarray(top_idx_product), g["top_idx"][:m]))
        if top_val_product is not None:
            out["top_k_api_rel_values"] = _rel(top_val_product, g["top_val"][:m])
    if alpha is not None:
        lead = min(6, len(alpha), len(g["alpha"]))
        out["alpha_lead_rel"] = float(np.max(np.abs(alpha[:lead] - g["alpha"][:lead]) / np.abs(g["alpha"][:lead])))
    if beta is not None and len(beta):
        lead = min(5, len(beta), len(g["beta"]))
        out["beta_lead_rel"] = float(np.max(np.abs(beta[:lead] - g["beta"][:lead]) / np.abs(g["beta"][:lead])))
    out["ok"] = bool(out["rel_2norm"] < TOL and out["rel_2norm_block_sums"] < TOL and out["rel_norm2"] < TOL and out["top100_identical"]
                     and out.get("top_k_api_identical", True))
    return out
