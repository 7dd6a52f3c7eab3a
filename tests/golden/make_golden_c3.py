"""Summary fixture of the UNMODIFIED reference on BASELINE.json configs[2] (C3: R-MAT 2^24, edge factor 8, seed 1, k=50).

    python tests/golden/make_golden_c3.py [--scale 24] [--k 50]

The full answer is 134 MB, so only a summary is committed (tests/golden/rmat_s24_k50_summary.npz; configs[1] = --scale 20 --k 30):
  alpha[k], beta[k-1]            the reference's tridiagonal (parallel-final host path, lanczosDecomp<double>(A,k,ones,false))
  norm2                          ||e^A 1||_2
  top_idx[256], top_val[256]     the 256 largest entries (argsort(-y), ties -> lower index) — top-100 ranking + margin
  top_gap                        smallest relative gap between consecutive entries of the top-101
  sample_idx[4096], sample_val   seeded random entries of y
  block_sums[1024]               sums of y over 1024 equal index blocks (every entry of y contributes to the fixture)
Generated in the build container from our deterministic host generator (bit-identical to the device generator:
tests/test_gpu_parity.py::test_device_generator_matches_host_generator) and oracle/_ref/ref_final (make -C oracle ref).
Needs ~12 GB of host memory and a few minutes of one core. The GPU box has no /root/reference: tests and bench.py only
read the .npz."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import __graft_entry__ as g  # noqa: E402
import oracle  # noqa: E402

TOP, SAMPLES, BLOCKS = 256, 4096, 1024


def summarize(y, alpha, beta, meta):
    n = len(y)
    order = np.lexsort((np.arange(n), -y))[:TOP].astype(np.uint32)      # descending value, ties -> lower index
    rng = np.random.default_rng(20261018)
    sidx = np.sort(rng.choice(n, SAMPLES, replace=False)).astype(np.uint32)
    s = y[order][:101]
    edges = np.linspace(0, n, BLOCKS + 1).astype(np.int64)
    return dict(alpha=alpha, beta=beta, norm2=np.float64(np.linalg.norm(y)), top_idx=order, top_val=y[order],
                top_gap=np.float64(np.min((s[:-1] - s[1:]) / np.abs(s[:-1]))), sample_idx=sidx, sample_val=y[sidx],
                block_sums=np.add.reduceat(y, edges[:-1]), n=np.uint64(n), meta=json.dumps(meta))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=24)
    ap.add_argument("--k", type=int, default=50)
    args = ap.parse_args()
    assert oracle.have_ref(), "build the reference first: make -C oracle ref"
    lz = g.load_package()
    spec = lz.GraphSpec.rmat(args.scale, 8, 1)
    t0 = time.time()
    n, ro, ci = lz.generate_host(spec)
    print(f"graph: n={n} nnz={len(ci)} ({time.time() - t0:.1f} s)", flush=True)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "g.bin")
        lz.write_bin(path, ro, ci)
        nnz = len(ci)
        del ro, ci
        t0 = time.time()
        r = oracle.run_ref_final(None, None, args.k, csr_path=path)
        print(f"reference: {r['timings']} ({time.time() - t0:.1f} s)", flush=True)
    y = r["ans"]
    assert np.isfinite(y).all()
    meta = {"graph": {"kind": "rmat", "scale": args.scale, "edge_factor": 8, "seed": 1}, "k": args.k, "x": "ones", "nnz": nnz,
            "source": "oracle/_ref/ref_final --csr (unmodified reference, parallel-final host path, cuda=false)",
            "timings": r["timings"]}
    out = os.path.join(HERE, f"rmat_s{args.scale}_k{args.k}_summary.npz")
    np.savez_compressed(out, **summarize(y, r["alpha"], r["beta"], meta))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
