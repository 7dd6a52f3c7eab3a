"""Run as a subprocess with LZ_REORTH_ALWAYS_TWICE=1 (the knob is read once per process): forces the second Gram-Schmidt pass
every step, so the rarely-taken path of LZ_REORTH_FULL is exercised; prints second_passes and the answer's hash-free summary."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

lz = g.load_package()
gl = np.load(os.path.join(ROOT, "tests", "golden", "rmat_s12_k30.npz"))
with lz.Context(0) as c:
    c.csr_upload(gl["row_offset"], gl["col_idx"])
    y = c.expv_host(None, int(gl["k"]), lz.REORTH_FULL)
    t = c.timings()
    Q = np.stack([c.get_basis(j) for j in range(int(gl["k"]))])
rel = float(np.linalg.norm(y - gl["ans"]) / np.linalg.norm(gl["ans"]))
print(json.dumps({"second_passes": int(t.reorth_second_passes), "rel": rel, "orth": float(np.abs(Q @ Q.T - np.eye(len(Q))).max())}))
