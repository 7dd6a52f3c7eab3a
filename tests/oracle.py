"""ctypes face of the CPU oracle (oracle/lanczos_oracle.c) and runner for the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs. The product never imports this."""
import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "liblzoracle.so")
REF_FINAL = os.path.join(ORACLE_DIR, "_ref", "ref_final")
REF_SERIAL = os.path.join(ORACLE_DIR, "_ref", "ref_serial")

if not os.path.exists(LIB):
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])
_lib = C.CDLL(LIB)
_u32p, _f64p = C.POINTER(C.c_uint32), C.POINTER(C.c_double)


def _u32(a):
    return a.ctypes.data_as(_u32p)


def _f64(a):
    return a.ctypes.data_as(_f64p)


_lib.lzo_spmv.argtypes = [C.c_uint32, _u32p, _u32p, _f64p, _f64p]
_lib.lzo_spmv.restype = None
_lib.lzo_expv.argtypes = [C.c_uint32, _u32p, _u32p, C.c_uint32, _f64p, C.c_int, _f64p, _f64p, _f64p]
_lib.lzo_expv.restype = C.c_int
for _n in ("lzo_lanczos", "lzo_lanczos_arnoldi", "lzo_lanczos_fullreorth"):
    getattr(_lib, _n).argtypes = [C.c_uint32, _u32p, _u32p, C.c_uint32, _f64p, _f64p, _f64p, _f64p]
    getattr(_lib, _n).restype = C.c_int
_lib.lzo_tridiag_eig.argtypes = [C.c_uint32, _f64p, _f64p, _f64p]
_lib.lzo_tridiag_eig.restype = C.c_int
_lib.lzo_multout.argtypes = [C.c_uint32, C.c_uint32, _f64p, _f64p, _f64p, C.c_double, C.c_int, _f64p, _f64p]
_lib.lzo_multout.restype = None
_lib.lzo_check_ans.argtypes = [C.c_uint32, _f64p, _f64p, _f64p, _u32p, _f64p, _f64p]
_lib.lzo_check_ans.restype = None
_lib.lzo_top_k.argtypes = [C.c_uint32, _f64p, C.c_uint32, _u32p]
_lib.lzo_top_k.restype = None
_lib.lzo_norm.argtypes = [_f64p, C.c_uint32]
_lib.lzo_norm.restype = C.c_double

PLAIN, ARNOLDI, FULL = 0, 1, 2


def _csr(ro, ci):
    return np.ascontiguousarray(ro, np.uint32), np.ascontiguousarray(ci, np.uint32)


def spmv(ro, ci, x):
    ro, ci = _csr(ro, ci)
    x = np.ascontiguousarray(x, np.float64)
    y = np.empty_like(x)
    _lib.lzo_spmv(len(ro) - 1, _u32(ro), _u32(ci), _f64(x), _f64(y))
    return y


def expv(ro, ci, k, x, reorth=PLAIN):
    """-> ans, alpha, beta  (reference pipeline parallel-final/main.cu:83-93)"""
    ro, ci = _csr(ro, ci)
    n = len(ro) - 1
    x = np.ascontiguousarray(x, np.float64)
    ans, a, b = np.empty(n), np.empty(k), np.empty(max(k - 1, 1))
    rc = _lib.lzo_expv(n, _u32(ro), _u32(ci), k, _f64(x), reorth, _f64(ans), _f64(a), _f64(b))
    assert rc == 0, f"oracle eigensolver failed ({rc})"
    return ans, a, b[:k - 1]


def lanczos(ro, ci, k, x, reorth=PLAIN):
    """-> alpha, beta, Q (n x k row-major)"""
    ro, ci = _csr(ro, ci)
    n = len(ro) - 1
    x = np.ascontiguousarray(x, np.float64)
    a, b, Q = np.empty(k), np.empty(max(k - 1, 1)), np.empty((n, k))
    fn = {PLAIN: _lib.lzo_lanczos, ARNOLDI: _lib.lzo_lanczos_arnoldi, FULL: _lib.lzo_lanczos_fullreorth}[reorth]
    fn(n, _u32(ro), _u32(ci), k, _f64(x), _f64(a), _f64(b), _f64(Q))
    return a, b[:k - 1], Q


def tridiag_eig(alpha, beta):
    """-> eigenvalues ascending, Z (k x k, Z[i, j] = component i of vector j) — dstevd's contract"""
    d = np.array(alpha, np.float64)
    k = len(d)
    e = np.zeros(max(k - 1, 1))
    e[:k - 1] = beta
    Z = np.empty((k, k))
    rc = _lib.lzo_tridiag_eig(k, _f64(d), _f64(e), _f64(Z))
    assert rc == 0
    return d, Z


def multout(eigvals, Z, Q, x_norm, qtrans=False):
    k = len(eigvals)
    Q = np.ascontiguousarray(Q, np.float64)
    n = Q.shape[1] if qtrans else Q.shape[0]
    ans, coef = np.empty(n), np.empty(k)
    ev, Zc = np.ascontiguousarray(eigvals, np.float64), np.ascontiguousarray(Z, np.float64)
    _lib.lzo_multout(n, k, _f64(ev), _f64(Zc), _f64(Q), x_norm, int(qtrans), _f64(ans), _f64(coef))
    return ans, coef


def check_ans(a, b):
    """-> max|a-b|, argmax, ||a-b||, ||a-b||/||b||  (parallel-final/lib/check_ans.cu:12-29)"""
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    mx, nd, rel, mi = C.c_double(), C.c_double(), C.c_double(), C.c_uint32()
    _lib.lzo_check_ans(len(a), _f64(a), _f64(b), C.byref(mx), C.byref(mi), C.byref(nd), C.byref(rel))
    return mx.value, mi.value, nd.value, rel.value


def top_k(y, top=100):
    y = np.ascontiguousarray(y, np.float64)
    out = np.empty(min(top, len(y)), np.uint32)
    _lib.lzo_top_k(len(y), _f64(y), len(out), _u32(out))
    return out


def top_gap(y, top=100):
    """Smallest relative gap between consecutive entries of the top-(top+1) — the ranking claim needs this >> 1e-9."""
    s = np.sort(np.asarray(y))[::-1][:top + 1]
    return float(np.min((s[:-1] - s[1:]) / np.abs(s[:-1])))


# ---- compiled reference (oracle/_ref) ------------------------------------------------------------------------------
def have_ref():
    return os.path.exists(REF_FINAL)


def write_csr_bin(path, ro, ci):
    ro, ci = _csr(ro, ci)
    with open(path, "wb") as f:
        f.write(b"LZCSR1\0\0")
        np.array([len(ro) - 1, len(ci)], np.uint64).tofile(f)
        ro.tofile(f)
        ci.tofile(f)


def run_ref_final(ro, ci, k, x=None, cuda=False, iters=0, reps=1, want_output=True, csr_path=None, single=False):
    """Runs the unmodified reference (parallel-final host path, or its CUDA path with cuda=True).
    -> dict(ans, alpha, beta, timings=[json per rep])"""
    with tempfile.TemporaryDirectory() as td:
        if csr_path is None:
            csr_path = os.path.join(td, "g.bin")
            write_csr_bin(csr_path, ro, ci)
        cmd = [REF_FINAL, "--csr", csr_path, "-k", str(k), "--reps", str(reps)]
        if x is not None:
            xp = os.path.join(td, "x.f64")
            np.ascontiguousarray(x, np.float64).tofile(xp)
            cmd += ["--x", xp]
        if cuda:
            cmd.append("--cuda")
        if single:
            cmd.append("--float")          # the reference's lanczosDecomp<float> instantiation
        if iters:
            cmd += ["--iters", str(iters)]
        out = os.path.join(td, "o")
        if want_output and not iters:
            cmd += ["--out", out]
        res = subprocess.run(cmd, check=True, capture_output=True, text=True)
        tim = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
        r = {"timings": tim}
        if want_output and not iters:
            r["ans"] = np.fromfile(out + ".ans.f64")
            r["alpha"] = np.fromfile(out + ".alpha.f64")
            r["beta"] = np.fromfile(out + ".beta.f64")
        return r
