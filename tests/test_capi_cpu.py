"""CPU-side checks of the C ABI: the library loads, exports every symbol include/lz.h declares, the host-only entry
points (generators, file formats) behave, and GPU entry points fail loudly instead of falling back."""
import ctypes
import os

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(lz):
    declared = lz.header_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lz.lib, s)]
    assert not missing, f"declared in include/lz.h but not exported: {missing}"
    assert lz.lib.lz_version() >= 100


def test_no_oracle_or_cpu_fallback_in_product(lz):
    """The product must not link or load anything under oracle/."""
    import subprocess
    out = subprocess.run(["ldd", lz.lib_path], capture_output=True, text=True).stdout
    assert "lzoracle" not in out and "oracle" not in out
    src_dir = os.path.join(os.path.dirname(lz.lib_path))
    for f in os.listdir(src_dir):
        if f.endswith((".cu", ".cc", ".h")):
            txt = open(os.path.join(src_dir, f)).read()
            assert "lanczos_oracle" not in txt and "oracle/" not in txt, f


def test_host_generators_are_deterministic_and_simple(lz):
    for spec in (lz.GraphSpec.er(5000, 20000, 42), lz.GraphSpec.rmat(11, 8, 42), lz.GraphSpec.band(5000, 42)):
        n, ro, ci = lz.generate_host(spec)
        n2, ro2, ci2 = lz.generate_host(spec)
        assert n == n2 and np.array_equal(ro, ro2) and np.array_equal(ci, ci2)
        assert ro[0] == 0 and ro[-1] == len(ci) and np.all(np.diff(ro.astype(np.int64)) >= 0)
        rows = np.repeat(np.arange(n), np.diff(ro))
        assert not np.any(rows == ci), "self loop"
        key = rows.astype(np.int64) * n + ci
        assert np.all(np.diff(key) > 0), "columns must be strictly ascending within rows (sorted, unique)"
        tkey = ci.astype(np.int64) * n + rows
        assert np.array_equal(np.sort(tkey), key), "not symmetric"
        assert ro[n] > ro[n - 1], "last vertex must not be isolated"
    # different seeds -> different graphs
    a = lz.generate_host(lz.GraphSpec.rmat(11, 8, 1))[2]
    b = lz.generate_host(lz.GraphSpec.rmat(11, 8, 2))[2]
    assert len(a) != len(b) or not np.array_equal(a, b)


def test_rmat_shape_matches_config(lz):
    n, ro, ci = lz.generate_host(lz.GraphSpec.rmat(16, 8, 1))
    assert n == 1 << 16
    assert 15.0 < len(ci) / n < 16.0                 # "~16 nnz/row" after symmetrise / de-dup
    deg = np.diff(ro)
    assert deg.max() > 8 * deg.mean()                # skewed


def test_band_is_irregular_and_low_degree(lz):
    n, ro, ci = lz.generate_host(lz.GraphSpec.band(1 << 14, 5))
    deg = np.diff(ro)
    assert 3.5 < deg.mean() < 4.2 and len(np.unique(deg)) >= 4   # ~4 nnz/row, NOT regular


def test_text_and_binary_roundtrip(lz, tmp_path):
    n, ro, ci = lz.generate_host(lz.GraphSpec.er(300, 900, 7))
    t, b = str(tmp_path / "g.mtx"), str(tmp_path / "g.bin")
    lz.write_text(t, ro, ci)
    lz.write_bin(b, ro, ci)
    first = open(t).readline().split()
    assert first == [str(n), str(n), str(len(ci) // 2)]          # "n n E" header (adjMatrix.cc:59)
    for reader, path in ((lz.read_text, t), (lz.read_bin, b)):
        n2, ro2, ci2 = reader(path)
        assert n2 == n and np.array_equal(ro2, ro) and np.array_equal(ci2, ci)


def test_text_reader_symmetrises_and_dedups(lz, tmp_path):
    p = tmp_path / "tiny.mtx"
    p.write_text("4 4 4\n2 1\n3 1\n1 2\n4 3\n")                   # edge (1,2) listed twice, in both orders
    n, ro, ci = lz.read_text(str(p))
    assert n == 4 and ro.tolist() == [0, 2, 3, 5, 6] and ci.tolist() == [1, 2, 0, 0, 3, 2]


def test_errors_are_reported_not_swallowed(lz, tmp_path):
    with pytest.raises(lz.LzError) as e:
        lz.read_text(str(tmp_path / "nope.mtx"))
    assert e.value.code == -5
    with pytest.raises(lz.LzError):
        lz.generate_host(lz.GraphSpec(99, 0, 10, 10, 1, 0, 0, 0))
    bad = tmp_path / "bad.mtx"
    bad.write_text("3 3 2\n1 2\n")
    with pytest.raises(lz.LzError):
        lz.read_text(str(bad))


def test_gpu_entry_points_fail_loudly_without_a_device(lz):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lz.LzError) as e:
        lz.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_text_reader_accepts_matrix_market_files(lz, tmp_path):
    """The reference's input format is MatrixMarket 'coordinate' with the banner stripped (serial/README.md:9). The reader takes
    both: '%' lines are skipped, a value column is ignored; stripped and unstripped files give the same CSR."""
    import numpy as np
    n, ro, ci = lz.generate_host(lz.GraphSpec.er(300, 900, 4))
    plain, mm = str(tmp_path / "g.txt"), str(tmp_path / "g.mtx")
    lz.write_text(plain, ro, ci)
    body = open(plain).read().splitlines()
    with open(mm, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n% a comment\n\n" + body[0] + "\n")
        for i, ln in enumerate(body[1:]):
            f.write(ln + " 1.0\n")
            if i == 3:
                f.write("% comment between entries\n")
    n1, ro1, ci1 = lz.read_text(plain)
    n2, ro2, ci2 = lz.read_text(mm)
    assert n1 == n2 == n and np.array_equal(ro1, ro) and np.array_equal(ci1, ci) and np.array_equal(ro2, ro) and np.array_equal(ci2, ci)
    bad = str(tmp_path / "bad.mtx")
    open(bad, "w").write("% only comments\n")
    import pytest
    with pytest.raises(lz.LzError):
        lz.read_text(bad)
