"""N > 1 path on real GPUs (skipped when the box has one GPU): torchrun over 2 (and 4) ranks, NCCL inside liblzb200."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_sharded_parity(lz, world):
    if lz.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"DIST_OK {world}" in r.stdout


def test_cpp_driver_threads_per_gpu(lz, tmp_path):
    """lib/final --gpus 2: one host thread per GPU in ONE process (peer access instead of CUDA IPC), checked against the oracle."""
    import numpy as np
    import oracle as orc
    if lz.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    libdir = os.path.join(ROOT, "msc-hpc-final-project_b200", "lib")
    subprocess.check_call(["make", "-s", "-C", libdir])
    n, ro, ci = lz.generate_host(lz.GraphSpec.rmat(15, 8, 1))
    ref, _, _ = orc.expv(ro, ci, 25, np.ones(n))
    ans = str(tmp_path / "ref.f64")
    ref.tofile(ans)
    for extra in ([], ["--reorth"]):      # --reorth: the coefficient reductions go through NCCL from two host threads of one process
        r = subprocess.run([os.path.join(libdir, "final"), "--graph", "rmat", "--scale", "15", "--ef", "8", "--seed", "1", "-k", "25",
                            "--gpus", "2", "--check", ans] + extra, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
