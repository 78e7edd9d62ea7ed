"""N>1 GPU path on real GPUs (skipped on boxes with a single GPU): torchrun launches tests/dist_check.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_gpu_slab_assembly_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "dist_check.py"), "12"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def _torchrun(script, args, port, env_extra=None, nproc=2, timeout=900):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", script)] + [str(a) for a in args]
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


ONE_GPU = {"ISL_DIST_BACKEND": "gloo", "ISL_DIST_SAME_GPU": "1"}


@pytest.mark.parametrize("mode", ["", "matrix_only"])
def test_two_ranks_on_one_gpu_slab_assembly(mode):
    """the N > 1 path on a box with ONE GPU: two processes share the device, every one with its own engine, element
    block, owned rows and pattern-only halo elements; gloo carries the ghost rows through the host (NCCL refuses two
    ranks per device).  matrix_only: the deferred Q1 stiffness launch is the last call before the exchange."""
    out = _torchrun("dist_check.py", [10] + ([mode] if mode else []), 29541, ONE_GPU)
    assert out.returncode == 0 and "DIST_CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.parametrize("name,n", [("stokes_p2p1_tet", 3), ("laplace_q1_hex", 8), ("stokes_slab", 4), ("stvenant_q1_hex", 4)])
def test_two_ranks_on_one_gpu_general_partition(name, n):
    out = _torchrun("dist_check_general.py", [name, n], 29542, ONE_GPU)
    assert out.returncode == 0 and "DIST_CHECK_GENERAL OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.parametrize("name,n", [("stokes_p2p1_tet", 4), ("laplace_q1_hex", 10), ("stokes_slab", 5)])
def test_two_gpu_general_partition_matches_oracle(name, n):
    """general element-block partition (Morton blocks of a permuted mesh, several fields, all-to-all ghost exchange)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29534", os.path.join(root, "tests", "dist_check_general.py"), name, str(n)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "DIST_CHECK_GENERAL OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
