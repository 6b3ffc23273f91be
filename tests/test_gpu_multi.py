"""Multi-GPU parity of the cluster-sharded path, as -m gpu tests: two ranks are spawned with torchrun when
at least two GPUs are visible (skipped otherwise); tests/gpu_multi.py compares the sharded solve with the single-GPU solve
to 1e-25 in the objectives and +-1 in the iteration count."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("nranks,big", [(2, None), (2, 20), (4, 20)])
def test_sharded_solve_matches_single_gpu(nranks, big):
    """big = CLRS_BIG_CLUSTER: clusters with at least that many constraints take the column-distributed path (factor broadcast,
    L^-1 B solved by column chunks on all ranks, all-gather, Q slab per rank); the default threshold (512) is above these test SDPs."""
    if _ngpus() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "gpu_multi.py")]
    env = dict(os.environ)
    if big is not None:
        env["CLRS_BIG_CLUSTER"] = str(big)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=280, cwd=ROOT, env=env)
    assert p.returncode == 0 and "MULTI-GPU PARITY PASSED" in p.stdout, (p.stdout[-3000:], p.stderr[-3000:])
