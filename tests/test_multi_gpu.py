"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the z-slab decomposition with
halo exchange and scalar reductions over NVLink peer memory must reproduce the single-GPU run BIT
FOR BIT (the dot products are partition independent) and meet the north-star bars against the CPU
oracle.  Runs scripts/multi_gpu_check.py under torchrun, one rank per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("what,comm", [("plate32", "p2p"), ("elmer", "p2p"), ("lim", "nccl")])
def test_two_ranks_equal_one_gpu_and_oracle(what, comm):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, EC3D_COMM=comm)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "multi_gpu_check.py"),
           what, "3"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("bit-identical to 2 GPUs: True") == 3
