"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the z-slab decomposition with
halo exchange and scalar reductions over NVLink peer memory must reproduce the single-GPU run BIT
FOR BIT (the dot products are partition independent) and meet the north-star bars against the CPU
oracle.  Runs scripts/multi_gpu_check.py under torchrun, one rank per GPU.

Cases: weighted default partition on 2 / 4 / 8 ranks, and explicit cuts (EC3D_KSTART) that put a slab
boundary exactly on the conductor's bottom / top face (the 2-plane U halo of the one-sided z
gradients, EC3D.f90:697-706), inside the conductor, and that leave slabs without any conductor cell."""
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


# plate(32): conductor planes 4..11 (0-based); compare_to_Elmer: 2..7; LIM: 9..12
CASES = [
    (2, "plate32", "p2p", None),
    (2, "plate32", "p2p", "12"),            # cut on the conductor's top face: rank 1 has no conductor cell
    (2, "plate32", "p2p", "4"),             # cut on the bottom face: rank 0 has no conductor cell
    (2, "plate32", "nccl", "8"),            # cut inside the conductor, NCCL exchange
    (2, "elmer", "p2p", None),
    (2, "lim", "nccl", None),
    (2, "move", "p2p", "8"),
    (4, "plate32", "p2p", "4,8,12"),        # conductor-free first and last slab, both faces on cuts
    (4, "lim", "p2p", None),
    (8, "plate32", "p2p", "4,8,10,12,16,20,26"),
    (8, "elmer", "p2p", None),
]


@pytest.mark.parametrize("nranks,what,comm,kstart", CASES)
def test_ranks_equal_one_gpu_and_oracle(nranks, what, comm, kstart):
    if _ngpus() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    if nranks < int(os.environ.get("EC3D_TEST_MIN_RANKS", "0")):
        pytest.skip("EC3D_TEST_MIN_RANKS")
    only = os.environ.get("EC3D_TEST_CASES")          # e.g. "0,1,4": indices into CASES (GPU-minute budgets)
    if only and str(CASES.index((nranks, what, comm, kstart))) not in only.split(","):
        pytest.skip("EC3D_TEST_CASES")
    env = dict(os.environ, EC3D_COMM=comm)
    if kstart:
        env["EC3D_KSTART"] = kstart
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "multi_gpu_check.py"),
           what, "3"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(f"bit-identical to {nranks} GPUs: True") == 3
