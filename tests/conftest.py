import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DECKS = ["compare_to_Elmer", "ec_src_move_hole", "LIM"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def deck_problems():
    from eddy_currents_3d_b200.problem import load_problem_npz
    return {d: load_problem_npz(os.path.join(GOLDEN, d + ".npz")) for d in DECKS}


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library; GPU tests must not silently fall back to anything else."""
    from eddy_currents_3d_b200 import lib
    L = lib.load()
    import ctypes
    assert isinstance(L, ctypes.CDLL)
    return lib
