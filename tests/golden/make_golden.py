"""Generates tests/golden/*.npz and *_oracle.json from the reference's shipped decks.

Run in the build container only (it reads /root/reference/src/*.vxc, which does not exist on the
GPU box):   python tests/golden/make_golden.py

  <deck>.npz          the Problem the reference's vxc2data would hand to the hot path (voxel map,
                      parameters, per-step source scalars on the reference's time grid)
  <deck>_oracle.json  what the CPU oracle computes for it: the structural counts the reference
                      prints (EC3D.f90:113,968-971,993), iteration counts and field norms of the
                      first steps, field samples at fixed indices.
The structural counts are ALSO hard-coded in tests/test_oracle.py from an independent derivation
(SURVEY.md section 8); the json pins the oracle against regressions.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from eddy_currents_3d_b200 import load_vxc  # noqa: E402
from eddy_currents_3d_b200.problem import load_problem_npz, save_problem_npz  # noqa: E402
from oracle import oracle  # noqa: E402

REF = "/root/reference/src"
DECKS = {"compare_to_Elmer": 3, "ec_src_move_hole": 3, "LIM": 4}


def main():
    for deck, nsteps in DECKS.items():
        p = load_vxc(f"{REF}/{deck}.vxc")
        npz = os.path.join(HERE, f"{deck}.npz")
        save_problem_npz(p, npz)
        q = load_problem_npz(npz)
        assert np.array_equal(p.geoPHYS_C, q.geoPHYS_C) and all(
            np.array_equal(a.nods, b.nods) for a, b in zip(p.sources, q.sources))
        run = oracle.OracleRun(q)
        A = run.A
        idx = np.linspace(0, q.nCellsGlob - 1, 64).astype(np.int64)
        out = {
            "deck": deck, "grid": [q.sdx, q.sdy, q.sdz], "nCells": q.nCells, "nCells0": q.nCells0,
            "nCellsGlob": q.nCellsGlob, "n_steps": q.n_steps(),
            "num_nz": [A.num_nzX, A.num_nzY, A.num_nzZ, A.num_nzU, A.num_nz],
            "num_bnd": [len(A.cel_bndX), len(A.cel_bndY), len(A.cel_bndZ), len(A.cel_bndUx), len(A.cel_bndUy),
                        len(A.cel_bndUz)],
            "irow_sum": int(A.irow.astype(np.int64).sum()), "jcol_sum": int(A.jcol.astype(np.int64).sum()),
            "valA_sum": float(A.valA.sum()), "valA_abs_sum": float(np.abs(A.valA).sum()),
            "sample_idx": idx.tolist(), "steps": [],
        }
        for s in range(nsteps):
            it = run.step()
            out["steps"].append({"iter": it, "Unorm": float(np.linalg.norm(run.Uaf)), "Jnorm": float(np.linalg.norm(run.Jaf)),
                                 "U_sample": run.Uaf[idx].tolist(), "J_sample": run.Jaf[idx].tolist()})
            print(deck, "step", s, "iter", it)
        with open(os.path.join(HERE, f"{deck}_oracle.json"), "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
