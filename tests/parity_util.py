"""Shared machinery of the parity-at-scale tests (tests/test_gpu_parity_scale.py) and of
scripts/parity_table.py: one timestep sequence run by four arms on identical inputs

    gpu      the CUDA path (libec3d_gpu.so through the C ABI)
    oracle   the reference's algorithm as written (sequential DOT_PRODUCT / NORM2), oracle/ec3d_oracle.c
    pairwise the same oracle with pairwise instead of sequential reductions (orc_set_dot_mode(1))
    bridge   the same oracle with EXACTLY ROUNDED inner products (orc_sprsBCGstabWR_exact_dots)

What the arms establish:
  * gpu == bridge bit for bit: the CUDA path IS the reference's algorithm; the only thing it does
    differently is that its inner products carry no summation error;
  * oracle vs pairwise / bridge: how far the reference moves away from ITSELF when nothing but the
    rounding of its reductions changes (the unpreconditioned BiCGSTABwr at tol = 5e-3 amplifies
    1e-16 perturbations of alpha / omega by many orders of magnitude, and the warm start plus the
    Jaf history carry the difference into the next step);
  * gpu vs oracle: the north-star quantities (relative L2 of the fields, iteration counts), to be
    read next to the previous line.

`reseed=True` copies the oracle's Uaf / Jaf into every other arm before each step, so every row is a
single solve from an identical start (no accumulated drift).
"""
from __future__ import annotations

import numpy as np


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run_arms(lib, oracle_mod, p, nsteps, reseed=False, with_pairwise=True, with_bridge=True, device=0, log=print):
    h = lib.Handle(p, device=device)
    ref = oracle_mod.OracleRun(p)
    alt = oracle_mod.OracleRun(p, ref.A) if with_pairwise else None
    brg = oracle_mod.OracleRun(p, ref.A, exact_dots=True) if with_bridge else None
    rows = []
    for s in range(nsteps):
        if reseed and s > 0:
            h.set_fields(ref.Uaf, ref.Jaf)
            for o in (alt, brg):
                if o is not None:
                    o.Uaf[:], o.Jaf[:] = ref.Uaf, ref.Jaf
        f, v = p.source_scalars(ref.T)
        it_g = h.step(f, v)
        U, J = h.get_fields()
        it_o = ref.step(f, v)
        row = {"step": s, "it_gpu": it_g, "it_oracle": it_o, "relU_gpu": rel(U, ref.Uaf), "relJ_gpu": rel(J, ref.Jaf)}
        if alt is not None:
            oracle_mod.set_dot_mode(1)
            try:
                row["it_pairwise"] = alt.step(f, v)
            finally:
                oracle_mod.set_dot_mode(0)
            row["relU_pairwise"] = rel(alt.Uaf, ref.Uaf)
            row["relJ_pairwise"] = rel(alt.Jaf, ref.Jaf)
        if brg is not None:
            row["it_bridge"] = brg.step(f, v)
            row["relU_bridge"] = rel(brg.Uaf, ref.Uaf)
            row["gpu_equals_bridge"] = bool(it_g == row["it_bridge"] and np.array_equal(U, brg.Uaf) and np.array_equal(J, brg.Jaf))
            row["relU_gpu_vs_bridge"] = rel(U, brg.Uaf)
        rows.append(row)
        if log:
            log(f"{p.name}{' reseeded' if reseed else ''} step {s}: " + ", ".join(
                f"{k}={v:.2e}" if isinstance(v, float) else f"{k}={v}" for k, v in row.items() if k != "step"))
    h.close()
    return rows


def iters_within(a, b, tol=0.05):
    return abs(a - b) <= max(1, int(np.ceil(tol * b)))
