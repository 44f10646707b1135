"""z-slab decomposition of the unknown vector -- the Python mirror of the
layout libec3d_gpu.so builds in ec3d_create (csrc/ec3d_gpu.cu, "slab partition").

Rank r owns planes [k0, k1): three contiguous A ranges (x-fastest cells of those planes) and one
contiguous U range (U unknowns are numbered in k,j,i order for a single conductor domain).  An SpMV
over the owned rows reads, besides owned entries, exactly: one plane of each A component below and
above, and the U unknowns of the two planes below and above (one-sided z-gradients reach two cells,
EC3D.f90:697-706).

Used by the tests only (tests/test_distributed_cpu.py checks the halo / ownership bookkeeping with a
world-size-2 gloo group on the CPU); the product path keeps its own copy of this layout in C++.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

from . import lib


@dataclass
class Slab:
    rank: int
    k0: int
    k1: int
    owned: List[slice]          # 4 global 0-based index ranges: Ax, Ay, Az, U
    halo_lo: List[slice]        # what is received from rank-1: Ax, Ay, Az plane k0-1; U of planes k0-2..k0-1
    halo_hi: List[slice]        # what is received from rank+1: planes k1 (A) and k1..k1+1 (U)
    send_lo: List[slice]        # what rank-1 needs from this rank
    send_hi: List[slice]

    @property
    def n_owned(self) -> int:
        return sum(s.stop - s.start for s in self.owned)


def make_slabs(problem, nranks: int) -> List[Slab]:
    p = problem
    kdz, nC, sdz = p.sdx * p.sdy, p.nCells, p.sdz
    cpp = np.zeros(sdz, np.int64)
    if p.cond_nod:
        np.add.at(cpp, (np.concatenate(p.cond_nod) - 1) // kdz, 1)
    ks = lib.partition_planes(p.sdx, p.sdy, sdz, cpp, nranks)
    ucum = np.concatenate([[0], np.cumsum(cpp)])

    def cl(k):
        return min(max(k, 0), sdz)

    def a_planes(ka, kb):
        return [slice(c * nC + ka * kdz, c * nC + kb * kdz) for c in range(3)]

    def u_planes(ka, kb):
        return slice(3 * nC + int(ucum[cl(ka)]), 3 * nC + int(ucum[cl(kb)]))

    slabs = []
    for r in range(nranks):
        k0, k1 = int(ks[r]), int(ks[r + 1])
        owned = a_planes(k0, k1) + [u_planes(k0, k1)]
        halo_lo = (a_planes(k0 - 1, k0) + [u_planes(k0 - 2, k0)]) if r > 0 else []
        halo_hi = (a_planes(k1, k1 + 1) + [u_planes(k1, k1 + 2)]) if r < nranks - 1 else []
        send_lo = (a_planes(k0, k0 + 1) + [u_planes(k0, k0 + 2)]) if r > 0 else []
        send_hi = (a_planes(k1 - 1, k1) + [u_planes(k1 - 2, k1)]) if r < nranks - 1 else []
        slabs.append(Slab(r, k0, k1, owned, halo_lo, halo_hi, send_lo, send_hi))
    return slabs
