"""Problem description consumed by the hot path -- the data the Fortran host holds after
``vxc2data`` (reference src/m_vxc2data.f90:17-54) -- plus the synthetic ``plate(N)`` generator used
by the benchmark.

All index values are 1-based exactly as in the reference arrays.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

MU0_LITERAL = 0.12566370964050292e-5  # EC3D.f90:254 / vxc2data.f90:402 (not 4*pi*1e-7)


@dataclass
class Source:
    """One source function bound to one material (tFun + tfun_nod, m_vxc2data.f90:9-31)."""
    name: str
    ex: str                      # 'X', 'Y' or 'Z' ('D' for SRCZ as in vxc2data.f90:488 -> STOP)
    nomsch: int                  # material id
    nods: np.ndarray             # int32 global unknown indices (nods_Fx / nods_Fy / nods_Fz)
    move: np.ndarray = field(default_factory=lambda: np.zeros(3, np.int32))
    num_Vmech: np.ndarray = field(default_factory=lambda: np.zeros(3, np.int32))
    vel_Vmech: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float64))


@dataclass
class Problem:
    sdx: int
    sdy: int
    sdz: int
    delta: np.ndarray            # float64[3]
    dt: float
    Time: float
    BND: np.ndarray              # float64[3,2]  BND[axis, side]
    tolerance: float
    itmax: int
    geoPHYS: np.ndarray          # int8[nC], x fastest
    geoPHYS_C: np.ndarray        # int32[nC]
    valPHYS: np.ndarray          # float64[nmat,5]
    cond_numdom: List[int]       # PHYS_C(:)%numdom
    cond_nod: List[np.ndarray]   # PHYS_C(:)%nod  (int32 cell numbers)
    cond_valdom: np.ndarray      # PHYS_C(:)%valdom = 2*C/dt
    sources: List[Source]
    numMech: int
    # t -> (fun_vely[numfun] WITHOUT the mu0 factor, vmech_vely[numMech])
    evaluate_functions: Callable[[float], tuple]
    name: str = "problem"

    @property
    def nCells(self) -> int:
        return self.sdx * self.sdy * self.sdz

    @property
    def nCells0(self) -> int:
        return int(sum(len(x) for x in self.cond_nod))

    @property
    def nCellsGlob(self) -> int:
        return 3 * self.nCells + self.nCells0

    @property
    def numfun(self) -> int:
        return len(self.sources)

    def source_scalars(self, t: float):
        """EC3D.f90:245-271: (Fun(:)%vely incl. the mu0 literal, Vmech(:)%vely) at time t."""
        f, v = self.evaluate_functions(t)
        f = np.asarray(f, np.float64) * MU0_LITERAL
        return f, np.asarray(v, np.float64)

    def n_steps(self) -> int:
        """EC3D.f90:452-455: T=T+DT; IF (T < Time) GOTO 2000 -- count of executed steps."""
        T, n = 0.0, 0
        while True:
            n += 1
            T = T + self.dt
            if not (T < self.Time):
                return n

    def flat_sources(self):
        """Flattened views for the C ABIs: ex bytes, nod_ptr, nods, num_Vmech, move, vel_Vmech."""
        nf = self.numfun
        ex = "".join(s.ex[0] for s in self.sources).encode()
        ptr = np.zeros(nf + 1, np.int32)
        for i, s in enumerate(self.sources):
            ptr[i + 1] = ptr[i] + len(s.nods)
        nods = (np.concatenate([s.nods for s in self.sources]).astype(np.int32)
                if nf else np.zeros(0, np.int32))
        numv = (np.stack([s.num_Vmech for s in self.sources]).astype(np.int32)
                if nf else np.zeros((0, 3), np.int32))
        move = (np.stack([s.move for s in self.sources]).astype(np.int32)
                if nf else np.zeros((0, 3), np.int32))
        vel = (np.stack([s.vel_Vmech for s in self.sources]).astype(np.float64)
               if nf else np.zeros((0, 3), np.float64))
        return ex, ptr, np.ascontiguousarray(nods), np.ascontiguousarray(numv), \
            np.ascontiguousarray(move), np.ascontiguousarray(vel)


def renumber_air(v: np.ndarray, nsub: int):
    """vxc2data.f90:320-338: air cells (0) get environment ids nsub+1, nsub+2, ..., a new id every
    500 000 air cells.  Returns (v, nsub_air)."""
    v = v.astype(np.int64).copy()
    air = np.flatnonzero(v == 0)
    ids = nsub + 1 + (np.arange(1, air.size + 1) // 500000)
    # the reference bumps k when the running count j hits 500000, i.e. AFTER assigning... it
    # assigns v(i)=nsub+k after the bump, so the 500000-th air cell already gets the new id.
    v[air] = ids
    j = air.size % 500000
    k = 1 + air.size // 500000
    if j == 0:
        k -= 1
    return v, k


def number_conductor(v: np.ndarray, cond_numdom: Sequence[int], nC: int):
    """vxc2data.f90:620-652: geoPHYS_C = 3*nC + m (m running over domains, then k,j,i) and the
    per-domain cell lists PHYS_C(:)%nod."""
    geoC = np.zeros(nC, np.int32)
    nod = []
    m = 0
    for np_ in cond_numdom:
        cells = np.flatnonzero(v == np_)
        geoC[cells] = 3 * nC + m + 1 + np.arange(cells.size, dtype=np.int64)
        m += cells.size
        nod.append((cells + 1).astype(np.int32))
    return geoC, nod


def source_nodes(v: np.ndarray, nomsch: int, ex: str, nC: int) -> np.ndarray:
    """vxc2data.f90:656-752: the node list is built by pushing cells in k,j,i order onto a linked
    list and popping it, i.e. DESCENDING cell order; Y/Z lists carry the +nC / +2nC offset."""
    cells = np.flatnonzero(v == nomsch)[::-1] + 1
    off = {"X": 0, "Y": nC, "Z": 2 * nC}.get(ex, 0)
    return (cells + off).astype(np.int32)


def plate(N: int, variant: str = "A", tol: float = 5e-3, itmax: int = 10000,
          steps_time: Optional[float] = None) -> Problem:
    """Synthetic conductor-in-air grid ``plate(N)``.

    N x N x N cells, a conducting plate with a through-hole under a square four-bar coil.
    variant 'A': Ve = 0, static coil.  'B': adds VEX = C*15 (convection terms, EC3D.f90:657-662).
    'M': static conductor, coil moving with constant Vsx = one cell per step.
    Deterministic -- no RNG.
    """
    if N % 16 != 0:
        raise ValueError("plate(N) needs N divisible by 16")
    d = 0.00333
    delta = np.array([d, d, d], np.float64)
    dt = 1e-3
    nC = N ** 3
    v = np.zeros((N, N, N), np.int64)  # [k, j, i] zero-based
    lo, hi = N // 8, 7 * N // 8            # i,j in [N/8+1, 7N/8] (1-based)
    v[N // 8:3 * N // 8, lo:hi, lo:hi] = 1
    v[N // 8:3 * N // 8, N // 2:3 * N // 4, N // 4:N // 2] = 0   # hole i in [N/4+1,N/2], j in [N/2+1,3N/4]
    w = N // 16
    a, b = N // 4 + 1, 3 * N // 4
    k0, k1 = 5 * N // 8, 5 * N // 8 + w       # k in [5N/8+1, 5N/8+w]
    def box(i0, i1, j0, j1, mat):             # 1-based inclusive ranges
        v[k0:k1, j0 - 1:j1, i0 - 1:i1] = mat
    box(a, b - w, a, a + w - 1, 2)            # mat 2: SRCx = +Fp
    box(b - w + 1, b, a, b - w, 4)            # mat 4: SRCy = +Fp
    box(a + w, b, b - w + 1, b, 3)            # mat 3: SRCx = Fm
    box(a, a + w - 1, a + w, b, 5)            # mat 5: SRCy = Fm
    v = v.reshape(-1)
    nsub = 5
    vv, nsub_air = renumber_air(v, nsub)
    if nsub + nsub_air > 127:
        # INTEGER(1) material map of the reference overflows here (m_vxc2data.f90:43); the ids of
        # air cells are never used by the hot path, so clamp them.
        vv = np.minimum(vv, 127)
        nmat = 127
    else:
        nmat = nsub + nsub_air
    valPHYS = np.zeros((nmat, 5), np.float64)
    valPHYS[:, 0] = 1.0
    C = MU0_LITERAL * 35.26e6
    valPHYS[0, 1] = C
    if variant == "B":
        valPHYS[0, 2] = C * 15
    geoC, nod = number_conductor(vv, [1], nC)
    valdom = np.array([2.0 * C / dt], np.float64)
    srcs = []
    for mat, ex, nm in ((2, "X", "FP"), (3, "X", "FM"), (4, "Y", "FP"), (5, "Y", "FM")):
        s = Source(name=nm, ex=ex, nomsch=mat, nods=source_nodes(vv, mat, ex, nC))
        if variant == "M":
            s.move = np.array([1, 0, 0], np.int32)
            s.vel_Vmech = np.array([d / dt, 0.0, 0.0], np.float64)
        srcs.append(s)
    a0 = 183.0 / (((w * d) * w) * d)
    p2, f = 2.0 * 3.1415926535897932384626433832795, 50.0

    def evaluate_functions(t: float):
        fp = a0 * math.cos((p2 * f) * t)
        fm = (-a0) * math.cos((p2 * f) * t)
        return np.array([fp, fm, fp, fm]), np.zeros(0)

    BND = np.full((3, 2), -0.95, np.float64)
    return Problem(sdx=N, sdy=N, sdz=N, delta=delta, dt=dt,
                   Time=steps_time if steps_time is not None else 11 * dt - 0.5 * dt,
                   BND=BND, tolerance=tol, itmax=itmax,
                   geoPHYS=vv.astype(np.int8), geoPHYS_C=geoC, valPHYS=valPHYS,
                   cond_numdom=[1], cond_nod=nod, cond_valdom=valdom, sources=srcs, numMech=0,
                   evaluate_functions=evaluate_functions, name=f"plate({N}){variant}")


# ---------------------------------------------------------------------------------------------
# compact on-disk form of a Problem (tests/golden/*.npz): the voxel map, parameters and the
# per-step source scalars tabulated on the reference's time grid (T = T + DT)
# ---------------------------------------------------------------------------------------------
def save_problem_npz(p: Problem, path: str, nsteps: Optional[int] = None) -> None:
    ns = p.n_steps() if nsteps is None else nsteps
    T, times = 0.0, []
    for _ in range(ns):
        times.append(T)
        T = T + p.dt
    tabs = [p.evaluate_functions(t) for t in times]
    np.savez_compressed(
        path, dims=np.array([p.sdx, p.sdy, p.sdz], np.int32), delta=p.delta, dt=p.dt, Time=p.Time, BND=p.BND,
        tolerance=p.tolerance, itmax=p.itmax, geoPHYS=p.geoPHYS, valPHYS=p.valPHYS,
        cond_numdom=np.array(p.cond_numdom, np.int32), cond_valdom=p.cond_valdom,
        src_name=np.array([s.name for s in p.sources]), src_ex=np.array([s.ex for s in p.sources]),
        src_nomsch=np.array([s.nomsch for s in p.sources], np.int32),
        src_move=np.array([s.move for s in p.sources], np.int32).reshape(-1, 3),
        src_numv=np.array([s.num_Vmech for s in p.sources], np.int32).reshape(-1, 3),
        src_velv=np.array([s.vel_Vmech for s in p.sources], np.float64).reshape(-1, 3),
        numMech=p.numMech, times=np.array(times),
        fun_table=np.array([np.asarray(a[0], np.float64) for a in tabs]).reshape(ns, -1),
        vmech_table=np.array([np.asarray(a[1], np.float64) for a in tabs]).reshape(ns, -1),
        name=p.name)


def load_problem_npz(path: str) -> Problem:
    z = np.load(path, allow_pickle=False)
    sdx, sdy, sdz = (int(x) for x in z["dims"])
    nC = sdx * sdy * sdz
    geoPHYS = z["geoPHYS"].astype(np.int8)
    v = geoPHYS.astype(np.int64)
    cond_numdom = [int(x) for x in z["cond_numdom"]]
    geoC, nod = number_conductor(v, cond_numdom, nC)
    sources = []
    for i in range(len(z["src_ex"])):
        ex, nomsch = str(z["src_ex"][i]), int(z["src_nomsch"][i])
        sources.append(Source(name=str(z["src_name"][i]), ex=ex, nomsch=nomsch, nods=source_nodes(v, nomsch, ex, nC),
                              move=z["src_move"][i].astype(np.int32), num_Vmech=z["src_numv"][i].astype(np.int32),
                              vel_Vmech=z["src_velv"][i].astype(np.float64)))
    times, ft, vt = z["times"], z["fun_table"], z["vmech_table"]
    dt = float(z["dt"])

    def evaluate_functions(t: float):
        s = int(round(t / dt))
        if s >= len(times) or abs(times[s] - t) > 1e-9 * max(1.0, abs(t)):
            raise ValueError(f"time {t} is not on the tabulated grid of {path}")
        return ft[s].copy(), vt[s].copy()

    return Problem(sdx=sdx, sdy=sdy, sdz=sdz, delta=z["delta"].astype(np.float64), dt=dt, Time=float(z["Time"]),
                   BND=z["BND"].astype(np.float64), tolerance=float(z["tolerance"]), itmax=int(z["itmax"]),
                   geoPHYS=geoPHYS, geoPHYS_C=geoC, valPHYS=z["valPHYS"].astype(np.float64),
                   cond_numdom=cond_numdom, cond_nod=nod, cond_valdom=z["cond_valdom"].astype(np.float64),
                   sources=sources, numMech=int(z["numMech"]), evaluate_functions=evaluate_functions,
                   name=str(z["name"]))
