"""ctypes front end of the CPU oracle (oracle/ec3d_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  PARITY UNPINNED by reference tests (see ec3d_oracle.h).

``OracleRun`` plays the role of the reference's main program (src/EC3D.f90:93-455) on top of the
oracle's restated subroutines.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libec3d_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ec3d_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _Grid(C.Structure):
    _fields_ = [("sdx", C.c_int32), ("sdy", C.c_int32), ("sdz", C.c_int32),
                ("delta", C.c_double * 3), ("dt", C.c_double), ("BND", (C.c_double * 2) * 3),
                ("nmat", C.c_int32), ("valPHYS", C.c_void_p), ("geoPHYS", C.c_void_p),
                ("geoPHYS_C", C.c_void_p), ("size_PHYS_C", C.c_int32), ("nCells0", C.c_int32)]


class _Csr(C.Structure):
    _fields_ = [("irow", C.c_void_p), ("jcol", C.c_void_p), ("valA", C.c_void_p),
                ("cel_bndX", C.c_void_p), ("cel_bndY", C.c_void_p), ("cel_bndZ", C.c_void_p),
                ("cel_bndUx", C.c_void_p), ("cel_bndUy", C.c_void_p), ("cel_bndUz", C.c_void_p),
                ("num_nzX", C.c_int64), ("num_nzY", C.c_int64), ("num_nzZ", C.c_int64),
                ("num_nzU", C.c_int64), ("num_nz", C.c_int64),
                ("num_bndX", C.c_int32), ("num_bndY", C.c_int32), ("num_bndZ", C.c_int32),
                ("num_bndUx", C.c_int32), ("num_bndUy", C.c_int32), ("num_bndUz", C.c_int32),
                ("err_cell", C.c_int32), ("err_col", C.c_int32)]


class _Sources(C.Structure):
    _fields_ = [("numfun", C.c_int32), ("ex", C.c_char_p), ("nod_ptr", C.c_void_p),
                ("nods", C.c_void_p), ("num_Vmech", C.c_void_p), ("move", C.c_void_p),
                ("vel_Vmech", C.c_void_p), ("Distance", C.c_void_p), ("shift", C.c_void_p),
                ("length", C.c_void_p), ("movestop", C.c_int32 * 3), ("flag_move", C.c_int32)]


class _Cond(C.Structure):
    _fields_ = [("size_PHYS_C", C.c_int32), ("nod_ptr", C.c_void_p), ("nod", C.c_void_p),
                ("valdom", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_nnz_upper_bound.restype = C.c_int64
        L.orc_nnz_upper_bound.argtypes = [C.POINTER(_Grid)]
        L.orc_gen_sparse_matrix.restype = C.c_int
        L.orc_gen_sparse_matrix.argtypes = [C.POINTER(_Grid), C.POINTER(_Csr)]
        L.orc_sprsBCGstabWR.restype = C.c_int
        L.orc_sprsBCGstabWR.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_double, C.c_int32, C.POINTER(C.c_int32)]
        L.orc_sprsBCGstabWR_exact_dots.restype = C.c_int
        L.orc_sprsBCGstabWR_exact_dots.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                                   C.c_void_p, C.c_double, C.c_int32, C.POINTER(C.c_int32),
                                                   C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_sprsAx.restype = None
        L.orc_sprsAx.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_set_dot_mode.restype = None
        L.orc_set_dot_mode.argtypes = [C.c_int]
        L.orc_norm2.restype = C.c_double
        L.orc_norm2.argtypes = [C.c_void_p, C.c_int64]
        L.orc_dot.restype = C.c_double
        L.orc_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_motion_prepare.restype = None
        L.orc_motion_prepare.argtypes = [C.POINTER(_Sources), C.c_void_p, C.c_double]
        L.orc_scatter_sources.restype = C.c_int
        L.orc_scatter_sources.argtypes = [C.POINTER(_Grid), C.POINTER(_Sources), C.POINTER(_Cond),
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_rhs_pre.restype = None
        L.orc_rhs_pre.argtypes = [C.POINTER(_Grid), C.POINTER(_Cond), C.POINTER(_Csr), C.c_void_p, C.c_void_p]
        L.orc_rhs_post.restype = None
        L.orc_rhs_post.argtypes = [C.POINTER(_Grid), C.POINTER(_Cond), C.POINTER(_Csr), C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def set_dot_mode(mode: int) -> None:
    """0 = the reference's sequential reductions (default); 1 = pairwise (sensitivity runs only)."""
    lib().orc_set_dot_mode(int(mode))


def norm2(x: np.ndarray) -> float:
    x = np.ascontiguousarray(x, np.float64)
    return lib().orc_norm2(_p(x), x.size)


def dot(a: np.ndarray, b: np.ndarray) -> float:
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    return lib().orc_dot(_p(a), _p(b), a.size)


def spmv(valA, irow, jcol, v):
    y = np.empty(irow.size - 1, np.float64)
    lib().orc_sprsAx(_p(valA), _p(irow), _p(jcol), irow.size - 1, _p(v), _p(y))
    return y


def bicgstabwr(valA, irow, jcol, b, x, tolerance: float, itmax: int) -> int:
    """solvers.f90:3 -- x is updated in place; returns iter."""
    assert x.dtype == np.float64 and x.flags.c_contiguous
    it = C.c_int32(0)
    rc = lib().orc_sprsBCGstabWR(_p(valA), _p(irow), _p(jcol), irow.size - 1, _p(b), _p(x),
                                 float(tolerance), int(itmax), C.byref(it))
    if rc != 0:
        raise MemoryError("oracle solver allocation failed")
    return it.value


def bicgstabwr_exact_dots(valA, irow, jcol, b, x, tolerance: float, itmax: int, problem) -> int:
    """The bridge solver (see ec3d_oracle.h): solvers.f90:3-50 with exactly rounded inner products."""
    assert x.dtype == np.float64 and x.flags.c_contiguous
    g = np.ascontiguousarray(problem.geoPHYS_C, np.int32)
    it = C.c_int32(0)
    rc = lib().orc_sprsBCGstabWR_exact_dots(_p(valA), _p(irow), _p(jcol), irow.size - 1, _p(b), _p(x),
                                            float(tolerance), int(itmax), C.byref(it),
                                            problem.sdx, problem.sdy, problem.sdz, _p(g))
    if rc != 0:
        raise RuntimeError(f"bridge solver failed (rc={rc}; -2: odd sdx)")
    return it.value


def vtk_fields(problem, Uaf: np.ndarray, Jaf: np.ndarray):
    """utilites.f90:222-290: (Field_A, Vector_field_eddy, Vector_field_SOURCE, Vector_field_B) as
    float32 arrays of shape (nCells, 3) in file order."""
    p = problem
    L = lib()
    L.orc_vtk_fields.restype = None
    L.orc_vtk_fields.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    nC = p.nCells
    out = [np.zeros((nC, 3), np.float32) for _ in range(4)]
    delta = np.ascontiguousarray(p.delta, np.float64)
    U = np.ascontiguousarray(Uaf, np.float64)
    J = np.ascontiguousarray(Jaf, np.float64)
    g = np.ascontiguousarray(p.geoPHYS_C, np.int32)
    L.orc_vtk_fields(p.sdx, p.sdy, p.sdz, _p(delta), _p(U), _p(J), _p(g), len(p.cond_numdom), *[_p(a) for a in out])
    return tuple(out)


class Assembled:
    """Result of gen_sparse_matrix (EC3D.f90:465-1049)."""

    def __init__(self, problem):
        p = problem
        self.problem = p
        self._keep = []
        g = _Grid()
        g.sdx, g.sdy, g.sdz = p.sdx, p.sdy, p.sdz
        for a in range(3):
            g.delta[a] = float(p.delta[a])
            for s in range(2):
                g.BND[a][s] = float(p.BND[a, s])
        g.dt = float(p.dt)
        self._valPHYS = np.ascontiguousarray(p.valPHYS, np.float64)
        self._geoPHYS = np.ascontiguousarray(p.geoPHYS, np.int8)
        self._geoPHYS_C = np.ascontiguousarray(p.geoPHYS_C, np.int32)
        g.nmat = self._valPHYS.shape[0]
        g.valPHYS, g.geoPHYS, g.geoPHYS_C = _p(self._valPHYS), _p(self._geoPHYS), _p(self._geoPHYS_C)
        g.size_PHYS_C = len(p.cond_numdom)
        g.nCells0 = p.nCells0
        self.grid = g
        n = p.nCellsGlob
        ub = lib().orc_nnz_upper_bound(C.byref(g))
        irow = np.empty(n + 1, np.int32)
        jcol = np.empty(ub, np.int32)
        valA = np.empty(ub, np.float64)
        nb = max(p.nCells0, 1)
        lists = [np.zeros(nb, np.int32) for _ in range(6)]
        c = _Csr()
        c.irow, c.jcol, c.valA = _p(irow), _p(jcol), _p(valA)
        (c.cel_bndX, c.cel_bndY, c.cel_bndZ, c.cel_bndUx, c.cel_bndUy, c.cel_bndUz) = [_p(a) for a in lists]
        self.rc = lib().orc_gen_sparse_matrix(C.byref(g), C.byref(c))
        self.err_cell, self.err_col = c.err_cell, c.err_col
        self.num_nz = int(c.num_nz)
        self.num_nzX, self.num_nzY, self.num_nzZ, self.num_nzU = (int(c.num_nzX), int(c.num_nzY),
                                                                 int(c.num_nzZ), int(c.num_nzU))
        self.irow = irow
        self.jcol = jcol[:self.num_nz].copy() if self.rc == 0 else jcol[:0]
        self.valA = valA[:self.num_nz].copy() if self.rc == 0 else valA[:0]
        del jcol, valA
        self.cel_bndX = lists[0][:c.num_bndX].copy()
        self.cel_bndY = lists[1][:c.num_bndY].copy()
        self.cel_bndZ = lists[2][:c.num_bndZ].copy()
        self.cel_bndUx = lists[3][:c.num_bndUx].copy()
        self.cel_bndUy = lists[4][:c.num_bndUy].copy()
        self.cel_bndUz = lists[5][:c.num_bndUz].copy()
        # a struct pointing at the trimmed arrays, for rhs_pre / rhs_post
        c2 = _Csr()
        c2.irow, c2.jcol, c2.valA = _p(self.irow), _p(self.jcol), _p(self.valA)
        c2.cel_bndX, c2.cel_bndY, c2.cel_bndZ = _p(self.cel_bndX), _p(self.cel_bndY), _p(self.cel_bndZ)
        c2.cel_bndUx, c2.cel_bndUy, c2.cel_bndUz = _p(self.cel_bndUx), _p(self.cel_bndUy), _p(self.cel_bndUz)
        c2.num_bndX, c2.num_bndY, c2.num_bndZ = c.num_bndX, c.num_bndY, c.num_bndZ
        c2.num_bndUx, c2.num_bndUy, c2.num_bndUz = c.num_bndUx, c.num_bndUy, c.num_bndUz
        c2.num_nz = c.num_nz
        self.csr = c2


class OracleRun:
    """The reference's main program (EC3D.f90:93-455) over the oracle subroutines."""

    def __init__(self, problem, assembled: Optional[Assembled] = None, exact_dots: bool = False, solver=None):
        p = problem
        self.p = p
        self.exact_dots = exact_dots          # True: the bridge solver instead of the reference's reductions
        # solver(valA, irow, jcol, n, b, x, tolerance, itmax) -> iter: stands in for the CALL at EC3D.f90:408
        # (bench.py times the GPU drop-in sprsbcgstabwr_ inside the reference's own time loop this way)
        self.solver = solver
        self.A = assembled if assembled is not None else Assembled(p)
        if self.A.rc != 0:
            raise RuntimeError(f"gen_sparse_matrix: reference would STOP (rc={self.A.rc}, "
                               f"cell={self.A.err_cell}, col={self.A.err_col})")
        n = p.nCellsGlob
        self.Uaf = np.zeros(n, np.float64)                                    # EC3D.f90:148
        self.Jaf = np.zeros(n, np.float64)
        self.Jafbuf = np.zeros(n if p.cond_numdom else 1, np.float64)
        ex, ptr, nods, numv, move, vel = p.flat_sources()
        self._src_arrays = (ex, ptr, nods, numv, move, vel)
        nf = p.numfun
        self.Distance = np.zeros((max(nf, 1), 3), np.float64)
        self.shift = np.zeros((max(nf, 1), 3), np.float64)
        self.length = np.zeros((max(nf, 1), 3), np.int32)
        s = _Sources()
        s.numfun = nf
        s.ex = ex
        s.nod_ptr, s.nods, s.num_Vmech, s.move, s.vel_Vmech = _p(ptr), _p(nods), _p(numv), _p(move), _p(vel)
        s.Distance, s.shift, s.length = _p(self.Distance), _p(self.shift), _p(self.length)
        self.src = s
        self._delta = np.ascontiguousarray(p.delta, np.float64)
        lib().orc_motion_prepare(C.byref(s), _p(self._delta), float(p.dt))
        cp = np.zeros(len(p.cond_nod) + 1, np.int32)
        for i, a in enumerate(p.cond_nod):
            cp[i + 1] = cp[i] + a.size
        cn = (np.concatenate(p.cond_nod).astype(np.int32) if p.cond_nod else np.zeros(1, np.int32))
        cv = np.ascontiguousarray(p.cond_valdom, np.float64) if p.cond_numdom else np.zeros(1)
        self._cond_arrays = (cp, cn, cv)
        c = _Cond()
        c.size_PHYS_C = len(p.cond_numdom)
        c.nod_ptr, c.nod, c.valdom = _p(cp), _p(cn), _p(cv)
        self.cond = c
        self.new_nodes = np.zeros(max(nods.size, 1), np.int32)
        self.T = 0.0
        self.Ntime = 0
        self.iters = []

    @property
    def flag_move(self) -> int:
        return int(self.src.flag_move)

    def step(self, fun_vely=None, vmech_vely=None, solve: bool = True) -> int:
        """One pass of the loop body EC3D.f90:241-455 (without output)."""
        p = self.p
        if fun_vely is None:
            fun_vely, vmech_vely = p.source_scalars(self.T)
        fv = np.ascontiguousarray(fun_vely, np.float64)
        vv = np.ascontiguousarray(vmech_vely if vmech_vely is not None and len(vmech_vely) else np.zeros(1), np.float64)
        L = lib()
        rc = L.orc_scatter_sources(C.byref(self.A.grid), C.byref(self.src), C.byref(self.cond), _p(fv), _p(vv),
                                   _p(self.Jaf), _p(self.Jafbuf), _p(self.new_nodes))
        if rc != 0:
            raise RuntimeError("source scatter: reference would STOP")
        L.orc_rhs_pre(C.byref(self.A.grid), C.byref(self.cond), C.byref(self.A.csr), _p(self.Uaf), _p(self.Jaf))
        self.rhs = self.Jaf.copy()
        it = 0
        if solve and self.solver is not None:
            it = self.solver(self.A.valA, self.A.irow, self.A.jcol, p.nCellsGlob, self.Jaf, self.Uaf, p.tolerance, p.itmax)
        elif solve and self.exact_dots:
            it = bicgstabwr_exact_dots(self.A.valA, self.A.irow, self.A.jcol, self.Jaf, self.Uaf, p.tolerance, p.itmax, p)
        elif solve:
            it = bicgstabwr(self.A.valA, self.A.irow, self.A.jcol, self.Jaf, self.Uaf, p.tolerance, p.itmax)
        L.orc_rhs_post(C.byref(self.A.grid), C.byref(self.cond), C.byref(self.A.csr), _p(self.Uaf), _p(self.Jaf))
        self.iters.append(it)
        self.Ntime += 1
        self.T = self.T + p.dt
        return it
