"""ctypes binding of libec3d_gpu.so (include/ec3d_gpu.h).

The library is the product path; there is no CPU fallback.  Loading fails loudly when the shared
object has not been built (``python -c "import __graft_entry__ as g; g.build()"``), and every
compute call raises :class:`Ec3dError` when CUDA is unavailable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libec3d_gpu.so")

EXPORTS = [
    "sprsbcgstabwr_", "SPRSBCGSTABWR", "ec3d_bicgstabwr_csr", "ec3d_csr_cache_clear", "ec3d_nccl_unique_id",
    "ec3d_create", "ec3d_destroy", "ec3d_set_preconditioner", "ec3d_sizes", "ec3d_assemble_csr", "ec3d_step", "ec3d_step_stage",
    "ec3d_get_fields", "ec3d_set_fields", "ec3d_get_vtk_fields", "ec3d_get_source_cells", "ec3d_apply_operator",
    "ec3d_solve_host", "ec3d_bench_kernel", "ec3d_counters", "ec3d_timer_start", "ec3d_timer_stop", "ec3d_global_launch_count",
    "ec3d_last_error", "ec3d_version", "ec3d_partition_planes", "ec3d_plan_spmv_items",
]


# ec3d_bench_kernel selectors of the kernels of one BiCGSTABwr iteration, in launch order:
# (which, kernel, algorithmic bytes per owned unknown, per owned cell, what it replaces in solvers.f90)
KERNELS = [
    (0, "k_spmv_tma<AP>", 24.0, 5.0, "AP = A*P fused with (AP,R0), solvers.f90:30-32"),
    (1, "k_spmv_tma<SAS>", 32.0, 5.0, "S = R - alpha*AP fused into AS = A*S with ||S||^2, (AS,S), (AS,AS), solvers.f90:33-40"),
    (3, "k_xr_update_tma", 56.0, 0.0, "X += alpha*P + omega*S; R = S - omega*AS; ||R||^2; (R,R0), solvers.f90:41-44"),
    (4, "k_p_update_tma", 32.0, 0.0, "exit tests, beta, P = R + beta*(P - omega*AP), restart, solvers.f90:43-49"),
]


class Ec3dError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ec3d_gpu error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("sdx", C.c_int32), ("sdy", C.c_int32), ("sdz", C.c_int32),
        ("delta", C.c_double * 3), ("dt", C.c_double), ("BND", (C.c_double * 2) * 3),
        ("tolerance", C.c_double), ("itmax", C.c_int32),
        ("nmat", C.c_int32), ("valPHYS", C.c_void_p), ("geoPHYS", C.c_void_p), ("geoPHYS_C", C.c_void_p),
        ("size_PHYS_C", C.c_int32), ("cond_nod_ptr", C.c_void_p), ("cond_nod", C.c_void_p),
        ("cond_valdom", C.c_void_p),
        ("numfun", C.c_int32), ("fun_ex", C.c_char_p), ("fun_nod_ptr", C.c_void_p), ("fun_nods", C.c_void_p),
        ("fun_num_Vmech", C.c_void_p), ("fun_move", C.c_void_p), ("fun_vel_Vmech", C.c_void_p),
        ("numMech", C.c_int32),
        ("nranks", C.c_int32), ("rank", C.c_int32), ("nccl_id", C.c_void_p), ("device", C.c_int32),
    ]


_lib = None


def load() -> C.CDLL:
    """Load libec3d_gpu.so; raises if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run __graft_entry__.build() "
                          "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback")
    # this host raises Ec3dError on failures; the Fortran symbol aborts instead (no status argument exists)
    os.environ.setdefault("EC3D_NO_ABORT", "1")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.sprsbcgstabwr_.restype = None
    L.sprsbcgstabwr_.argtypes = [vp, vp, vp, C.POINTER(i32), vp, vp, C.POINTER(dbl), C.POINTER(i32), C.POINTER(i32)]
    L.ec3d_bicgstabwr_csr.restype = C.c_int
    L.ec3d_bicgstabwr_csr.argtypes = [vp, vp, vp, i32, vp, vp, dbl, i32, C.POINTER(i32)]
    L.ec3d_csr_cache_clear.restype = None
    L.ec3d_nccl_unique_id.restype = C.c_int
    L.ec3d_nccl_unique_id.argtypes = [vp]
    L.ec3d_create.restype = C.c_int
    L.ec3d_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.ec3d_destroy.restype = C.c_int
    L.ec3d_destroy.argtypes = [vp]
    L.ec3d_set_preconditioner.restype = C.c_int
    L.ec3d_set_preconditioner.argtypes = [vp, i32]
    L.ec3d_sizes.restype = C.c_int
    L.ec3d_sizes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
    L.ec3d_assemble_csr.restype = C.c_int
    L.ec3d_assemble_csr.argtypes = [vp, C.POINTER(i64), C.POINTER(i32)] + [vp] * 9
    L.ec3d_step.restype = C.c_int
    L.ec3d_step.argtypes = [vp, vp, vp, C.POINTER(i32)]
    L.ec3d_step_stage.restype = C.c_int
    L.ec3d_step_stage.argtypes = [vp, i32, vp, vp, C.POINTER(i32)]
    L.ec3d_get_fields.restype = C.c_int
    L.ec3d_get_fields.argtypes = [vp, vp, vp]
    L.ec3d_set_fields.restype = C.c_int
    L.ec3d_set_fields.argtypes = [vp, vp, vp]
    L.ec3d_get_vtk_fields.restype = C.c_int
    L.ec3d_get_vtk_fields.argtypes = [vp, vp, vp, vp, vp, i32]
    L.ec3d_get_source_cells.restype = C.c_int
    L.ec3d_get_source_cells.argtypes = [vp, vp]
    L.ec3d_apply_operator.restype = C.c_int
    L.ec3d_apply_operator.argtypes = [vp, vp, vp]
    L.ec3d_solve_host.restype = C.c_int
    L.ec3d_solve_host.argtypes = [vp, vp, vp, C.POINTER(i32)]
    L.ec3d_bench_kernel.restype = C.c_int
    L.ec3d_bench_kernel.argtypes = [vp, i32, i32, i32, C.POINTER(dbl)]
    L.ec3d_counters.restype = C.c_int
    L.ec3d_counters.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(dbl), C.POINTER(dbl)]
    L.ec3d_timer_start.restype = C.c_int
    L.ec3d_timer_start.argtypes = [vp]
    L.ec3d_timer_stop.restype = C.c_int
    L.ec3d_timer_stop.argtypes = [vp, C.POINTER(dbl)]
    L.ec3d_global_launch_count.restype = i64
    L.ec3d_last_error.restype = C.c_char_p
    L.ec3d_version.restype = C.c_char_p
    L.ec3d_partition_planes.restype = C.c_int
    L.ec3d_partition_planes.argtypes = [i32, i32, i32, vp, i32, vp]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise Ec3dError(rc, load().ec3d_last_error().decode(errors="replace"))


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def partition_planes(sdx: int, sdy: int, sdz: int, cond_per_plane: np.ndarray, nranks: int) -> np.ndarray:
    """Host-only: weighted z-slab partition (first plane of each rank, 0-based, plus sdz)."""
    cpp = np.ascontiguousarray(cond_per_plane, np.int64)
    ks = np.zeros(nranks + 1, np.int32)
    _check(load().ec3d_partition_planes(sdx, sdy, sdz, _p(cpp), nranks, _p(ks)))
    return ks


def plan_spmv_items(sdx: int, sdy: int, k0: int, k1: int, box, zc: int = 0, plane_major: bool = True) -> np.ndarray:
    """Host-only: work list of the TMA SpMV, rows {x0, y0, kb, ke, has_u} in launch order."""
    L = load()
    L.ec3d_plan_spmv_items.restype = C.c_int
    L.ec3d_plan_spmv_items.argtypes = [C.c_int32] * 4 + [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    b = np.ascontiguousarray(box, np.int32)
    n = C.c_int32(0)
    _check(L.ec3d_plan_spmv_items(sdx, sdy, k0, k1, _p(b), zc, 1 if plane_major else 0, None, 0, C.byref(n)))
    out = np.zeros((max(n.value, 1), 5), np.int32)
    _check(L.ec3d_plan_spmv_items(sdx, sdy, k0, k1, _p(b), zc, 1 if plane_major else 0, _p(out), n.value, C.byref(n)))
    return out[:n.value]


def nccl_unique_id() -> bytes:
    buf = (C.c_char * 128)()
    _check(load().ec3d_nccl_unique_id(C.addressof(buf)))
    return bytes(buf)


def sprsBCGstabWR(valA: np.ndarray, irow: np.ndarray, jcol: np.ndarray, n: int, b: np.ndarray, x: np.ndarray,
                  tolerance: float, itmax: int) -> int:
    """The reference's solver entry point (solvers.f90:3), same argument order and meaning, called
    through the gfortran-mangled C symbol ``sprsbcgstabwr_`` exactly as EC3D.f90:408 would.
    ``x`` is updated in place; returns ``iter``."""
    L = load()
    for a, dt in ((valA, np.float64), (irow, np.int32), (jcol, np.int32), (b, np.float64), (x, np.float64)):
        if a.dtype != dt or not a.flags.c_contiguous:
            raise TypeError("arrays must be contiguous float64 / int32")
    nn, tol, im, it = C.c_int32(n), C.c_double(tolerance), C.c_int32(itmax), C.c_int32(0)
    L.sprsbcgstabwr_(_p(valA), _p(irow), _p(jcol), C.byref(nn), _p(b), _p(x), C.byref(tol), C.byref(im), C.byref(it))
    if it.value < 0:
        raise Ec3dError(1, L.ec3d_last_error().decode(errors="replace"))
    return it.value


class Handle:
    """GPU-resident EC3D state (ec3d_create ... ec3d_destroy)."""

    def __init__(self, problem, nranks: int = 1, rank: int = 0, nccl_id: Optional[bytes] = None, device: int = -1):
        L = load()
        p = problem
        self.problem = p
        cfg = Config()
        cfg.sdx, cfg.sdy, cfg.sdz = p.sdx, p.sdy, p.sdz
        for a in range(3):
            cfg.delta[a] = float(p.delta[a])
            for s in range(2):
                cfg.BND[a][s] = float(p.BND[a, s])
        cfg.dt, cfg.tolerance, cfg.itmax = float(p.dt), float(p.tolerance), int(p.itmax)
        keep = []
        vp_ = np.ascontiguousarray(p.valPHYS, np.float64); keep.append(vp_)
        gp = np.ascontiguousarray(p.geoPHYS, np.int8); keep.append(gp)
        gc = np.ascontiguousarray(p.geoPHYS_C, np.int32); keep.append(gc)
        cfg.nmat, cfg.valPHYS, cfg.geoPHYS, cfg.geoPHYS_C = vp_.shape[0], _p(vp_), _p(gp), _p(gc)
        cfg.size_PHYS_C = len(p.cond_numdom)
        if p.cond_numdom:
            cptr = np.zeros(len(p.cond_nod) + 1, np.int32)
            for i, a in enumerate(p.cond_nod):
                cptr[i + 1] = cptr[i] + a.size
            cnod = np.ascontiguousarray(np.concatenate(p.cond_nod), np.int32)
            cval = np.ascontiguousarray(p.cond_valdom, np.float64)
            keep += [cptr, cnod, cval]
            cfg.cond_nod_ptr, cfg.cond_nod, cfg.cond_valdom = _p(cptr), _p(cnod), _p(cval)
        ex, ptr, nods, numv, move, vel = p.flat_sources()
        keep += [ex, ptr, nods, numv, move, vel]
        cfg.numfun = p.numfun
        cfg.fun_ex = ex
        cfg.fun_nod_ptr, cfg.fun_nods, cfg.fun_num_Vmech = _p(ptr), _p(nods), _p(numv)
        cfg.fun_move, cfg.fun_vel_Vmech = _p(move), _p(vel)
        cfg.numMech = p.numMech
        cfg.nranks, cfg.rank, cfg.device = nranks, rank, device
        if nccl_id is not None:
            idbuf = C.create_string_buffer(nccl_id, 128); keep.append(idbuf)
            cfg.nccl_id = C.cast(idbuf, C.c_void_p)
        self._keep = keep
        h = C.c_void_p()
        _check(L.ec3d_create(C.byref(cfg), C.byref(h)))
        self._h = h
        nC, n0, ng, k0, k1, nown = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32(), C.c_int64()
        _check(L.ec3d_sizes(h, C.byref(nC), C.byref(n0), C.byref(ng), C.byref(k0), C.byref(k1), C.byref(nown)))
        self.nCells, self.nCells0, self.nCellsGlob = nC.value, n0.value, ng.value
        self.k0, self.k1, self.n_owned = k0.value, k1.value, nown.value
        self.T = 0.0
        self.Ntime = 0
        self.iters = []

    def close(self):
        if getattr(self, "_h", None):
            load().ec3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_preconditioner(self, kind: int):
        """0 = none (the reference's algorithm, default), 1 = Jacobi (changes iterates)."""
        _check(load().ec3d_set_preconditioner(self._h, int(kind)))

    # -- gen_sparse_matrix ---------------------------------------------------------------------
    def assemble_csr(self) -> dict:
        """GPU assembly kernel; returns the reference's CSR arrays and boundary-cell lists."""
        L = load()
        nz = (C.c_int64 * 5)()
        nb = (C.c_int32 * 6)()
        _check(L.ec3d_assemble_csr(self._h, nz, nb, *([None] * 9)))
        irow = np.empty(self.nCellsGlob + 1, np.int32)
        jcol = np.empty(nz[4], np.int32)
        valA = np.empty(nz[4], np.float64)
        lists = [np.empty(max(nb[i], 1), np.int32) for i in range(6)]
        _check(L.ec3d_assemble_csr(self._h, nz, nb, _p(irow), _p(jcol), _p(valA), *[_p(a) for a in lists]))
        names = ["cel_bndX", "cel_bndY", "cel_bndZ", "cel_bndUx", "cel_bndUy", "cel_bndUz"]
        out = {"irow": irow, "jcol": jcol, "valA": valA, "num_nz": list(nz), "num_bnd": list(nb)}
        for i, nm in enumerate(names):
            out[nm] = lists[i][:nb[i]]
        return out

    # -- time loop body ------------------------------------------------------------------------
    def step(self, fun_vely=None, vmech_vely=None) -> int:
        p = self.problem
        if fun_vely is None:
            fun_vely, vmech_vely = p.source_scalars(self.T)
        fv = np.ascontiguousarray(fun_vely, np.float64)
        vv = np.ascontiguousarray(vmech_vely, np.float64) if vmech_vely is not None and len(vmech_vely) else None
        it = C.c_int32(0)
        _check(load().ec3d_step(self._h, _p(fv) if fv.size else None, _p(vv), C.byref(it)))
        self.iters.append(it.value)
        self.Ntime += 1
        self.T = self.T + p.dt
        return it.value

    def stage(self, what: int, fun_vely=None, vmech_vely=None) -> int:
        fv = np.ascontiguousarray(fun_vely, np.float64) if fun_vely is not None else None
        vv = np.ascontiguousarray(vmech_vely, np.float64) if vmech_vely is not None and len(vmech_vely) else None
        it = C.c_int32(0)
        _check(load().ec3d_step_stage(self._h, what, _p(fv), _p(vv), C.byref(it)))
        return it.value

    def get_fields(self, want_U: bool = True, want_J: bool = True) -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
        U = np.zeros(self.nCellsGlob, np.float64) if want_U else None
        J = np.zeros(self.nCellsGlob, np.float64) if want_J else None
        _check(load().ec3d_get_fields(self._h, _p(U), _p(J)))
        return U, J

    def get_fields_into(self, U: Optional[np.ndarray], J: Optional[np.ndarray]):
        _check(load().ec3d_get_fields(self._h, _p(U), _p(J)))

    def set_fields(self, U: Optional[np.ndarray] = None, J: Optional[np.ndarray] = None):
        U = None if U is None else np.ascontiguousarray(U, np.float64)
        J = None if J is None else np.ascontiguousarray(J, np.float64)
        _check(load().ec3d_set_fields(self._h, _p(U), _p(J)))

    def vtk_fields(self, big_endian: bool = False):
        """writeVtk_field's arrays (Field_A, Vector_field_eddy, Vector_field_SOURCE, Vector_field_B),
        float32, shape (nCells, 3), computed on the device (utilites.f90:222-290)."""
        out = [np.zeros((self.nCells, 3), np.float32) for _ in range(4)]
        _check(load().ec3d_get_vtk_fields(self._h, *[_p(a) for a in out], 1 if big_endian else 0))
        return tuple(out)

    def source_cells(self) -> np.ndarray:
        n = sum(len(s.nods) for s in self.problem.sources)
        out = np.zeros(max(n, 1), np.int32)
        _check(load().ec3d_get_source_cells(self._h, _p(out)))
        return out[:n]

    def apply_operator(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x)
        _check(load().ec3d_apply_operator(self._h, _p(x), _p(y)))
        return y

    def solve(self, b: np.ndarray, x: np.ndarray) -> int:
        b = np.ascontiguousarray(b, np.float64)
        assert x.dtype == np.float64 and x.flags.c_contiguous
        it = C.c_int32(0)
        _check(load().ec3d_solve_host(self._h, _p(b), _p(x), C.byref(it)))
        return it.value

    def bench_kernel(self, which: int, warm: int = 3, reps: int = 20) -> float:
        ms = C.c_double(0.0)
        _check(load().ec3d_bench_kernel(self._h, which, warm, reps, C.byref(ms)))
        return ms.value

    def timer_start(self):
        _check(load().ec3d_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double(0.0)
        _check(load().ec3d_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def step_raw(self, fv_ptr: int, vv_ptr: int) -> int:
        """ec3d_step on raw host pointers (pinned buffers in bench.py)."""
        it = C.c_int32(0)
        _check(load().ec3d_step(self._h, fv_ptr, vv_ptr, C.byref(it)))
        return it.value

    def get_fields_raw(self, u_ptr: int, j_ptr: int):
        _check(load().ec3d_get_fields(self._h, u_ptr, j_ptr))

    def set_fields_raw(self, u_ptr: int, j_ptr: int):
        _check(load().ec3d_set_fields(self._h, u_ptr, j_ptr))

    def counters(self) -> dict:
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_double(), C.c_double()
        _check(load().ec3d_counters(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"launches": a.value, "iterations": b.value, "last_step_ms": c.value, "last_solve_ms": d.value}
