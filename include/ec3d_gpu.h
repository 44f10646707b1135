/*
 * ec3d_gpu.h -- C ABI of libec3d_gpu.so, the B200 (sm_100a) drop-in for the hot path of
 * JNSresearcher/eddy_currents_3d (EC3D).  Plain pointers and sizes only; every array argument is a
 * HOST pointer owned by the caller unless stated otherwise.  All index VALUES are 1-based exactly
 * as in the Fortran arrays they mirror.  There is no CPU fallback: every compute entry point
 * returns EC3D_ERR_CUDA when no sm_100-class device / driver is usable.
 *
 * Reference interfaces replaced (paths relative to the reference's src/):
 *   sprsbcgstabwr_ / ec3d_bicgstabwr_csr .... SUBROUTINE sprsBCGstabWR, solvers.f90:3-63,
 *                                             called from EC3D.f90:408
 *   ec3d_create ............................. state built by EC3D.f90:93-106,137-202 from the
 *                                             m_vxc2data module arrays (m_vxc2data.f90:17-54)
 *   ec3d_assemble_csr ....................... SUBROUTINE gen_sparse_matrix, EC3D.f90:465-1049
 *   ec3d_step ............................... loop body EC3D.f90:275-433 (source scatter with
 *                                             motion_calc/new_m :1052-1114, inertial sources,
 *                                             solve, history update)
 *   ec3d_get_fields / ec3d_set_fields ....... the Uaf / Jaf arrays, EC3D.f90:55,148
 *   ec3d_get_vtk_fields ..................... the field arithmetic of SUBROUTINE writeVtk_field,
 *                                             utilites.f90:222-290 (output step, EC3D.f90:436-444)
 */
#ifndef EC3D_GPU_H
#define EC3D_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    EC3D_OK = 0,
    EC3D_ERR_CUDA = 1,        /* CUDA runtime / driver error (message via ec3d_last_error) */
    EC3D_ERR_ARG = 2,         /* invalid argument */
    EC3D_ERR_GEOMETRY = 3,    /* the reference would STOP: column <= 0 / duplicate U column
                                 (EC3D.f90:717-720, 924-936) or conductor on a domain face */
    EC3D_ERR_NNZ_OVERFLOW = 4,/* nnz does not fit the reference's default INTEGER */
    EC3D_ERR_UNSUPPORTED = 5, /* e.g. more than one conductor domain, SRCZ (B5) */
    EC3D_ERR_NCCL = 6         /* NCCL error, or a peer-to-peer exchange timed out waiting for a neighbour rank */
};

/* ------------------------------------------------------------------------------------------ */
/* 1. Strict drop-in for solvers.f90:3                                                          */
/* ------------------------------------------------------------------------------------------ */

/* gfortran / Linux mangling of the external procedure sprsBCGstabWR (implicit interface, F77
 * calling convention: everything by reference, bare contiguous arrays).  Same semantics as the
 * reference: x is initial guess and result, b is not modified, iter = iterations performed
 * (0 and x untouched when ||b|| == 0), up to itmax+1 iterations, on non-convergence the residual
 * norm is printed to stdout and the call returns normally.  The CSR arrays are immutable during a
 * run, so a device copy is cached and reused while (pointers, n, nnz) stay the same.
 * The device copy is keyed on a content hash of irow / jcol / valA (one host pass per call), so a
 * re-assembled matrix in the same arrays is uploaded again (values only when the sparsity is unchanged).
 * On a CUDA failure the message goes to stderr and the process ABORTS: the interface has no status
 * argument and the reference's host ignores iter (EC3D.f90:408), so returning would let it time-step on
 * a stale x.  (EC3D_NO_ABORT=1 in the environment: iter = -1 and return; the Python binding sets it and
 * raises.)  SPRSBCGSTABWR is the same procedure under ifort / Windows mangling (reference Makefile:30-47). */
void sprsbcgstabwr_(double *valA, int32_t *irow, int32_t *jcol, int32_t *n, double *b, double *x,
                    double *tolerance, int32_t *itmax, int32_t *iter);
void SPRSBCGSTABWR(double *valA, int32_t *irow, int32_t *jcol, int32_t *n, double *b, double *x,
                   double *tolerance, int32_t *itmax, int32_t *iter);

/* Same operation for ISO_C_BINDING callers; returns a status code. */
int ec3d_bicgstabwr_csr(const double *valA, const int32_t *irow, const int32_t *jcol, int32_t n,
                        const double *b, double *x, double tolerance, int32_t itmax,
                        int32_t *iter);

/* Drops the cached device copy of the CSR matrix. */
void ec3d_csr_cache_clear(void);

/* ------------------------------------------------------------------------------------------ */
/* 2. GPU-resident time stepping (EC3D.f90:275-433)                                             */
/* ------------------------------------------------------------------------------------------ */

typedef struct ec3d_handle ec3d_handle;

typedef struct {
    /* grid (vxc2data outputs, EC3D.f90:86-89) */
    int32_t sdx, sdy, sdz;
    double delta[3];
    double dt;
    double BND[3][2];            /* BND[axis][side] == Fortran BND(axis+1, side+1) */
    double tolerance;
    int32_t itmax;
    /* materials (m_vxc2data.f90:43-52) */
    int32_t nmat;
    const double *valPHYS;       /* [nmat][5], valPHYS(m,c) at [(m-1)*5 + (c-1)] */
    const int8_t *geoPHYS;       /* [sdx*sdy*sdz] x fastest */
    const int32_t *geoPHYS_C;    /* [sdx*sdy*sdz] 0 or 3*nC + m */
    /* conductor domains (tPHYS, m_vxc2data.f90:34-39); exactly 0 or 1 domain is supported */
    int32_t size_PHYS_C;
    const int32_t *cond_nod_ptr; /* [size_PHYS_C+1] */
    const int32_t *cond_nod;     /* PHYS_C(:)%nod, cell numbers */
    const double *cond_valdom;   /* PHYS_C(:)%valdom = 2*C/dt */
    /* source functions (tFun/tfun_nod, m_vxc2data.f90:9-31) */
    int32_t numfun;
    const char *fun_ex;          /* [numfun] 'X' or 'Y' */
    const int32_t *fun_nod_ptr;  /* [numfun+1] */
    const int32_t *fun_nods;     /* nods_Fx / nods_Fy (global unknown indices) */
    const int32_t *fun_num_Vmech;/* [numfun][3] */
    const int32_t *fun_move;     /* [numfun][3] */
    const double *fun_vel_Vmech; /* [numfun][3] */
    int32_t numMech;
    /* multi-GPU z-slab decomposition: this process is `rank` of `nranks` (one process per GPU);
     * nccl_id = 128 bytes from ec3d_nccl_unique_id() of rank 0 (ignored when nranks == 1).  NCCL is
     * used for set-up; halo planes and solver scalars then travel through NVLink peer memory (CUDA
     * IPC), with NCCL send/recv/all-gather as the fallback.  Results are bit-identical for any nranks. */
    int32_t nranks, rank;
    const void *nccl_id;
    int32_t device;              /* CUDA device ordinal, -1 = current */
} ec3d_config;

int ec3d_nccl_unique_id(void *id128);

/* With nranks > 1 the following calls are COLLECTIVE: every rank must make them in the same order with
 * compatible arguments, or the ranks' exchange epochs desynchronise and the waiting side gives up after
 * about 30 s with EC3D_ERR_NCCL: ec3d_create, ec3d_step / ec3d_step_stage (stages 1 and 2), ec3d_set_fields
 * (one exchange per non-NULL field), ec3d_apply_operator, ec3d_solve_host, ec3d_get_vtk_fields with
 * field_B != NULL, ec3d_set_preconditioner, ec3d_bench_kernel, ec3d_destroy.  ec3d_get_fields,
 * ec3d_get_source_cells, ec3d_counters and the timers are local. */

int ec3d_create(const ec3d_config *cfg, ec3d_handle **out);
int ec3d_destroy(ec3d_handle *h);

/* Optional preconditioning of the resident solver.  kind 0 = none (default: the
 * reference's algorithm, solvers.f90:3-50), 1 = Jacobi: BiCGSTABwr on D^-1 A x = D^-1 b with D the matrix
 * diagonal (EC3D.f90:533-663).  It changes the iterates and iteration counts, so parity with the reference
 * only holds for kind 0.  With nranks > 1 every rank must make the same call. */
int ec3d_set_preconditioner(ec3d_handle *h, int32_t kind);

/* Sizes: n = nCellsGlob (global), and this rank's slab [k0,k1) (0-based planes) and owned unknown
 * count. */
int ec3d_sizes(const ec3d_handle *h, int64_t *nCells, int64_t *nCells0, int64_t *nCellsGlob,
               int32_t *k0, int32_t *k1, int64_t *n_owned);

/* GPU assembly kernel: emits the reference's CSR (EC3D.f90:465-1049).  Call once with all
 * output pointers NULL to obtain the counts, then with caller-allocated arrays
 * (irow: nCellsGlob+1, jcol/valA: num_nz[4], lists: num_bnd[0..5] = X,Y,Z,Ux,Uy,Uz).
 * num_nz[0..3] = X,Y,Z,U block counts, num_nz[4] = total.  Single-rank handles only. */
int ec3d_assemble_csr(ec3d_handle *h, int64_t num_nz[5], int32_t num_bnd[6], int32_t *irow,
                      int32_t *jcol, double *valA, int32_t *cel_bndX, int32_t *cel_bndY,
                      int32_t *cel_bndZ, int32_t *cel_bndUx, int32_t *cel_bndUy,
                      int32_t *cel_bndUz);

/* One timestep, EC3D.f90:275-433.  fun_vely[numfun] = Fun(:)%vely (already times the mu0 literal,
 * EC3D.f90:254); vmech_vely[numMech] = Vmech(:)%vely.  iter receives the BiCGSTABwr iteration
 * count.  Uaf/Jaf and coil positions stay on the device. */
int ec3d_step(ec3d_handle *h, const double *fun_vely, const double *vmech_vely, int32_t *iter);

/* Stage-wise access for parity tests: what = 0 scatter sources (EC3D.f90:275-367),
 * 1 inertial sources / rhs (370-404), 2 solve (408), 3 history update (412-433). */
int ec3d_step_stage(ec3d_handle *h, int32_t what, const double *fun_vely,
                    const double *vmech_vely, int32_t *iter);

/* Full-length (nCellsGlob) host arrays in the reference's layout [Ax|Ay|Az|U].  With nranks > 1
 * each rank fills / reads only its owned entries (others are left untouched on get). Either
 * pointer may be NULL. */
int ec3d_get_fields(ec3d_handle *h, double *Uaf, double *Jaf);
int ec3d_set_fields(ec3d_handle *h, const double *Uaf, const double *Jaf);

/* Output post-processing on the device (writeVtk_field, utilites.f90:222-290): the per-point float32
 * triples of a field_N.vtk in file order (x fastest; 3 components per point; 3*nCells floats per
 * array): field_A = Uaf; field_eddy = s*Jaf on conductor cells, 0 elsewhere (:237-250, meaningful
 * when size_PHYS_C != 0); field_source = Jaf on non-conductor cells (:253-274); field_B = curl A by
 * central differences with clamped indices (:276-290).  Computed in fp64 in the reference's order and
 * rounded once.  big_endian != 0 stores the bytes as the reference's convert="big_endian" stream
 * expects, so the host can write the arrays straight into the file.  Any pointer may be NULL.  With
 * nranks > 1 each rank fills only the points of its z-slab. */
int ec3d_get_vtk_fields(ec3d_handle *h, float *field_A, float *field_eddy, float *field_source,
                        float *field_B, int32_t big_endian);

/* Moved source cells of the last step in (function, node) order (new_nodesX/Y, EC3D.f90:311-320):
 * cell numbers 1..nC. */
int ec3d_get_source_cells(ec3d_handle *h, int32_t *cells);

/* y = A*x with the matrix-free operator on full-length host vectors (parity tests). */
int ec3d_apply_operator(ec3d_handle *h, const double *x, double *y);

/* Solve A x = b on full-length host vectors with the matrix-free operator (x in/out). */
int ec3d_solve_host(ec3d_handle *h, const double *b, double *x, int32_t *iter);

/* ------------------------------------------------------------------------------------------ */
/* 3. Measurement hooks (bench.py)                                                              */
/* ------------------------------------------------------------------------------------------ */

/* Times `reps` launches of one kernel class with CUDA events on the library's stream after
 * `warm` untimed launches; writes the mean ms per launch.  which: 0 = stencil SpMV A*p fused with
 * (Ap,r0); 1 = stencil SpMV A*s fused with s = r - alpha*Ap, ||s||^2, (As,s), (As,As) (the unfused
 * A*s kernel when the s-update is not fused: odd grids, EC3D_FUSE_S=0); 2 = stand-alone
 * s = r - alpha*Ap with ||s||^2; 3 = x,r update with 2 dots; 4 = p update; 5 = one whole BiCGSTABwr
 * iteration, every kernel in sequence without exit (single rank only). */
int ec3d_bench_kernel(ec3d_handle *h, int32_t which, int32_t warm, int32_t reps, double *ms);

/* Counters since creation: kernels launched by this library, solver iterations, and device ms
 * (CUDA events) of the last ec3d_step: total and solve only. */
int ec3d_counters(const ec3d_handle *h, int64_t *launches, int64_t *iterations,
                  double *last_step_ms, double *last_solve_ms);

/* CUDA-event stopwatch on the library's stream (the stream every kernel of `h` is launched on):
 * start synchronises the stream and records; stop records, waits and returns the elapsed ms. */
int ec3d_timer_start(ec3d_handle *h);
int ec3d_timer_stop(ec3d_handle *h, double *ms);

int64_t ec3d_global_launch_count(void);

const char *ec3d_last_error(void);
const char *ec3d_version(void);

/* ------------------------------------------------------------------------------------------ */
/* 4. Host-only helpers (no GPU needed)                                                         */
/* ------------------------------------------------------------------------------------------ */

/* Weighted z-slab partition: planes are split so every rank moves about
 * the same HBM bytes per iteration.  cond_per_plane[sdz] = conductor cells in each plane;
 * kstart[nranks+1] receives the 0-based first plane of each rank (kstart[nranks] = sdz). */
int ec3d_partition_planes(int32_t sdx, int32_t sdy, int32_t sdz, const int64_t *cond_per_plane,
                          int32_t nranks, int32_t *kstart);

/* Work list of the TMA-staged SpMV for a slab of planes [k0,k1) of an sdx x sdy x . grid whose conductor
 * bounding box is box = {i0,i1,j0,j1,k0,k1} (half open, 0-based; k0 == k1: no conductor): items5 receives
 * {x0, y0, kb, ke, has_u} per item in launch order (pass NULL to get only the count).  zc <= 0 selects the
 * default item length.  Every (64 x 8 tile, plane) is covered by exactly one item. */
int ec3d_plan_spmv_items(int32_t sdx, int32_t sdy, int32_t k0, int32_t k1, const int32_t box[6], int32_t zc,
                         int32_t plane_major, int32_t *items5, int32_t max_items, int32_t *n_items);

#ifdef __cplusplus
}
#endif
#endif
