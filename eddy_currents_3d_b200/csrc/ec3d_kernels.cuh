// ec3d_kernels.cuh -- sm_100a kernels of the EC3D hot path: generic matrix-free SpMV (odd grids; the
// main TMA-staged SpMV is in ec3d_tma.cuh), CSR SpMV (drop-in mode), fused BiCGSTABwr vector kernels,
// double-double reductions.
//
// Arithmetic: fp64 with explicit round-to-nearest mul/add (__dmul_rn/__dadd_rn are never
// contracted into FMA), each row summed in ascending column order starting from 0 -- the order of
// the reference's sprsAx (solvers.f90:54-61) -- so the matrix-free operator is bit-identical to a
// sequential CSR row sum.  Dot products: the products are formed per element / per cell pair in a
// fixed order that does not depend on how the grid is cut into slabs, chunks or blocks, and
// everything above that level (per thread across planes, warp, block, blocks, ranks) is summed in
// double-double (error-free TwoSum), whose result rounded to fp64 is the same for any summation
// order (up to ties at the 1e-32 level).  So the solver scalars -- and with them iteration counts
// and fields -- are reproducible run to run AND identical for 1, 2, 4, 8 GPUs; they differ from the
// reference's sequential dot_product only by the reference's own rounding error.
#pragma once
#include "ec3d_async.cuh"
#include "ec3d_common.cuh"
#include "ec3d_comm.cuh"
#include "ec3d_rows.cuh"

#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))

// MODE_SAS (TMA SpMV only): the input is s = r - alpha*Ap formed on the fly from x = r and x2 = Ap
// (solvers.f90:33 fused into the A*s of :39); the kernel also stores s and reduces ||s||^2.
enum { MODE_PLAIN = 0, MODE_AP = 1, MODE_AS = 2, MODE_INIT = 3, MODE_SAS = 4 };

struct VecSet {
    const double *x;   // SpMV input (MODE_SAS: r)
    double *y;         // SpMV output (AP / AS / plain)
    const double *r0;  // MODE_AP
    const double *b;   // MODE_INIT
    double *R, *R0, *P;// MODE_INIT outputs
    const double *x2;  // MODE_SAS: Ap
    double *S;         // MODE_SAS: s output
    int vi_y, vi_R, vi_P;   // fused halo push (several ranks): index of y / R / P among the handle's vectors
};

// --------------------------------------------------------------------------------------------
// reductions (double-double)
// --------------------------------------------------------------------------------------------
struct dd {
    double hi, lo;
};
__device__ __forceinline__ dd dd_zero() { return dd{0.0, 0.0}; }

// a += b, b a double: hi + b is split exactly into the rounded sum and its error (Knuth TwoSum)
__device__ __forceinline__ void dd_add_d(dd &a, const double b)
{
    const double s = DADD(a.hi, b);
    const double bb = DSUB(s, a.hi);
    const double e = DADD(DSUB(a.hi, DSUB(s, bb)), DSUB(b, bb));
    a.hi = s;
    a.lo = DADD(a.lo, e);
}
// a += b, both double-double
__device__ __forceinline__ void dd_add_dd(dd &a, const dd b)
{
    dd_add_d(a, b.hi);
    a.lo = DADD(a.lo, b.lo);
}
__device__ __forceinline__ double dd_round(const dd a) { return DADD(a.hi, a.lo); }

__device__ __forceinline__ dd warp_sum(dd v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dd w;
        w.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
        w.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
        dd_add_dd(v, w);
    }
    return v;
}

// Sum over the block; result valid in thread 0.  blockDim.x*blockDim.y*blockDim.z <= 1024.
__device__ __forceinline__ dd block_sum(dd v, double *sh /* >= 64 doubles */)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, w = tid >> 5, nw = (nthr + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) { sh[2 * w] = v.hi; sh[2 * w + 1] = v.lo; }
    __syncthreads();
    dd r = dd_zero();
    if (tid == 0) {
        for (int q = 0; q < nw; ++q) dd_add_dd(r, dd{sh[2 * q], sh[2 * q + 1]});
    }
    return r;
}

// Cross-rank part of a reduction, run by warp 0 of the last block (ec3d_comm.cuh): this rank's
// double-double partials sh[0 .. 2*NRED) go to every rank's CommBlock, all contributions are awaited and
// summed in rank order; the rounded sums land in sc->red[slot*].
template <int NRED>
__device__ __forceinline__ void xchg_reduce_warp(const PeerTable &pt, CommLocal *cl, Scal *sc, const double *sh,
                                                 const int slot0, const int slot1, const int slot2)
{
    const int lane = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    unsigned long long e = 0;
    if (lane == 0) { e = cl->red_epoch + 1ull; cl->red_epoch = e; }
    e = __shfl_sync(0xffffffffu, e, 0);
    const int par = (int)(e & 1ull);
    if (lane < pt.nranks) {
        CommBlock *dst = pt.cb[lane];
#pragma unroll
        for (int q = 0; q < 2 * NRED; ++q) dst->red_val[par][pt.rank][q] = sh[q];
        __threadfence_system();
        st_release_sys(&dst->red_flag[pt.rank], e);
    }
    CommBlock *me = pt.cb[pt.rank];
    bool ok = true;
    if (lane < pt.nranks) ok = wait_epoch(&me->red_flag[lane], e);
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        if (!ok) cl->error = 1;
        const int slots[3] = {slot0, slot1, slot2};
#pragma unroll
        for (int q = 0; q < NRED; ++q) {
            dd t = dd_zero();                        // double-double sum in rank order, rounded once
            for (int r = 0; r < pt.nranks; ++r)
                dd_add_dd(t, dd{*(volatile double *)&me->red_val[par][r][2 * q], *(volatile double *)&me->red_val[par][r][2 * q + 1]});
            sc->red[slots[q]] = dd_round(t);
        }
        __threadfence();
    }
}

// Fused-exchange context of a kernel (several ranks over NVLink peer memory): pt == nullptr -> none.
// raise = bit mask (1 << HALO_*) of the halo kinds this kernel has pushed to the neighbours.
struct XchgCtx {
    const PeerTable *pt;
    CommLocal *cl;
    int raise;
};
__device__ __forceinline__ XchgCtx no_xchg() { return XchgCtx{nullptr, nullptr, 0}; }

// Stores this block's partial(s); the last block of the GROUP of kernels that share `sc->counter`
// (expected = total number of blocks that will call this with the same partials array) sums all
// partials and writes sc->red[slot0..2] (rounded).  Several ranks: with a fused-exchange context the
// last block also raises the halo flags of the vectors this kernel pushed and does the cross-rank sum
// itself; without one the unrounded double-double goes to red / red_lo for the exchange kernels.
// partials: 2 * NRED * pstride doubles.
template <int NRED>
__device__ __forceinline__ void reduce_epilogue(dd a0, dd a1, dd a2, double *partials, int pstride, int pidx,
                                                unsigned expected, Scal *sc, int slot0, int slot1, int slot2, double *sh,
                                                const XchgCtx xc = XchgCtx{nullptr, nullptr, 0})
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    dd s0 = block_sum(a0, sh);
    dd s1 = dd_zero(), s2 = dd_zero();
    if (NRED > 1) s1 = block_sum(a1, sh);
    if (NRED > 2) s2 = block_sum(a2, sh);
    __shared__ unsigned ticket_s;
    if (tid == 0) {
        partials[2 * pidx] = s0.hi; partials[2 * pidx + 1] = s0.lo;
        if (NRED > 1) { partials[2 * (pstride + pidx)] = s1.hi; partials[2 * (pstride + pidx) + 1] = s1.lo; }
        if (NRED > 2) { partials[2 * (2 * pstride + pidx)] = s2.hi; partials[2 * (2 * pstride + pidx) + 1] = s2.lo; }
        // thread 0 fences for the whole block: the barriers inside block_sum order every thread's stores
        // (own partials, fused halo push into the neighbours' memory) before this fence, and fences are
        // cumulative -- the pattern of a grid-wide barrier.  One system-scope fence per block, not per
        // thread (256 x thousands of blocks of membar.sys cost ~0.1 ms per SpMV on 8 GPUs).
        if (xc.pt) __threadfence_system();
        else __threadfence();
        ticket_s = atomicAdd(&sc->counter, 1u);
    }
    __syncthreads();
    if (ticket_s != expected - 1) return;
    __threadfence();
    if (xc.pt && xc.raise && tid == 0) {        // every block has fenced its stores: the halos are complete
        __threadfence_system();
        for (int k = 1; k < HALO_KINDS; ++k)
            if (xc.raise & (1 << k)) halo_raise(*xc.pt, xc.cl, k);
    }
    dd t0 = dd_zero(), t1 = dd_zero(), t2 = dd_zero();
    for (unsigned q = tid; q < expected; q += nthr) {
        dd_add_dd(t0, dd{__ldcg(partials + 2 * q), __ldcg(partials + 2 * q + 1)});
        if (NRED > 1) dd_add_dd(t1, dd{__ldcg(partials + 2 * (pstride + q)), __ldcg(partials + 2 * (pstride + q) + 1)});
        if (NRED > 2) dd_add_dd(t2, dd{__ldcg(partials + 2 * (2 * pstride + q)), __ldcg(partials + 2 * (2 * pstride + q) + 1)});
    }
    t0 = block_sum(t0, sh);
    if (NRED > 1) t1 = block_sum(t1, sh);
    if (NRED > 2) t2 = block_sum(t2, sh);
    if (xc.pt) {
        __syncthreads();
        if (tid == 0) {
            sh[0] = t0.hi; sh[1] = t0.lo; sh[2] = t1.hi; sh[3] = t1.lo; sh[4] = t2.hi; sh[5] = t2.lo;
            sc->counter = 0u;
        }
        __syncthreads();
        if (tid < 32) xchg_reduce_warp<NRED>(*xc.pt, xc.cl, sc, sh, slot0, slot1, slot2);
        return;
    }
    if (tid == 0) {
        if (sc->multi) {
            sc->red[slot0] = t0.hi; sc->red_lo[slot0] = t0.lo;
            if (NRED > 1) { sc->red[slot1] = t1.hi; sc->red_lo[slot1] = t1.lo; }
            if (NRED > 2) { sc->red[slot2] = t2.hi; sc->red_lo[slot2] = t2.lo; }
        } else {
            sc->red[slot0] = dd_round(t0);
            if (NRED > 1) sc->red[slot1] = dd_round(t1);
            if (NRED > 2) sc->red[slot2] = dd_round(t2);
        }
        sc->counter = 0u;
        __threadfence();
    }
}

// --------------------------------------------------------------------------------------------
// solver control: every kernel of iteration `it` decides from device scalars whether to run.
// --------------------------------------------------------------------------------------------
struct IterCtl {
    Scal *sc;
    const int *iter_base;  // device counter: iterations completed by previous launches
    int it_off;            // this kernel belongs to iteration *iter_base + it_off (1-based)
};

__device__ __forceinline__ bool is_block0()
{
    return (blockIdx.x | blockIdx.y | blockIdx.z) == 0 &&
           (threadIdx.x | threadIdx.y | threadIdx.z) == 0;
}

// guard of the first kernel of an iteration (SpMV A*p).  solvers.f90:23-29.
__device__ __forceinline__ bool guard_ap(const IterCtl &c, int &it)
{
    Scal *sc = c.sc;
    if (sc->done) return false;
    it = *c.iter_base + c.it_off;
    if (it == 1 && sc->red[RED_BB] == 0.0) {            // Bnorm == 0 -> RETURN with iter = 0
        if (is_block0()) { sc->exit_kind = 4; sc->final_iter = 0; sc->done = 1; }
        return false;
    }
    if (it - 1 > sc->itmax) {                           // IF (iter > itmax) ... EXIT
        if (is_block0()) { sc->exit_kind = 3; sc->final_iter = it - 1; sc->done = 1; }
        return false;
    }
    return true;
}

__device__ __forceinline__ bool s_converged(const Scal *sc)
{
    return sqrt(sc->red[RED_SS]) / sqrt(sc->red[RED_BB]) < sc->tol;   // solvers.f90:34
}

// --------------------------------------------------------------------------------------------
// SpMV row epilogue
// --------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void row_epilogue(double y, long long idx, double xc, const VecSet &vs, double &a0,
                                             double &a1)
{
    if (MODE == MODE_PLAIN) {
        vs.y[idx] = y;
    } else if (MODE == MODE_AP) {
        vs.y[idx] = y;
        a0 = DADD(a0, DMUL(y, __ldg(vs.r0 + idx)));              // (AP,R0)   solvers.f90:32
    } else if (MODE == MODE_AS) {
        vs.y[idx] = y;
        a0 = DADD(a0, DMUL(y, xc));                              // (AS,S)    solvers.f90:40
        a1 = DADD(a1, DMUL(y, y));                               // (AS,AS)
    } else {
        const double bb = __ldg(vs.b + idx);
        const double r = DSUB(bb, y);                            // R = B - A*X  solvers.f90:14-19
        vs.R[idx] = r; vs.R0[idx] = r; vs.P[idx] = r;
        a0 = DADD(a0, DMUL(bb, bb));                             // ||b||^2
        a1 = DADD(a1, DMUL(r, r));                               // (R,R0) with R0 = R
    }
}

template <int MODE>
__device__ __forceinline__ bool spmv_guard(const IterCtl &c)
{
    if (MODE == MODE_AP) { int it; return guard_ap(c, it); }
    if (MODE == MODE_AS) {
        if (c.sc->done) return false;
        if (s_converged(c.sc)) return false;
    }
    if (MODE == MODE_SAS) return !c.sc->done;      // ||s|| is only known after this kernel: A*s is speculative
    return true;
}

// --------------------------------------------------------------------------------------------
// K2a: matrix-free SpMV, non-conductor cells (7-point Laplacian with the reference's boundary
// rows, EC3D.f90:528-654).  One thread per (i,j) column of a TX x TY tile marching over `zc`
// planes with the k-1/k/k+1 values of the three components in registers; x/y neighbours come
// through L1.  Conductor cells are skipped here and handled by k_cond_spmv.
// --------------------------------------------------------------------------------------------
template <int MODE, int TX, int TY>
__global__ void __launch_bounds__(TX *TY)
k_air_spmv(const SlabGeom G, const Coef cf, const int *__restrict__ geo, const VecSet vs, const IterCtl ctl,
           const int zc, double *partials, const int pstride, const unsigned expected, const int finalize_here)
{
    __shared__ double sh[64];
    if (!spmv_guard<MODE>(ctl)) return;
    const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
    const int kb = G.k0 + blockIdx.z * zc;
    const int ke = min(kb + zc, G.k1);
    double a0 = 0.0, a1 = 0.0;
    if (i < G.sdx && j < G.sdy) {
        const int sdx = G.sdx, kdz = G.kdz;
        const long long col = (long long)j * sdx + i;
        const bool xl = (i == 0), xh = (i == sdx - 1), yl = (j == 0), yh = (j == G.sdy - 1);
        const double cxm = xh ? cf.bhi[0] : cf.msx, cxp = xl ? cf.blo[0] : cf.msx;
        const double cym = yh ? cf.bhi[1] : cf.msy, cyp = yl ? cf.blo[1] : cf.msy;
        const int bxy = (int)(xl | xh) | ((int)(yl | yh) << 1);
        const double *__restrict__ x0 = vs.x;
        const double *__restrict__ x1 = vs.x + G.segA;
        const double *__restrict__ x2 = vs.x + 2 * G.segA;
        long long p = (long long)(kb - G.k0 + 1) * kdz + col;     // local offset of (i,j,kb)
        const int *gp = geo + (long long)(kb - G.k0 + 2) * kdz + col;
        double c0 = x0[p], c1 = x1[p], c2 = x2[p];
        double m0 = 0.0, m1 = 0.0, m2 = 0.0;
        if (kb > 0) { m0 = x0[p - kdz]; m1 = x1[p - kdz]; m2 = x2[p - kdz]; }
        for (int k = kb; k < ke; ++k, p += kdz, gp += kdz) {
            const bool zl = (k == 0), zh = (k == G.sdz - 1);
            double p0 = 0.0, p1 = 0.0, p2 = 0.0;
            if (!zh) { p0 = x0[p + kdz]; p1 = x1[p + kdz]; p2 = x2[p + kdz]; }
            const int g = __ldg(gp);
            if (g == 0) {
                const double czm = zh ? cf.bhi[2] : cf.msz, czp = zl ? cf.blo[2] : cf.msz;
                const int bm = bxy | ((int)(zl | zh) << 2);
                const double dg = bm ? cf.diag_b[bm] : cf.diag_int;
#define AIR_ROW(XC, CM, CC, CP, COMP)                                                   \
                {                                                                       \
                    double s = 0.0;                                                     \
                    if (!zl) s = DADD(s, DMUL(czm, CM));                                \
                    if (!yl) s = DADD(s, DMUL(cym, XC[p - sdx]));                       \
                    if (!xl) s = DADD(s, DMUL(cxm, XC[p - 1]));                         \
                    s = DADD(s, DMUL(dg, CC));                                          \
                    if (!xh) s = DADD(s, DMUL(cxp, XC[p + 1]));                         \
                    if (!yh) s = DADD(s, DMUL(cyp, XC[p + sdx]));                       \
                    if (!zh) s = DADD(s, DMUL(czp, CP));                                \
                    row_epilogue<MODE>(s, (long long)(COMP) * G.segA + p, CC, vs, a0, a1); \
                }
                AIR_ROW(x0, m0, c0, p0, 0)
                AIR_ROW(x1, m1, c1, p1, 1)
                AIR_ROW(x2, m2, c2, p2, 2)
#undef AIR_ROW
            }
            m0 = c0; m1 = c1; m2 = c2;
            c0 = p0; c1 = p1; c2 = p2;
        }
    }
    if (MODE != MODE_PLAIN) {
        const int pidx = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        // INIT reduces (bb, rr), AS reduces (ass, asas), AP reduces (apr0)
        if (MODE == MODE_AP)
            reduce_epilogue<1>(dd{a0, 0.0}, dd_zero(), dd_zero(), partials, pstride, pidx, finalize_here ? expected : 0xffffffffu, ctl.sc,
                               RED_APR0, RED_APR0, RED_APR0, sh);
        else if (MODE == MODE_AS)
            reduce_epilogue<2>(dd{a0, 0.0}, dd{a1, 0.0}, dd_zero(), partials, pstride, pidx, finalize_here ? expected : 0xffffffffu, ctl.sc,
                               RED_ASS, RED_ASAS, RED_ASAS, sh);
        else
            reduce_epilogue<2>(dd{a0, 0.0}, dd{a1, 0.0}, dd_zero(), partials, pstride, pidx, finalize_here ? expected : 0xffffffffu, ctl.sc,
                               RED_BB, RED_RR_INIT, RED_RR_INIT, sh);
    }
}

// --------------------------------------------------------------------------------------------
// 16-byte helpers of the kernels that own two adjacent entries per thread.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }

// --------------------------------------------------------------------------------------------
// K2b: matrix-free SpMV, conductor cells: three A rows with convection, 2C/dt and grad-U coupling
// (EC3D.f90:656-710) plus the U row (EC3D.f90:766-922).  One thread per owned conductor cell.
// --------------------------------------------------------------------------------------------
struct GatherVisitor {      // used with a dense GeoView: U column g = 1 + offset in the dense U box
    const double *__restrict__ x;
    long long segA, offU, cell_shift;   // local A index = comp*segA + cell0 - cell_shift
    double s;
    __device__ __forceinline__ void a(int comp, long long cell0, double coef)
    {
        s = DADD(s, DMUL(coef, x[comp * segA + (cell0 - cell_shift)]));
    }
    __device__ __forceinline__ void u(int g, double coef) { s = DADD(s, DMUL(coef, x[offU + (g - 1)])); }
};

template <int MODE>
__global__ void __launch_bounds__(256)
k_cond_spmv(const SlabGeom G, const Coef cf, const MatCoef *__restrict__ mcs, const int *__restrict__ geo,
            const signed char *__restrict__ mat, const int *__restrict__ cond_cells, const int ncond,
            const VecSet vs, const IterCtl ctl, double *partials, const int pstride, const int pbase,
            const unsigned expected)
{
    __shared__ double sh[64];
    if (!spmv_guard<MODE>(ctl)) return;
    double a0 = 0.0, a1 = 0.0;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ncond) {
        const int cell0 = cond_cells[t];
        const int k = cell0 / G.kdz, rem = cell0 - k * G.kdz, j = rem / G.sdx, i = rem - j * G.sdx;
        const GeoView gv = dense_view(G, geo);
        const long long cell_shift = (long long)(G.k0 - 1) * G.kdz;
        const long long lp = (long long)cell0 - cell_shift;
        const long long lmat = (long long)cell0 - (long long)(G.k0 - 2) * G.kdz;
        const MatCoef mc = mcs[mat[lmat] - 1];
        GatherVisitor v{vs.x, G.segA, G.offU, cell_shift, 0.0};
#pragma unroll
        for (int comp = 0; comp < 3; ++comp) {
            v.s = 0.0;
            cond_a_row(mc, gv, i, j, k, comp, v);
            row_epilogue<MODE>(v.s, comp * G.segA + lp, vs.x[comp * G.segA + lp], vs, a0, a1);
        }
        v.s = 0.0;
        cond_u_row(cf, gv, i, j, k, 3, v);
        const long long lu = u_local(G, i, j, k);
        row_epilogue<MODE>(v.s, lu, vs.x[lu], vs, a0, a1);
    }
    if (MODE != MODE_PLAIN) {
        const int pidx = pbase + blockIdx.x;
        if (MODE == MODE_AP)
            reduce_epilogue<1>(dd{a0, 0.0}, dd_zero(), dd_zero(), partials, pstride, pidx, expected, ctl.sc, RED_APR0, RED_APR0, RED_APR0, sh);
        else if (MODE == MODE_AS)
            reduce_epilogue<2>(dd{a0, 0.0}, dd{a1, 0.0}, dd_zero(), partials, pstride, pidx, expected, ctl.sc, RED_ASS, RED_ASAS, RED_ASAS, sh);
        else
            reduce_epilogue<2>(dd{a0, 0.0}, dd{a1, 0.0}, dd_zero(), partials, pstride, pidx, expected, ctl.sc, RED_BB, RED_RR_INIT, RED_RR_INIT, sh);
    }
}

// --------------------------------------------------------------------------------------------
// K2': CSR SpMV on the reference's own arrays (drop-in mode, sprsbcgstabwr_).  A block owns 256
// consecutive rows; their (value, column) entries are one contiguous range of the CSR arrays, which
// the block streams through shared memory with coalesced loads; each thread then sums ITS row
// sequentially in stored order like sprsAx (solvers.f90:54-61), gathering x through L1/L2.
// --------------------------------------------------------------------------------------------
#define CSR_CH 4096
template <int MODE>
__global__ void __launch_bounds__(256)
k_csr_spmv(const int n, const int *__restrict__ irow, const int *__restrict__ jcol, const double *__restrict__ valA,
           const VecSet vs, const IterCtl ctl, double *partials, const int pstride, const unsigned expected)
{
    __shared__ double sh[64];
    __shared__ double sprod[CSR_CH];
    if (!spmv_guard<MODE>(ctl)) return;
    double a0 = 0.0, a1 = 0.0;
    const int r0 = blockIdx.x * blockDim.x;
    const int r = r0 + threadIdx.x;
    const int rlast = min(r0 + (int)blockDim.x, n);
    const long long lo = (long long)irow[r0] - 1, hi = (long long)irow[rlast] - 1;   // entries of this block's rows
    long long i1 = 0, i2 = 0;
    if (r < n) { i1 = (long long)irow[r] - 1; i2 = (long long)irow[r + 1] - 1; }
    double s = 0.0;
    for (long long c0 = lo; c0 < hi; c0 += CSR_CH) {
        const int len = (int)min((long long)CSR_CH, hi - c0);
        __syncthreads();
        // products of the whole range: coalesced matrix stream, independent gathers of x
        for (int q = threadIdx.x; q < len; q += blockDim.x)
            sprod[q] = DMUL(__ldg(valA + c0 + q), vs.x[__ldg(jcol + c0 + q) - 1]);
        __syncthreads();
        // this thread's row: sequential sum in stored order (valA*x rounded, then added -- as sprsAx)
        const long long ma = max(i1, c0), mb = min(i2, c0 + len);
        for (long long m = ma; m < mb; ++m) s = DADD(s, sprod[m - c0]);
    }
    if (r < n) row_epilogue<MODE>(s, r, vs.x[r], vs, a0, a1);
    if (MODE != MODE_PLAIN) {
        if (MODE == MODE_AP)
            reduce_epilogue<1>(dd{a0, 0.0}, dd_zero(), dd_zero(), partials, pstride, blockIdx.x, expected, ctl.sc, RED_APR0, RED_APR0, RED_APR0, sh);
        else if (MODE == MODE_AS)
            reduce_epilogue<2>(dd{a0, 0.0}, dd{a1, 0.0}, dd_zero(), partials, pstride, blockIdx.x, expected, ctl.sc, RED_ASS, RED_ASAS, RED_ASAS, sh);
        else
            reduce_epilogue<2>(dd{a0, 0.0}, dd{a1, 0.0}, dd_zero(), partials, pstride, blockIdx.x, expected, ctl.sc, RED_BB, RED_RR_INIT, RED_RR_INIT, sh);
    }
}

// --------------------------------------------------------------------------------------------
// fused BiCGSTABwr vector kernels over the owned ranges of the segmented local vector
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ long long own_to_local(const SlabGeom &G, long long e)
{
    const int s = (int)(e >= G.own_cum[1]) + (int)(e >= G.own_cum[2]) + (int)(e >= G.own_cum[3]);
    return G.own_off[s] + (e - G.own_cum[s]);
}

// iteration space of a BLAS-1 kernel over `units` work units: grid-stride, or one contiguous range
// per block (each SM then touches few 2 MB pages instead of all of them)
struct QRange {
    long long q0, q1, step;
};
__device__ __forceinline__ QRange q_range(const SlabGeom &G, const long long units)
{
    if (G.vmap) {
        const long long per = (units + gridDim.x - 1) / gridDim.x;
        const long long a = blockIdx.x * per;
        return QRange{a + threadIdx.x, min(units, a + per), (long long)blockDim.x};
    }
    return QRange{blockIdx.x * (long long)blockDim.x + threadIdx.x, units, (long long)gridDim.x * blockDim.x};
}

// K3: alpha = rr0/(AP,R0); S = R - alpha*AP; ||S||^2.            solvers.f90:31-34
template <int VEC>
__global__ void __launch_bounds__(256)
k_s_update(const SlabGeom G, const double *__restrict__ R, const double *__restrict__ AP, double *__restrict__ S,
           const IterCtl ctl, double *partials, const int pstride, const unsigned expected)
{
    __shared__ double sh[64];
    Scal *sc = ctl.sc;
    if (sc->done) return;
    const int it = *ctl.iter_base + ctl.it_off;
    const double rr0 = (it == 1) ? sc->red[RED_RR_INIT] : sc->rr0[it & 1];
    const double alpha = rr0 / sc->red[RED_APR0];
    if (is_block0()) sc->alpha = alpha;
    dd acc = dd_zero();
    const long long units = G.n_own / VEC;
    const QRange qr = q_range(G, units);
    for (long long q = qr.q0; q < qr.q1; q += qr.step) {
        const long long l = own_to_local(G, q * VEC);
        if (VEC == 2) {
            const double2 r = *reinterpret_cast<const double2 *>(R + l);
            const double2 ap = *reinterpret_cast<const double2 *>(AP + l);
            double2 s;
            s.x = DSUB(r.x, DMUL(alpha, ap.x));
            s.y = DSUB(r.y, DMUL(alpha, ap.y));
            *reinterpret_cast<double2 *>(S + l) = s;
            dd_add_d(acc, __fma_rn(s.y, s.y, DMUL(s.x, s.x)));      // pair sum: fixed, partition independent
        } else {
            const double s = DSUB(R[l], DMUL(alpha, AP[l]));
            S[l] = s;
            dd_add_d(acc, DMUL(s, s));
        }
    }
    reduce_epilogue<1>(acc, dd_zero(), dd_zero(), partials, pstride, blockIdx.x, expected, sc, RED_SS, RED_SS, RED_SS, sh);
}

// K5: if ||S|| converged: X = X + alpha*P (solvers.f90:36).  Else omega = (AS,S)/(AS,AS);
// X = X + alpha*P + omega*S; R = S - omega*AS; ||R||^2, (R,R0).   solvers.f90:40-44
template <int VEC>
__global__ void __launch_bounds__(256)
k_xr_update(const SlabGeom G, double *__restrict__ X, const double *__restrict__ P, const double *__restrict__ S,
            const double *__restrict__ AS, double *__restrict__ R, const double *__restrict__ R0,
            const IterCtl ctl, double *partials, const int pstride, const unsigned expected)
{
    __shared__ double sh[64];
    Scal *sc = ctl.sc;
    if (sc->done) return;
    const double alpha = sc->alpha;
    const long long units = G.n_own / VEC;
    const QRange qr = q_range(G, units);
    const long long q0 = qr.q0, qs = qr.step, qe = qr.q1;
    if (s_converged(sc)) {
        for (long long q = q0; q < qe; q += qs) {
            const long long l = own_to_local(G, q * VEC);
#pragma unroll
            for (int v = 0; v < VEC; ++v) X[l + v] = DADD(X[l + v], DMUL(alpha, P[l + v]));
        }
        return;
    }
    const double omega = sc->red[RED_ASS] / sc->red[RED_ASAS];
    if (is_block0()) sc->omega = omega;
    dd a0 = dd_zero(), a1 = dd_zero();
    for (long long q = q0; q < qe; q += qs) {
        const long long l = own_to_local(G, q * VEC);
        if (VEC == 2) {
            double2 x = *reinterpret_cast<const double2 *>(X + l);
            const double2 p = *reinterpret_cast<const double2 *>(P + l);
            const double2 s = *reinterpret_cast<const double2 *>(S + l);
            const double2 as = *reinterpret_cast<const double2 *>(AS + l);
            const double2 r0 = *reinterpret_cast<const double2 *>(R0 + l);
            double2 r;
            x.x = DADD(DADD(x.x, DMUL(alpha, p.x)), DMUL(omega, s.x));
            x.y = DADD(DADD(x.y, DMUL(alpha, p.y)), DMUL(omega, s.y));
            r.x = DSUB(s.x, DMUL(omega, as.x));
            r.y = DSUB(s.y, DMUL(omega, as.y));
            *reinterpret_cast<double2 *>(X + l) = x;
            *reinterpret_cast<double2 *>(R + l) = r;
            dd_add_d(a0, __fma_rn(r.y, r.y, DMUL(r.x, r.x)));
            dd_add_d(a1, __fma_rn(r.y, r0.y, DMUL(r.x, r0.x)));
        } else {
            const double s = S[l];
            X[l] = DADD(DADD(X[l], DMUL(alpha, P[l])), DMUL(omega, s));
            const double r = DSUB(s, DMUL(omega, AS[l]));
            R[l] = r;
            dd_add_d(a0, DMUL(r, r));
            dd_add_d(a1, DMUL(r, R0[l]));
        }
    }
    reduce_epilogue<2>(a0, a1, dd_zero(), partials, pstride, blockIdx.x, expected, sc, RED_RR, RED_RR0N, RED_RR0N, sh);
}

// K6: exit tests, beta, P = R + beta*(P - omega*AP), restart.     solvers.f90:34-49
template <int VEC>
__global__ void __launch_bounds__(256)
k_p_update(const SlabGeom G, double *__restrict__ P, const double *__restrict__ R, const double *__restrict__ AP,
           double *__restrict__ R0, const IterCtl ctl)
{
    Scal *sc = ctl.sc;
    if (sc->done) return;
    const int it = *ctl.iter_base + ctl.it_off;
    const double bnorm = sqrt(sc->red[RED_BB]);
    if (s_converged(sc)) {                                              // exit taken at solvers.f90:34-38
        if (is_block0()) { sc->exit_kind = 1; sc->final_iter = it; sc->done = 1; }
        return;
    }
    if (sqrt(sc->red[RED_RR]) / bnorm < sc->tol) {                      // solvers.f90:43
        if (is_block0()) { sc->exit_kind = 2; sc->final_iter = it; sc->done = 1; }
        return;
    }
    const double rr0 = (it == 1) ? sc->red[RED_RR_INIT] : sc->rr0[it & 1];
    const double rr0n = sc->red[RED_RR0N];
    const double alpha = sc->alpha, omega = sc->omega;
    const double beta = (alpha / omega) * rr0n / rr0;                   // solvers.f90:45
    const bool restart = fabs(rr0n) / bnorm < sc->tol;                  // solvers.f90:47
    if (is_block0()) {
        sc->beta = beta;
        // next (R,R0): after a restart R0 = R so it is ||R||^2, otherwise rr0_new (solvers.f90:31)
        sc->rr0[(it + 1) & 1] = restart ? sc->red[RED_RR] : rr0n;
        if (restart) sc->restarts += 1;
    }
    const long long units = G.n_own / VEC;
    const QRange qr = q_range(G, units);
    for (long long q = qr.q0; q < qr.q1; q += qr.step) {
        const long long l = own_to_local(G, q * VEC);
        if (VEC == 2) {
            const double2 r = *reinterpret_cast<const double2 *>(R + l);
            if (restart) {
                *reinterpret_cast<double2 *>(R0 + l) = r;
                *reinterpret_cast<double2 *>(P + l) = r;
            } else {
                const double2 p = *reinterpret_cast<const double2 *>(P + l);
                const double2 ap = *reinterpret_cast<const double2 *>(AP + l);
                double2 o;
                o.x = DADD(r.x, DMUL(beta, DSUB(p.x, DMUL(omega, ap.x))));
                o.y = DADD(r.y, DMUL(beta, DSUB(p.y, DMUL(omega, ap.y))));
                *reinterpret_cast<double2 *>(P + l) = o;
            }
        } else {
            const double r = R[l];
            if (restart) {
                R0[l] = r;
                P[l] = r;
            } else {
                P[l] = DADD(r, DMUL(beta, DSUB(P[l], DMUL(omega, AP[l]))));
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// TMA-ring BLAS-1 kernels (main path).  The grid-stride kernels above keep one 16-byte load per
// stream and thread in flight and, on 3.4 GB vectors whose bases differ by multiples of 0.5 MiB,
// collide in the L2-slice / DRAM-channel hash: 0.78-0.83 of the copy peak on plate(512) with DRAM
// reads 15-21 % above the algorithmic bytes (profiles/r02_ncu_blas1_plate512_r01kernels.txt).
// Here a persistent CTA per SM streams chunks of CH doubles of every input vector into a ring of
// shared-memory stages with linear bulk copies (cp.async.bulk, completed on an mbarrier); the
// copy engine keeps (NST-1) x NIN x 8 KB in flight per SM regardless of registers, and the same
// arithmetic runs out of shared memory.  Measured in isolation (scripts/stream_probe.cu): 1.00-1.05
// of the measured copy peak at 0.4 GB and 3.4 GB per vector, any vector spacing.
// Chunks never straddle the four owned segments of the local layout; chunk c is processed by CTA
// c % gridDim.x so that concurrently running CTAs stream neighbouring addresses.  The dot products
// keep the pair grouping of the grid-stride kernels (entries 2u, 2u+1 of the owned numbering, one
// fma; everything above in double-double), i.e. the same rounded results.
// --------------------------------------------------------------------------------------------
constexpr int V1_CH = 1024;                      // doubles per chunk and stream (8 KB)

struct ChunkMap {                                // chunks of the four owned segments
    long long cum[5];
    __device__ __forceinline__ void init(const SlabGeom &G)
    {
        cum[0] = 0;
#pragma unroll
        for (int s = 0; s < 4; ++s) cum[s + 1] = cum[s] + (G.own_len[s] + V1_CH - 1) / V1_CH;
    }
    // local offset and length (doubles, even) of chunk c; seg / o = owned segment and offset inside it
    __device__ __forceinline__ void locate(const SlabGeom &G, long long c, long long &loc, int &len, int &seg, long long &o) const
    {
        const int s = (int)(c >= cum[1]) + (int)(c >= cum[2]) + (int)(c >= cum[3]);
        const long long first = (s == 0) ? 0 : (s == 1) ? cum[1] : (s == 2) ? cum[2] : cum[3];   // (no dynamic indexing: registers)
        o = (c - first) * V1_CH;
        seg = s;
        loc = G.own_off[s] + o;
        len = (int)min((long long)V1_CH, G.own_len[s] - o);
    }
    __device__ __forceinline__ void locate(const SlabGeom &G, long long c, long long &loc, int &len) const
    {
        int seg; long long o;
        locate(G, c, loc, len, seg, o);
    }
};

template <int NIN, int NST>
struct BulkRing {
    static constexpr int STAGE_BYTES = NIN * V1_CH * 8;
    unsigned char *smem;
    unsigned long long *full, *empty;
    const double *in[NIN];
    ChunkMap cm;
    long long c0, cs, mine;

    __device__ __forceinline__ void issue(const SlabGeom &G, long long i)
    {
        const int s = (int)(i % NST);
        long long loc; int len;
        cm.locate(G, c0 + i * cs, loc, len);
        unsigned char *st = smem + s * STAGE_BYTES;
        tma::mbar_expect_tx(full + s, (uint32_t)(NIN * len * 8));
#pragma unroll
        for (int v = 0; v < NIN; ++v) tma::bulk_load(st + v * V1_CH * 8, in[v] + loc, (uint32_t)(len * 8), full + s);
    }
    __device__ __forceinline__ void start(const SlabGeom &G, unsigned char *smem_, unsigned long long *full_,
                                          unsigned long long *empty_)
    {
        smem = smem_; full = full_; empty = empty_;
        cm.init(G);
        c0 = blockIdx.x; cs = gridDim.x;
        const long long nch = cm.cum[4];
        mine = (nch > c0) ? (nch - c0 + cs - 1) / cs : 0;
        if (threadIdx.x == 0) {
            for (int s = 0; s < NST; ++s) { tma::mbar_init(full + s, 1); tma::mbar_init(empty + s, blockDim.x / 32); }
            tma::fence_barrier_init();
            tma::fence_proxy_async();
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (long long i = 0; i < min((long long)NST, mine); ++i) issue(G, i);
    }
    // wait for chunk i; returns its stage
    __device__ __forceinline__ const double *acquire(long long i) const
    {
        tma::mbar_wait(full + (int)(i % NST), (uint32_t)((i / NST) & 1));
        return reinterpret_cast<const double *>(smem + (int)(i % NST) * STAGE_BYTES);
    }
    // every warp calls this once it has read chunk i into registers; thread 0 re-arms the stage
    // (consumer arrive + producer wait on the `empty` barrier order the generic-proxy reads before
    // the async-proxy refill)
    __device__ __forceinline__ void release(const SlabGeom &G, long long i)
    {
        const int s = (int)(i % NST);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) tma::mbar_arrive(empty + s);
        if (threadIdx.x == 0 && i + NST < mine) {
            tma::mbar_wait(empty + s, (uint32_t)((i / NST) & 1));
            issue(G, i + NST);
        }
    }
};

constexpr int XR_NST = 3, P_NST = 4, S_NST = 4;
constexpr int v1_smem_bytes(int nin, int nst) { return nin * V1_CH * 8 * nst + 128; }

__device__ __forceinline__ unsigned char *align128(unsigned char *p)
{
    return p + ((128u - (tma::smem_u32(p) & 127u)) & 127u);
}

// K5 (TMA ring): X = X + alpha*P [+ omega*S]; R = S - omega*AS; ||R||^2, (R,R0).   solvers.f90:34-44
__global__ void __launch_bounds__(256, 1)
k_xr_update_tma(const SlabGeom G, double *__restrict__ X, const double *__restrict__ P, const double *__restrict__ S,
                const double *__restrict__ AS, double *__restrict__ R, const double *__restrict__ R0, const IterCtl ctl,
                double *partials, const int pstride, const unsigned expected,
                const __grid_constant__ PeerTable pt, CommLocal *cl, const int xf, const int vi_R)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ double sh[64];
    __shared__ __align__(8) unsigned long long full[XR_NST], empty[XR_NST];
    Scal *sc = ctl.sc;
    if (sc->done) return;
    const double alpha = sc->alpha;
    const int tid = threadIdx.x;
    if (s_converged(sc)) {                                  // exit of solvers.f90:34-38: X = X + alpha*P only
        ChunkMap cm; cm.init(G);
        for (long long c = blockIdx.x; c < cm.cum[4]; c += gridDim.x) {
            long long loc; int len;
            cm.locate(G, c, loc, len);
            for (int e = 2 * tid; e < len; e += 512) {
                double2 x = ld2(X + loc + e);
                const double2 p = ld2(P + loc + e);
                st2(X + loc + e, DADD(x.x, DMUL(alpha, p.x)), DADD(x.y, DMUL(alpha, p.y)));
            }
        }
        return;
    }
    const double omega = sc->red[RED_ASS] / sc->red[RED_ASAS];
    if (is_block0()) sc->omega = omega;
    BulkRing<5, XR_NST> ring;
    ring.in[0] = X; ring.in[1] = P; ring.in[2] = S; ring.in[3] = AS; ring.in[4] = R0;
    ring.start(G, align128(smem_raw), full, empty);
    dd a0 = dd_zero(), a1 = dd_zero();
    for (long long i = 0; i < ring.mine; ++i) {
        const double *st = ring.acquire(i);
        long long loc, o0; int len, seg;
        ring.cm.locate(G, ring.c0 + i * ring.cs, loc, len, seg, o0);
        double2 x[2], p[2], s[2], as[2], r0[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = 2 * (tid + u * 256);
            if (e < len) {
                x[u] = ld2(st + e); p[u] = ld2(st + V1_CH + e); s[u] = ld2(st + 2 * V1_CH + e);
                as[u] = ld2(st + 3 * V1_CH + e); r0[u] = ld2(st + 4 * V1_CH + e);
            }
        }
        ring.release(G, i);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = 2 * (tid + u * 256);
            if (e < len) {
                double2 r;
                x[u].x = DADD(DADD(x[u].x, DMUL(alpha, p[u].x)), DMUL(omega, s[u].x));
                x[u].y = DADD(DADD(x[u].y, DMUL(alpha, p[u].y)), DMUL(omega, s[u].y));
                r.x = DSUB(s[u].x, DMUL(omega, as[u].x));
                r.y = DSUB(s[u].y, DMUL(omega, as[u].y));
                st2(X + loc + e, x[u].x, x[u].y);
                st2(R + loc + e, r.x, r.y);
                if (xf) peer_push2(pt, G, vi_R, seg, o0 + e, r.x, r.y);      // R is an input of the next A*s SpMV
                dd_add_d(a0, __fma_rn(r.y, r.y, DMUL(r.x, r.x)));
                dd_add_d(a1, __fma_rn(r.y, r0[u].y, DMUL(r.x, r0[u].x)));
            }
        }
    }
    reduce_epilogue<2>(a0, a1, dd_zero(), partials, pstride, blockIdx.x, expected, sc, RED_RR, RED_RR0N, RED_RR0N, sh,
                       xf ? XchgCtx{&pt, cl, 1 << HALO_R} : no_xchg());
}

// K6 (TMA ring): exit tests, beta, P = R + beta*(P - omega*AP), restart.     solvers.f90:34-49
// all blocks of the p-update have fenced their peer stores: the last one tells the neighbours
__device__ __forceinline__ void p_push_done(const PeerTable &pt, CommLocal *cl)
{
    __syncthreads();                                   // (thread 0's fence below covers the block, see reduce_epilogue)
    if (threadIdx.x != 0) return;
    __threadfence_system();
    if (atomicAdd(&cl->ticket_p, 1u) != gridDim.x - 1) return;
    cl->ticket_p = 0u;
    __threadfence_system();
    halo_raise(pt, cl, HALO_P);
}

__global__ void __launch_bounds__(256, 2)
k_p_update_tma(const SlabGeom G, double *__restrict__ P, const double *__restrict__ R, const double *__restrict__ AP,
               double *__restrict__ R0, const IterCtl ctl,
               const __grid_constant__ PeerTable pt, CommLocal *cl, const int xf, const int vi_P)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full[P_NST], empty[P_NST];
    Scal *sc = ctl.sc;
    if (sc->done) return;
    const int it = *ctl.iter_base + ctl.it_off;
    const double bnorm = sqrt(sc->red[RED_BB]);
    if (s_converged(sc)) {                                              // exit taken at solvers.f90:34-38
        if (is_block0()) { sc->exit_kind = 1; sc->final_iter = it; sc->done = 1; }
        return;
    }
    if (sqrt(sc->red[RED_RR]) / bnorm < sc->tol) {                      // solvers.f90:43
        if (is_block0()) { sc->exit_kind = 2; sc->final_iter = it; sc->done = 1; }
        return;
    }
    const double rr0 = (it == 1) ? sc->red[RED_RR_INIT] : sc->rr0[it & 1];
    const double rr0n = sc->red[RED_RR0N];
    const double alpha = sc->alpha, omega = sc->omega;
    const double beta = (alpha / omega) * rr0n / rr0;                   // solvers.f90:45
    const bool restart = fabs(rr0n) / bnorm < sc->tol;                  // solvers.f90:47
    if (is_block0()) {
        sc->beta = beta;
        sc->rr0[(it + 1) & 1] = restart ? sc->red[RED_RR] : rr0n;      // next (R,R0), solvers.f90:31
        if (restart) sc->restarts += 1;
    }
    const int tid = threadIdx.x;
    if (restart) {                                                      // R0 = R; P = R   (rare: plain loads)
        ChunkMap cm; cm.init(G);
        for (long long c = blockIdx.x; c < cm.cum[4]; c += gridDim.x) {
            long long loc, o0; int len, seg;
            cm.locate(G, c, loc, len, seg, o0);
            for (int e = 2 * tid; e < len; e += 512) {
                const double2 r = ld2(R + loc + e);
                st2(R0 + loc + e, r.x, r.y);
                st2(P + loc + e, r.x, r.y);
                if (xf) peer_push2(pt, G, vi_P, seg, o0 + e, r.x, r.y);
            }
        }
        if (xf) p_push_done(pt, cl);
        return;
    }
    BulkRing<3, P_NST> ring;
    ring.in[0] = R; ring.in[1] = P; ring.in[2] = AP;
    ring.start(G, align128(smem_raw), full, empty);
    for (long long i = 0; i < ring.mine; ++i) {
        const double *st = ring.acquire(i);
        long long loc, o0; int len, seg;
        ring.cm.locate(G, ring.c0 + i * ring.cs, loc, len, seg, o0);
        double2 r[2], p[2], ap[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = 2 * (tid + u * 256);
            if (e < len) { r[u] = ld2(st + e); p[u] = ld2(st + V1_CH + e); ap[u] = ld2(st + 2 * V1_CH + e); }
        }
        ring.release(G, i);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = 2 * (tid + u * 256);
            if (e < len) {
                const double px = DADD(r[u].x, DMUL(beta, DSUB(p[u].x, DMUL(omega, ap[u].x))));
                const double py = DADD(r[u].y, DMUL(beta, DSUB(p[u].y, DMUL(omega, ap[u].y))));
                st2(P + loc + e, px, py);
                if (xf) peer_push2(pt, G, vi_P, seg, o0 + e, px, py);        // P is the input of the next A*p SpMV
            }
        }
    }
    if (xf) p_push_done(pt, cl);
}

// K3 (TMA ring; CSR drop-in path, where the SpMV is not the fused MODE_SAS kernel):
// alpha = rr0/(AP,R0); S = R - alpha*AP; ||S||^2.            solvers.f90:31-34
__global__ void __launch_bounds__(256, 2)
k_s_update_tma(const SlabGeom G, const double *__restrict__ R, const double *__restrict__ AP, double *__restrict__ S,
               const IterCtl ctl, double *partials, const int pstride, const unsigned expected)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ double sh[64];
    __shared__ __align__(8) unsigned long long full[S_NST], empty[S_NST];
    Scal *sc = ctl.sc;
    if (sc->done) return;
    const int it = *ctl.iter_base + ctl.it_off;
    const double rr0 = (it == 1) ? sc->red[RED_RR_INIT] : sc->rr0[it & 1];
    const double alpha = rr0 / sc->red[RED_APR0];
    if (is_block0()) sc->alpha = alpha;
    const int tid = threadIdx.x;
    BulkRing<2, S_NST> ring;
    ring.in[0] = R; ring.in[1] = AP;
    ring.start(G, align128(smem_raw), full, empty);
    dd acc = dd_zero();
    for (long long i = 0; i < ring.mine; ++i) {
        const double *st = ring.acquire(i);
        long long loc; int len;
        ring.cm.locate(G, ring.c0 + i * ring.cs, loc, len);
        double2 r[2], ap[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = 2 * (tid + u * 256);
            if (e < len) { r[u] = ld2(st + e); ap[u] = ld2(st + V1_CH + e); }
        }
        ring.release(G, i);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = 2 * (tid + u * 256);
            if (e < len) {
                const double sx = DSUB(r[u].x, DMUL(alpha, ap[u].x)), sy = DSUB(r[u].y, DMUL(alpha, ap[u].y));
                st2(S + loc + e, sx, sy);
                dd_add_d(acc, __fma_rn(sy, sy, DMUL(sx, sx)));
            }
        }
    }
    reduce_epilogue<1>(acc, dd_zero(), dd_zero(), partials, pstride, blockIdx.x, expected, sc, RED_SS, RED_SS, RED_SS, sh);
}

__global__ void k_solver_reset(Scal *sc, int *iter_base, double tol, int itmax, int multi)
{
    for (int q = 0; q < 8; ++q) { sc->red[q] = 0.0; sc->red_lo[q] = 0.0; }
    sc->multi = multi;
    sc->rr0[0] = sc->rr0[1] = 0.0;
    sc->alpha = sc->omega = sc->beta = 0.0;
    sc->tol = tol; sc->itmax = itmax;
    sc->done = 0; sc->final_iter = 0; sc->exit_kind = 0; sc->restarts = 0; sc->counter = 0u;
    *iter_base = 0;
}

__global__ void k_iter_advance(int *iter_base, int by) { *iter_base += by; }

// NCCL path of the cross-rank reduction: (hi, lo) of up to three results <-> gather buffer
#define RED_W 6     // doubles exchanged per rank and reduction point: up to 3 results x (hi, lo)
__global__ void k_red_pack(const Scal *sc, int slot, int count, double *mine /* RED_W doubles */)
{
    if (threadIdx.x != 0) return;
    for (int q = 0; q < RED_W / 2; ++q) {
        mine[2 * q] = q < count ? sc->red[slot + q] : 0.0;
        mine[2 * q + 1] = q < count ? sc->red_lo[slot + q] : 0.0;
    }
}
__global__ void k_red_unpack(Scal *sc, int slot, int count, const double *all /* [nranks][RED_W] */, int nranks)
{
    if (threadIdx.x != 0) return;
    for (int q = 0; q < count; ++q) {
        dd t = dd_zero();
        for (int r = 0; r < nranks; ++r) dd_add_dd(t, dd{all[RED_W * r + 2 * q], all[RED_W * r + 2 * q + 1]});
        sc->red[slot + q] = dd_round(t);
    }
}

// generic helpers
__global__ void k_fill(double *p, long long n, double v)
{
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
        p[q] = v;
}

// Deterministic pseudo-random fill, uniform in (-1, 1) (splitmix64 of the index): the measurement hooks time
// the kernels on data with realistic bit activity -- constant-valued vectors draw ~25 % less power on a
// B200 and keep it out of the power cap the real solves run into.
__global__ void k_fill_random(double *p, long long n, unsigned long long seed)
{
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(q + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        p[q] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
}
