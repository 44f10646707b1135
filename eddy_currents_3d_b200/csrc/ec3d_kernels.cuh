// ec3d_kernels.cuh -- sm_100a kernels of the EC3D hot path: matrix-free SpMV (air + conductor
// parts), CSR SpMV (drop-in mode), fused BiCGSTABwr vector kernels, deterministic reductions.
//
// Arithmetic: fp64 with explicit round-to-nearest mul/add (__dmul_rn/__dadd_rn are never
// contracted into FMA), each row summed in ascending column order starting from 0 -- the order of
// the reference's sprsAx (solvers.f90:54-61) -- so the matrix-free operator is bit-identical to a
// sequential CSR row sum.  Dot products use a fixed reduction tree (thread-sequential, warp
// butterfly, warp-0 across warps, last block over the per-block partials in index order), hence
// results are reproducible run to run; they differ from the reference's sequential dot_product
// only by summation order.
#pragma once
#include "ec3d_common.cuh"
#include "ec3d_rows.cuh"

#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))

enum { MODE_PLAIN = 0, MODE_AP = 1, MODE_AS = 2, MODE_INIT = 3 };

struct VecSet {
    const double *x;   // SpMV input
    double *y;         // SpMV output (AP / AS / plain)
    const double *r0;  // MODE_AP
    const double *b;   // MODE_INIT
    double *R, *R0, *P;// MODE_INIT outputs
};

// --------------------------------------------------------------------------------------------
// reductions
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = DADD(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum over the block; result valid in thread 0.  blockDim.x*blockDim.y*blockDim.z <= 1024.
__device__ __forceinline__ double block_sum(double v, double *sh /* >= 32 doubles */)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, w = tid >> 5, nw = (nthr + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (tid == 0) {
        for (int q = 0; q < nw; ++q) r = DADD(r, sh[q]);
    }
    return r;
}

// Stores this block's partial(s); the last block of the GROUP of kernels that share `sc->counter`
// (expected = total number of blocks that will call this with the same partials array) sums all
// partials in index order and writes sc->red[slot0], sc->red[slot1].
template <int NRED>
__device__ __forceinline__ void reduce_epilogue(double a0, double a1, double *partials, int pstride, int pidx,
                                                unsigned expected, Scal *sc, int slot0, int slot1, double *sh)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    double s0 = block_sum(a0, sh);
    double s1 = 0.0;
    if (NRED > 1) s1 = block_sum(a1, sh);
    __shared__ unsigned ticket_s;
    if (tid == 0) {
        partials[pidx] = s0;
        if (NRED > 1) partials[pstride + pidx] = s1;
        __threadfence();
        ticket_s = atomicAdd(&sc->counter, 1u);
    }
    __syncthreads();
    if (ticket_s != expected - 1) return;
    __threadfence();
    double t0 = 0.0, t1 = 0.0;
    for (unsigned q = tid; q < expected; q += nthr) {
        t0 = DADD(t0, __ldcg(partials + q));
        if (NRED > 1) t1 = DADD(t1, __ldcg(partials + pstride + q));
    }
    t0 = block_sum(t0, sh);
    if (NRED > 1) t1 = block_sum(t1, sh);
    if (tid == 0) {
        sc->red[slot0] = t0;
        if (NRED > 1) sc->red[slot1] = t1;
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
    __shared__ double sh[32];
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
            reduce_epilogue<1>(a0, 0.0, partials, pstride, pidx, finalize_here ? expected : 0xffffffffu, ctl.sc,
                               RED_APR0, RED_APR0, sh);
        else if (MODE == MODE_AS)
            reduce_epilogue<2>(a0, a1, partials, pstride, pidx, finalize_here ? expected : 0xffffffffu, ctl.sc,
                               RED_ASS, RED_ASAS, sh);
        else
            reduce_epilogue<2>(a0, a1, partials, pstride, pidx, finalize_here ? expected : 0xffffffffu, ctl.sc,
                               RED_BB, RED_RR_INIT, sh);
    }
}

// --------------------------------------------------------------------------------------------
// K2 (main path): fused matrix-free SpMV.  Each thread owns TWO x-adjacent cells of an (i,j)
// column and marches over `zc` planes: the k-1 / k / k+1 values of the three components live in
// registers, plane k+2 is prefetched one iteration ahead (16-byte loads), x/y neighbours come
// through L1.  Per cell the 1-byte class map selects
//   0  air or domain-face cell ...... the reference's boundary / interior Laplacian rows
//   1  interior conductor cell ...... convection + 2C/dt + central grad-U rows and the 13-entry
//                                     U row (EC3D.f90:649-680, 917-922), all gathers inline
//   2  conductor-surface cell ....... skipped here; k_cond_spmv handles the (short) surface list
// Rows are summed in ascending column order with unfused mul/add, so results equal the CSR sums.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }

template <int MODE>
__device__ __forceinline__ void pair_epilogue(double ya, double yb, bool wa, bool wb, long long idx, double xa,
                                              double xb, const VecSet &vs, double &a0, double &a1)
{
    if (wa && wb) {
        if (MODE == MODE_PLAIN) {
            st2(vs.y + idx, ya, yb);
        } else if (MODE == MODE_AP) {
            st2(vs.y + idx, ya, yb);
            const double2 r0 = ld2(vs.r0 + idx);
            a0 = DADD(a0, DMUL(ya, r0.x));
            a0 = DADD(a0, DMUL(yb, r0.y));
        } else if (MODE == MODE_AS) {
            st2(vs.y + idx, ya, yb);
            a0 = DADD(a0, DMUL(ya, xa)); a1 = DADD(a1, DMUL(ya, ya));
            a0 = DADD(a0, DMUL(yb, xb)); a1 = DADD(a1, DMUL(yb, yb));
        } else {
            const double2 bb = ld2(vs.b + idx);
            const double ra = DSUB(bb.x, ya), rb = DSUB(bb.y, yb);
            st2(vs.R + idx, ra, rb); st2(vs.R0 + idx, ra, rb); st2(vs.P + idx, ra, rb);
            a0 = DADD(a0, DMUL(bb.x, bb.x)); a1 = DADD(a1, DMUL(ra, ra));
            a0 = DADD(a0, DMUL(bb.y, bb.y)); a1 = DADD(a1, DMUL(rb, rb));
        }
    } else {
        if (wa) row_epilogue<MODE>(ya, idx, xa, vs, a0, a1);
        if (wb) row_epilogue<MODE>(yb, idx + 1, xb, vs, a0, a1);
    }
}

template <int MODE>
__device__ __forceinline__ void pair_epilogue_pre(double ya, double yb, bool wa, bool wb, long long idx, double xa,
                                                  double xb, double2 aux /* r0 (AP) or b (INIT) at idx */,
                                                  const VecSet &vs, double &a0, double &a1)
{
    // same arithmetic and accumulation order as row_epilogue for cell a then cell b
    if (MODE == MODE_INIT) {
        const double ra = DSUB(aux.x, ya), rb = DSUB(aux.y, yb);
        if (wa && wb) { st2(vs.R + idx, ra, rb); st2(vs.R0 + idx, ra, rb); st2(vs.P + idx, ra, rb); }
        else {
            if (wa) { vs.R[idx] = ra; vs.R0[idx] = ra; vs.P[idx] = ra; }
            if (wb) { vs.R[idx + 1] = rb; vs.R0[idx + 1] = rb; vs.P[idx + 1] = rb; }
        }
        if (wa) { a0 = DADD(a0, DMUL(aux.x, aux.x)); a1 = DADD(a1, DMUL(ra, ra)); }
        if (wb) { a0 = DADD(a0, DMUL(aux.y, aux.y)); a1 = DADD(a1, DMUL(rb, rb)); }
        return;
    }
    if (wa && wb) st2(vs.y + idx, ya, yb);
    else {
        if (wa) vs.y[idx] = ya;
        if (wb) vs.y[idx + 1] = yb;
    }
    if (MODE == MODE_AP) {
        if (wa) a0 = DADD(a0, DMUL(ya, aux.x));
        if (wb) a0 = DADD(a0, DMUL(yb, aux.y));
    } else if (MODE == MODE_AS) {
        if (wa) { a0 = DADD(a0, DMUL(ya, xa)); a1 = DADD(a1, DMUL(ya, ya)); }
        if (wb) { a0 = DADD(a0, DMUL(yb, xb)); a1 = DADD(a1, DMUL(yb, yb)); }
    }
}

template <int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_stencil2_spmv(const SlabGeom G, const Coef cf, const MatCoef mc, const unsigned char *__restrict__ cls,
                const int *__restrict__ geo, const VecSet vs, const IterCtl ctl, const int zc, double *partials,
                const int pstride, const unsigned expected, const int finalize_here)
{
    __shared__ double sh[32];
    if (!spmv_guard<MODE>(ctl)) return;
    const int i0 = (blockIdx.x * 32 + threadIdx.x) * 2, j = blockIdx.y * 8 + threadIdx.y;
    const int kb = G.k0 + blockIdx.z * zc;
    const int ke = min(kb + zc, G.k1);
    double a0 = 0.0, a1 = 0.0;
    if (i0 < G.sdx && j < G.sdy) {
        const int sdx = G.sdx, kdz = G.kdz, sdz = G.sdz;
        const long long col = (long long)j * sdx + i0;
        const bool xlA = (i0 == 0), xhB = (i0 + 2 == sdx), yl = (j == 0), yh = (j == G.sdy - 1);
        const double cxpA = xlA ? cf.blo[0] : cf.msx;       // on cell a's right neighbour (= cell b)
        const double cxmB = xhB ? cf.bhi[0] : cf.msx;       // on cell b's left neighbour (= cell a)
        const double cym = yh ? cf.bhi[1] : cf.msy, cyp = yl ? cf.blo[1] : cf.msy;
        const int by = (int)(yl | yh) << 1;
        const double *__restrict__ xv = vs.x;
        const long long ub = G.offU - (long long)G.gbase;    // U value of geoPHYS_C id g is xv[ub + g]
        long long p = (long long)(kb - G.k0 + 1) * kdz + col;
        double2 m[3], c[3], z1[3], z2[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double *X = xv + q * G.segA + p;
            c[q] = ld2(X);
            m[q] = (kb > 0) ? ld2(X - kdz) : make_double2(0.0, 0.0);
            z1[q] = (kb + 1 < sdz) ? ld2(X + kdz) : make_double2(0.0, 0.0);
        }
        for (int k = kb; k < ke; ++k, p += kdz) {
            const bool zl = (k == 0), zh = (k == sdz - 1);
            const bool pf = (k + 1 < ke) && (k + 2 < sdz);
#pragma unroll
            for (int q = 0; q < 3; ++q)
                z2[q] = pf ? ld2(xv + q * G.segA + p + 2 * (long long)kdz) : make_double2(0.0, 0.0);
            const uchar2 cl = *reinterpret_cast<const uchar2 *>(cls + (long long)(k - G.k0) * kdz + col);
            const double czm = zh ? cf.bhi[2] : cf.msz, czp = zl ? cf.blo[2] : cf.msz;
            const int bz = (int)(zl | zh) << 2;
            const int bmA = (int)xlA | by | bz, bmB = (int)xhB | by | bz;
            const double dgA = bmA ? cf.diag_b[bmA] : cf.diag_int, dgB = bmB ? cf.diag_b[bmB] : cf.diag_int;
            const bool fA = (cl.x == 1), fB = (cl.y == 1);
            const bool wA = (cl.x != 2), wB = (cl.y != 2);
            // U gathers of the interior-conductor path (whenever one cell of the pair is interior
            // conductor the other is a conductor cell too, numbered consecutively along x)
            int gA = 0;
            double uxm = 0.0, uA = 0.0, uB = 0.0, uxp = 0.0;
            double2 uym = make_double2(0.0, 0.0), uyp = uym, uzm = uym, uzp = uym;
            if (fA | fB) {
                const int *gp = geo + (long long)(k - G.k0 + 2) * kdz + col;
                gA = gp[0];
                const int2 gjm = *reinterpret_cast<const int2 *>(gp - sdx), gjp = *reinterpret_cast<const int2 *>(gp + sdx);
                const int2 gkm = *reinterpret_cast<const int2 *>(gp - kdz), gkp = *reinterpret_cast<const int2 *>(gp + kdz);
                uA = xv[ub + gA]; uB = xv[ub + gA + 1];
                if (fA) { uxm = xv[ub + gA - 1]; uym.x = xv[ub + gjm.x]; uyp.x = xv[ub + gjp.x]; uzm.x = xv[ub + gkm.x]; uzp.x = xv[ub + gkp.x]; }
                if (fB) { uxp = xv[ub + gA + 2]; uym.y = xv[ub + gjm.y]; uyp.y = xv[ub + gjp.y]; uzm.y = xv[ub + gkm.y]; uzp.y = xv[ub + gkp.y]; }
            }
            double suA = 0.0, suB = 0.0;      // running A part of the U rows (ascending columns)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const double *X = xv + q * G.segA + p;
                const double2 ym = yl ? make_double2(0.0, 0.0) : ld2(X - sdx);
                const double2 yp = yh ? make_double2(0.0, 0.0) : ld2(X + sdx);
                const double xm = xlA ? 0.0 : X[-1];
                const double xp = xhB ? 0.0 : X[2];
                double ya = 0.0, yb = 0.0;
                if (fA) {
                    ya = DADD(0.0, DMUL(mc.cm[2], m[q].x));
                    ya = DADD(ya, DMUL(mc.cm[1], ym.x));
                    ya = DADD(ya, DMUL(mc.cm[0], xm));
                    ya = DADD(ya, DMUL(mc.diag, c[q].x));
                    ya = DADD(ya, DMUL(mc.cp[0], c[q].y));
                    ya = DADD(ya, DMUL(mc.cp[1], yp.x));
                    ya = DADD(ya, DMUL(mc.cp[2], z1[q].x));
                    const double um = (q == 0) ? uxm : (q == 1) ? uym.x : uzm.x;
                    const double up = (q == 0) ? uB : (q == 1) ? uyp.x : uzp.x;
                    ya = DADD(ya, DMUL(mc.g1[q], um));
                    ya = DADD(ya, DMUL(-mc.g1[q], up));
                    const double am = (q == 0) ? xm : (q == 1) ? ym.x : m[q].x;
                    const double ap = (q == 0) ? c[q].y : (q == 1) ? yp.x : z1[q].x;
                    suA = DADD(suA, DMUL(cf.ua_p[q], am));
                    suA = DADD(suA, DMUL(cf.ua_m[q], ap));
                } else if (wA) {
                    if (!zl) ya = DADD(ya, DMUL(czm, m[q].x));
                    if (!yl) ya = DADD(ya, DMUL(cym, ym.x));
                    if (!xlA) ya = DADD(ya, DMUL(cf.msx, xm));
                    ya = DADD(ya, DMUL(dgA, c[q].x));
                    ya = DADD(ya, DMUL(cxpA, c[q].y));
                    if (!yh) ya = DADD(ya, DMUL(cyp, yp.x));
                    if (!zh) ya = DADD(ya, DMUL(czp, z1[q].x));
                }
                if (fB) {
                    yb = DADD(0.0, DMUL(mc.cm[2], m[q].y));
                    yb = DADD(yb, DMUL(mc.cm[1], ym.y));
                    yb = DADD(yb, DMUL(mc.cm[0], c[q].x));
                    yb = DADD(yb, DMUL(mc.diag, c[q].y));
                    yb = DADD(yb, DMUL(mc.cp[0], xp));
                    yb = DADD(yb, DMUL(mc.cp[1], yp.y));
                    yb = DADD(yb, DMUL(mc.cp[2], z1[q].y));
                    const double um = (q == 0) ? uA : (q == 1) ? uym.y : uzm.y;
                    const double up = (q == 0) ? uxp : (q == 1) ? uyp.y : uzp.y;
                    yb = DADD(yb, DMUL(mc.g1[q], um));
                    yb = DADD(yb, DMUL(-mc.g1[q], up));
                    const double am = (q == 0) ? c[q].x : (q == 1) ? ym.y : m[q].y;
                    const double ap = (q == 0) ? xp : (q == 1) ? yp.y : z1[q].y;
                    suB = DADD(suB, DMUL(cf.ua_p[q], am));
                    suB = DADD(suB, DMUL(cf.ua_m[q], ap));
                } else if (wB) {
                    if (!zl) yb = DADD(yb, DMUL(czm, m[q].y));
                    if (!yl) yb = DADD(yb, DMUL(cym, ym.y));
                    yb = DADD(yb, DMUL(cxmB, c[q].x));
                    yb = DADD(yb, DMUL(dgB, c[q].y));
                    if (!xhB) yb = DADD(yb, DMUL(cf.msx, xp));
                    if (!yh) yb = DADD(yb, DMUL(cyp, yp.y));
                    if (!zh) yb = DADD(yb, DMUL(czp, z1[q].y));
                }
                pair_epilogue<MODE>(ya, yb, wA, wB, q * G.segA + p, c[q].x, c[q].y, vs, a0, a1);
            }
            if (fA) {       // U row of cell a: U columns k-1, j-1, i-1, centre, i+1, j+1, k+1
                double s = suA;
                s = DADD(s, DMUL(cf.msz, uzm.x)); s = DADD(s, DMUL(cf.msy, uym.x)); s = DADD(s, DMUL(cf.msx, uxm));
                s = DADD(s, DMUL(cf.diag_int, uA));
                s = DADD(s, DMUL(cf.msx, uB)); s = DADD(s, DMUL(cf.msy, uyp.x)); s = DADD(s, DMUL(cf.msz, uzp.x));
                row_epilogue<MODE>(s, ub + gA, uA, vs, a0, a1);
            }
            if (fB) {
                double s = suB;
                s = DADD(s, DMUL(cf.msz, uzm.y)); s = DADD(s, DMUL(cf.msy, uym.y)); s = DADD(s, DMUL(cf.msx, uA));
                s = DADD(s, DMUL(cf.diag_int, uB));
                s = DADD(s, DMUL(cf.msx, uxp)); s = DADD(s, DMUL(cf.msy, uyp.y)); s = DADD(s, DMUL(cf.msz, uzp.y));
                row_epilogue<MODE>(s, ub + gA + 1, uB, vs, a0, a1);
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) { m[q] = c[q]; c[q] = z1[q]; z1[q] = z2[q]; }
        }
    }
    if (MODE != MODE_PLAIN) {
        const int pidx = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        const unsigned ex = finalize_here ? expected : 0xffffffffu;
        if (MODE == MODE_AP)
            reduce_epilogue<1>(a0, 0.0, partials, pstride, pidx, ex, ctl.sc, RED_APR0, RED_APR0, sh);
        else if (MODE == MODE_AS)
            reduce_epilogue<2>(a0, a1, partials, pstride, pidx, ex, ctl.sc, RED_ASS, RED_ASAS, sh);
        else
            reduce_epilogue<2>(a0, a1, partials, pstride, pidx, ex, ctl.sc, RED_BB, RED_RR_INIT, sh);
    }
}

// --------------------------------------------------------------------------------------------
// K2 (main path, v3): like k_stencil2_spmv but ONE vector component per thread (blockIdx.z =
// 3*zchunk + comp), so a thread keeps only 4 double2 of marching state and ~64 registers: four
// CTAs per SM, every thread with a 16-byte prefetch plus its neighbour loads in flight.  The three
// A rows of a cell are independent of each other; the Az threads also compute the 13-entry U row
// of interior conductor cells (Ax(i+-1), Ay(j+-1) come through L1/L2).
// --------------------------------------------------------------------------------------------
template <int MODE, int MINB, int DBG = 0>
__global__ void __launch_bounds__(256, MINB)
k_stencil3_spmv(const SlabGeom G, const Coef cf, const MatCoef mc, const unsigned char *__restrict__ cls,
                const int *__restrict__ geo, const VecSet vs, const IterCtl ctl, const int zc, double *partials,
                const int pstride, const unsigned expected, const int finalize_here)
{
    __shared__ double sh[32];
    if (!spmv_guard<MODE>(ctl)) return;
    const int comp = blockIdx.z % 3, zchunk = blockIdx.z / 3;
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2, j = blockIdx.y * blockDim.y + threadIdx.y;
    const int kb = G.k0 + zchunk * zc;
    const int ke = min(kb + zc, G.k1);
    double a0 = 0.0, a1 = 0.0;
    if (i0 < G.sdx && j < G.sdy) {
        const int sdx = G.sdx, kdz = G.kdz, sdz = G.sdz;
        const long long col = (long long)j * sdx + i0;
        const bool xlA = (i0 == 0), xhB = (i0 + 2 == sdx), yl = (j == 0), yh = (j == G.sdy - 1);
        const double cxpA = xlA ? cf.blo[0] : cf.msx;
        const double cxmB = xhB ? cf.bhi[0] : cf.msx;
        const double cym = yh ? cf.bhi[1] : cf.msy, cyp = yl ? cf.blo[1] : cf.msy;
        const int by = (int)(yl | yh) << 1;
        const double *__restrict__ xv = vs.x;
        const long long ub = G.offU - (long long)G.gbase;
        const long long cb = (long long)comp * G.segA;                // this component's segment
        long long p = (long long)(kb - G.k0 + 1) * kdz + col;
        const double g1c = mc.g1[comp];
        const double2 zero2 = make_double2(0.0, 0.0);
        // marching state: planes k-1, k in m, c; planes k+1 .. k+PF in zq[]; plane k+PF+1 is issued at
        // the top of iteration k, so PF 16-byte DRAM loads per thread are always in flight
        constexpr int PF = 1;
        double2 m, c, zq[PF];
        const double *__restrict__ aux = (MODE == MODE_AP) ? vs.r0 : (MODE == MODE_INIT) ? vs.b : nullptr;
        double2 auxc = zero2;
        {
            const double *X = xv + cb + p;
            c = ld2(X);
            m = (kb > 0) ? ld2(X - kdz) : zero2;
#pragma unroll
            for (int d = 0; d < PF; ++d)
                zq[d] = ((d == 0 || kb + d < ke) && (kb + d + 1 < sdz)) ? ld2(X + (long long)(d + 1) * kdz) : zero2;
            if (MODE == MODE_AP || MODE == MODE_INIT) auxc = ld2(aux + cb + p);
        }
        for (int k = kb; k < ke; ++k, p += kdz) {
            const bool zl = (k == 0), zh = (k == sdz - 1);
            const double *X = xv + cb + p;
            const double2 znew = ((k + PF < ke) && (k + PF + 1 < sdz)) ? ld2(X + (long long)(PF + 1) * kdz) : zero2;
            double2 auxn = zero2;
            if ((MODE == MODE_AP || MODE == MODE_INIT) && (k + 1 < ke)) auxn = ld2(aux + cb + p + kdz);
            const double2 z1 = zq[0];
            const uchar2 cl = *reinterpret_cast<const uchar2 *>(cls + (long long)(k - G.k0) * kdz + col);
            const double2 ym = (DBG & 1) ? m : yl ? make_double2(0.0, 0.0) : ld2(X - sdx);
            const double2 yp = (DBG & 1) ? z1 : yh ? make_double2(0.0, 0.0) : ld2(X + sdx);
            const double xm = (DBG & 1) ? c.y : xlA ? 0.0 : X[-1];
            const double xp = (DBG & 1) ? c.x : xhB ? 0.0 : X[2];
            const bool fA = (DBG & 4) ? false : (cl.x == 1), fB = (DBG & 4) ? false : (cl.y == 1);
            const bool wA = (DBG & 2) ? (cl.x == 77) : (cl.x != 2), wB = (DBG & 2) ? (cl.y == 77) : (cl.y != 2);
            double ya = 0.0, yb = 0.0;
            if (DBG & 2) { a0 = DADD(a0, DADD(DADD(m.x, c.y), DADD(z1.x, ym.y))); a1 = DADD(a1, DADD(DADD(yp.x, xm), xp)); }
            if (fA | fB) {
                // interior conductor: both cells of the pair are conductor cells, numbered
                // consecutively along x.  Gather the U values this component's rows need.
                const int *gp = geo + (long long)(k - G.k0 + 2) * kdz + col;
                const int gA = gp[0];
                double umA = 0.0, upA = 0.0, umB = 0.0, upB = 0.0;   // U(-1), U(+1) along this axis
                if (comp == 0) {
                    const double uA = xv[ub + gA], uB = xv[ub + gA + 1];
                    umB = uA; upA = uB;
                    if (fA) umA = xv[ub + gA - 1];
                    if (fB) upB = xv[ub + gA + 2];
                } else {
                    const int st = (comp == 1) ? sdx : kdz;
                    const int2 gm = *reinterpret_cast<const int2 *>(gp - st), gq = *reinterpret_cast<const int2 *>(gp + st);
                    if (fA) { umA = xv[ub + gm.x]; upA = xv[ub + gq.x]; }
                    if (fB) { umB = xv[ub + gm.y]; upB = xv[ub + gq.y]; }
                }
                if (fA) {
                    ya = DADD(0.0, DMUL(mc.cm[2], m.x));
                    ya = DADD(ya, DMUL(mc.cm[1], ym.x));
                    ya = DADD(ya, DMUL(mc.cm[0], xm));
                    ya = DADD(ya, DMUL(mc.diag, c.x));
                    ya = DADD(ya, DMUL(mc.cp[0], c.y));
                    ya = DADD(ya, DMUL(mc.cp[1], yp.x));
                    ya = DADD(ya, DMUL(mc.cp[2], z1.x));
                    ya = DADD(ya, DMUL(g1c, umA));
                    ya = DADD(ya, DMUL(-g1c, upA));
                }
                if (fB) {
                    yb = DADD(0.0, DMUL(mc.cm[2], m.y));
                    yb = DADD(yb, DMUL(mc.cm[1], ym.y));
                    yb = DADD(yb, DMUL(mc.cm[0], c.x));
                    yb = DADD(yb, DMUL(mc.diag, c.y));
                    yb = DADD(yb, DMUL(mc.cp[0], xp));
                    yb = DADD(yb, DMUL(mc.cp[1], yp.y));
                    yb = DADD(yb, DMUL(mc.cp[2], z1.y));
                    yb = DADD(yb, DMUL(g1c, umB));
                    yb = DADD(yb, DMUL(-g1c, upB));
                }
                if (comp == 2) {
                    // U rows (EC3D.f90:917-922): A columns Ax(i-1),Ax(i+1),Ay(j-1),Ay(j+1),Az(k-1),Az(k+1),
                    // then U columns k-1, j-1, i-1, centre, i+1, j+1, k+1.  umA/upA hold U(k-1)/U(k+1).
                    const double *X0 = xv + p, *X1 = xv + G.segA + p;
                    const double2 a0c = ld2(X0);
                    const double2 a1m = ld2(X1 - sdx), a1p = ld2(X1 + sdx);
                    const int2 gjm = *reinterpret_cast<const int2 *>(gp - sdx), gjp = *reinterpret_cast<const int2 *>(gp + sdx);
                    const double uA = xv[ub + gA], uB = xv[ub + gA + 1];
                    if (fA) {
                        double s = DADD(0.0, DMUL(cf.ua_p[0], X0[-1]));
                        s = DADD(s, DMUL(cf.ua_m[0], a0c.y));
                        s = DADD(s, DMUL(cf.ua_p[1], a1m.x));
                        s = DADD(s, DMUL(cf.ua_m[1], a1p.x));
                        s = DADD(s, DMUL(cf.ua_p[2], m.x));
                        s = DADD(s, DMUL(cf.ua_m[2], z1.x));
                        s = DADD(s, DMUL(cf.msz, umA));
                        s = DADD(s, DMUL(cf.msy, xv[ub + gjm.x]));
                        s = DADD(s, DMUL(cf.msx, xv[ub + gA - 1]));
                        s = DADD(s, DMUL(cf.diag_int, uA));
                        s = DADD(s, DMUL(cf.msx, uB));
                        s = DADD(s, DMUL(cf.msy, xv[ub + gjp.x]));
                        s = DADD(s, DMUL(cf.msz, upA));
                        row_epilogue<MODE>(s, ub + gA, uA, vs, a0, a1);
                    }
                    if (fB) {
                        double s = DADD(0.0, DMUL(cf.ua_p[0], a0c.x));
                        s = DADD(s, DMUL(cf.ua_m[0], X0[2]));
                        s = DADD(s, DMUL(cf.ua_p[1], a1m.y));
                        s = DADD(s, DMUL(cf.ua_m[1], a1p.y));
                        s = DADD(s, DMUL(cf.ua_p[2], m.y));
                        s = DADD(s, DMUL(cf.ua_m[2], z1.y));
                        s = DADD(s, DMUL(cf.msz, umB));
                        s = DADD(s, DMUL(cf.msy, xv[ub + gjm.y]));
                        s = DADD(s, DMUL(cf.msx, uA));
                        s = DADD(s, DMUL(cf.diag_int, uB));
                        s = DADD(s, DMUL(cf.msx, xv[ub + gA + 2]));
                        s = DADD(s, DMUL(cf.msy, xv[ub + gjp.y]));
                        s = DADD(s, DMUL(cf.msz, upB));
                        row_epilogue<MODE>(s, ub + gA + 1, uB, vs, a0, a1);
                    }
                }
            }
            if (!fA | !fB) {
                const double czm = zh ? cf.bhi[2] : cf.msz, czp = zl ? cf.blo[2] : cf.msz;
                const int bz = (int)(zl | zh) << 2;
                if (!fA && wA) {
                    const int bm = (int)xlA | by | bz;
                    const double dg = bm ? cf.diag_b[bm] : cf.diag_int;
                    if (!zl) ya = DADD(ya, DMUL(czm, m.x));
                    if (!yl) ya = DADD(ya, DMUL(cym, ym.x));
                    if (!xlA) ya = DADD(ya, DMUL(cf.msx, xm));
                    ya = DADD(ya, DMUL(dg, c.x));
                    ya = DADD(ya, DMUL(cxpA, c.y));
                    if (!yh) ya = DADD(ya, DMUL(cyp, yp.x));
                    if (!zh) ya = DADD(ya, DMUL(czp, z1.x));
                }
                if (!fB && wB) {
                    const int bm = (int)xhB | by | bz;
                    const double dg = bm ? cf.diag_b[bm] : cf.diag_int;
                    if (!zl) yb = DADD(yb, DMUL(czm, m.y));
                    if (!yl) yb = DADD(yb, DMUL(cym, ym.y));
                    yb = DADD(yb, DMUL(cxmB, c.x));
                    yb = DADD(yb, DMUL(dg, c.y));
                    if (!xhB) yb = DADD(yb, DMUL(cf.msx, xp));
                    if (!yh) yb = DADD(yb, DMUL(cyp, yp.y));
                    if (!zh) yb = DADD(yb, DMUL(czp, z1.y));
                }
            }
            pair_epilogue_pre<MODE>(ya, yb, wA, wB, cb + p, c.x, c.y, auxc, vs, a0, a1);
            m = c; c = z1;
#pragma unroll
            for (int d = 0; d + 1 < PF; ++d) zq[d] = zq[d + 1];
            zq[PF - 1] = znew;
            auxc = auxn;
        }
    }
    if (MODE != MODE_PLAIN) {
        const int pidx = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        const unsigned ex = finalize_here ? expected : 0xffffffffu;
        if (MODE == MODE_AP)
            reduce_epilogue<1>(a0, 0.0, partials, pstride, pidx, ex, ctl.sc, RED_APR0, RED_APR0, sh);
        else if (MODE == MODE_AS)
            reduce_epilogue<2>(a0, a1, partials, pstride, pidx, ex, ctl.sc, RED_ASS, RED_ASAS, sh);
        else
            reduce_epilogue<2>(a0, a1, partials, pstride, pidx, ex, ctl.sc, RED_BB, RED_RR_INIT, sh);
    }
}

// --------------------------------------------------------------------------------------------
// K2b: matrix-free SpMV, conductor cells: three A rows with convection, 2C/dt and grad-U coupling
// (EC3D.f90:656-710) plus the U row (EC3D.f90:766-922).  One thread per owned conductor cell.
// --------------------------------------------------------------------------------------------
struct GatherVisitor {
    const double *__restrict__ x;
    long long segA, offU, cell_shift;   // local A index = comp*segA + cell0 - cell_shift
    int gbase;
    double s;
    __device__ __forceinline__ void a(int comp, long long cell0, double coef)
    {
        s = DADD(s, DMUL(coef, x[comp * segA + (cell0 - cell_shift)]));
    }
    __device__ __forceinline__ void u(int g, double coef) { s = DADD(s, DMUL(coef, x[offU + (g - gbase)])); }
};

template <int MODE>
__global__ void __launch_bounds__(256)
k_cond_spmv(const SlabGeom G, const Coef cf, const MatCoef *__restrict__ mcs, const int *__restrict__ geo,
            const signed char *__restrict__ mat, const int *__restrict__ cond_cells, const int ncond,
            const VecSet vs, const IterCtl ctl, double *partials, const int pstride, const int pbase,
            const unsigned expected)
{
    __shared__ double sh[32];
    if (!spmv_guard<MODE>(ctl)) return;
    double a0 = 0.0, a1 = 0.0;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ncond) {
        const int cell0 = cond_cells[t];
        const int k = cell0 / G.kdz, rem = cell0 - k * G.kdz, j = rem / G.sdx, i = rem - j * G.sdx;
        GeoView gv{geo, G.sdx, G.sdy, G.kdz, G.k0 - 2, G.nzl + 4};
        const long long cell_shift = (long long)(G.k0 - 1) * G.kdz;
        const long long lp = (long long)cell0 - cell_shift;
        const long long lmat = (long long)cell0 - (long long)(G.k0 - 2) * G.kdz;
        const MatCoef mc = mcs[mat[lmat] - 1];
        GatherVisitor v{vs.x, G.segA, G.offU, cell_shift, G.gbase, 0.0};
#pragma unroll
        for (int comp = 0; comp < 3; ++comp) {
            v.s = 0.0;
            cond_a_row(mc, gv, i, j, k, comp, v);
            row_epilogue<MODE>(v.s, comp * G.segA + lp, vs.x[comp * G.segA + lp], vs, a0, a1);
        }
        v.s = 0.0;
        cond_u_row(cf, gv, i, j, k, 3, v);
        const long long lu = G.offU + (gv.at(i, j, k) - G.gbase);
        row_epilogue<MODE>(v.s, lu, vs.x[lu], vs, a0, a1);
    }
    if (MODE != MODE_PLAIN) {
        const int pidx = pbase + blockIdx.x;
        if (MODE == MODE_AP)
            reduce_epilogue<1>(a0, 0.0, partials, pstride, pidx, expected, ctl.sc, RED_APR0, RED_APR0, sh);
        else if (MODE == MODE_AS)
            reduce_epilogue<2>(a0, a1, partials, pstride, pidx, expected, ctl.sc, RED_ASS, RED_ASAS, sh);
        else
            reduce_epilogue<2>(a0, a1, partials, pstride, pidx, expected, ctl.sc, RED_BB, RED_RR_INIT, sh);
    }
}

// --------------------------------------------------------------------------------------------
// K2': CSR SpMV on the reference's own arrays (drop-in mode), one thread per row, entries summed
// sequentially in stored order like sprsAx (solvers.f90:54-61).
// --------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
k_csr_spmv(const int n, const int *__restrict__ irow, const int *__restrict__ jcol, const double *__restrict__ valA,
           const VecSet vs, const IterCtl ctl, double *partials, const int pstride, const unsigned expected)
{
    __shared__ double sh[32];
    if (!spmv_guard<MODE>(ctl)) return;
    double a0 = 0.0, a1 = 0.0;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) {
        const int i1 = irow[r] - 1, i2 = irow[r + 1] - 1;
        double s = 0.0;
        for (int m = i1; m < i2; ++m) s = DADD(s, DMUL(__ldg(valA + m), vs.x[__ldg(jcol + m) - 1]));
        row_epilogue<MODE>(s, r, vs.x[r], vs, a0, a1);
    }
    if (MODE != MODE_PLAIN) {
        if (MODE == MODE_AP)
            reduce_epilogue<1>(a0, 0.0, partials, pstride, blockIdx.x, expected, ctl.sc, RED_APR0, RED_APR0, sh);
        else if (MODE == MODE_AS)
            reduce_epilogue<2>(a0, a1, partials, pstride, blockIdx.x, expected, ctl.sc, RED_ASS, RED_ASAS, sh);
        else
            reduce_epilogue<2>(a0, a1, partials, pstride, blockIdx.x, expected, ctl.sc, RED_BB, RED_RR_INIT, sh);
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

// K3: alpha = rr0/(AP,R0); S = R - alpha*AP; ||S||^2.            solvers.f90:31-34
template <int VEC>
__global__ void __launch_bounds__(256)
k_s_update(const SlabGeom G, const double *__restrict__ R, const double *__restrict__ AP, double *__restrict__ S,
           const IterCtl ctl, double *partials, const int pstride, const unsigned expected)
{
    __shared__ double sh[32];
    Scal *sc = ctl.sc;
    if (sc->done) return;
    const int it = *ctl.iter_base + ctl.it_off;
    const double rr0 = (it == 1) ? sc->red[RED_RR_INIT] : sc->rr0[it & 1];
    const double alpha = rr0 / sc->red[RED_APR0];
    if (is_block0()) sc->alpha = alpha;
    double acc = 0.0;
    const long long units = G.n_own / VEC;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < units;
         q += (long long)gridDim.x * blockDim.x) {
        const long long l = own_to_local(G, q * VEC);
        if (VEC == 2) {
            const double2 r = *reinterpret_cast<const double2 *>(R + l);
            const double2 ap = *reinterpret_cast<const double2 *>(AP + l);
            double2 s;
            s.x = DSUB(r.x, DMUL(alpha, ap.x));
            s.y = DSUB(r.y, DMUL(alpha, ap.y));
            *reinterpret_cast<double2 *>(S + l) = s;
            acc = DADD(acc, DMUL(s.x, s.x));
            acc = DADD(acc, DMUL(s.y, s.y));
        } else {
            const double s = DSUB(R[l], DMUL(alpha, AP[l]));
            S[l] = s;
            acc = DADD(acc, DMUL(s, s));
        }
    }
    reduce_epilogue<1>(acc, 0.0, partials, pstride, blockIdx.x, expected, sc, RED_SS, RED_SS, sh);
}

// K5: if ||S|| converged: X = X + alpha*P (solvers.f90:36).  Else omega = (AS,S)/(AS,AS);
// X = X + alpha*P + omega*S; R = S - omega*AS; ||R||^2, (R,R0).   solvers.f90:40-44
template <int VEC>
__global__ void __launch_bounds__(256)
k_xr_update(const SlabGeom G, double *__restrict__ X, const double *__restrict__ P, const double *__restrict__ S,
            const double *__restrict__ AS, double *__restrict__ R, const double *__restrict__ R0,
            const IterCtl ctl, double *partials, const int pstride, const unsigned expected)
{
    __shared__ double sh[32];
    Scal *sc = ctl.sc;
    if (sc->done) return;
    const double alpha = sc->alpha;
    const long long units = G.n_own / VEC;
    const long long q0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long qs = (long long)gridDim.x * blockDim.x;
    if (s_converged(sc)) {
        for (long long q = q0; q < units; q += qs) {
            const long long l = own_to_local(G, q * VEC);
#pragma unroll
            for (int v = 0; v < VEC; ++v) X[l + v] = DADD(X[l + v], DMUL(alpha, P[l + v]));
        }
        return;
    }
    const double omega = sc->red[RED_ASS] / sc->red[RED_ASAS];
    if (is_block0()) sc->omega = omega;
    double a0 = 0.0, a1 = 0.0;
    for (long long q = q0; q < units; q += qs) {
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
            a0 = DADD(a0, DMUL(r.x, r.x)); a0 = DADD(a0, DMUL(r.y, r.y));
            a1 = DADD(a1, DMUL(r.x, r0.x)); a1 = DADD(a1, DMUL(r.y, r0.y));
        } else {
            const double s = S[l];
            X[l] = DADD(DADD(X[l], DMUL(alpha, P[l])), DMUL(omega, s));
            const double r = DSUB(s, DMUL(omega, AS[l]));
            R[l] = r;
            a0 = DADD(a0, DMUL(r, r));
            a1 = DADD(a1, DMUL(r, R0[l]));
        }
    }
    reduce_epilogue<2>(a0, a1, partials, pstride, blockIdx.x, expected, sc, RED_RR, RED_RR0N, sh);
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
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < units;
         q += (long long)gridDim.x * blockDim.x) {
        const long long l = own_to_local(G, q * VEC);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const double r = R[l + v];
            if (restart) {
                R0[l + v] = r;
                P[l + v] = r;
            } else {
                P[l + v] = DADD(r, DMUL(beta, DSUB(P[l + v], DMUL(omega, AP[l + v]))));
            }
        }
    }
}

__global__ void k_solver_reset(Scal *sc, int *iter_base, double tol, int itmax)
{
    for (int q = 0; q < 8; ++q) sc->red[q] = 0.0;
    sc->rr0[0] = sc->rr0[1] = 0.0;
    sc->alpha = sc->omega = sc->beta = 0.0;
    sc->tol = tol; sc->itmax = itmax;
    sc->done = 0; sc->final_iter = 0; sc->exit_kind = 0; sc->restarts = 0; sc->counter = 0u;
    *iter_base = 0;
}

__global__ void k_iter_advance(int *iter_base, int by) { *iter_base += by; }

// generic helpers
__global__ void k_fill(double *p, long long n, double v)
{
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
        p[q] = v;
}
