// ec3d_tma.cuh -- K2 (main path): TMA-staged matrix-free SpMV of the coupled A-U operator, sm_100a.
//
// Why TMA: a register-marching stencil is latency bound -- every plane needs fresh cache lines and a
// thread keeps only one or two 16-byte loads in flight (ncu on the first version: long-scoreboard
// stalls, 24 % warps active, DRAM at 30 %).  Here ONE elected thread per CTA posts
// `cp.async.bulk.tensor` copies of whole (x,y) tiles -- with their halo, zero-filled outside the
// domain by the hardware -- into a ring of shared-memory stages several planes ahead; each stage is
// guarded by an mbarrier that the copy engine completes (complete_tx).  The 256 compute threads only
// touch shared memory and registers, so the bytes in flight per SM are set by the ring depth
// (NSTAGE-2 stages x 16-22 KB x 2 CTAs), not by registers.
//
// One CTA = a 64 x 8 (x,y) tile, all three vector components and the U block, marching over `zc`
// planes; each thread owns two x-adjacent cells (16-byte shared loads / global stores) and keeps the
// k-1 / k / k+1 values of Ax, Ay, Az and U of its pair in registers.  Per plane the ring delivers
//     xs : 68 x 10 doubles x 3 components  (one 4-D TMA box of the input vector's A part)
//     us : 68 x 10 doubles                 (3-D TMA box of the dense U box; only for tiles that
//                                           touch the conductor's bounding box)
//     cs : 64 x 8 class bytes              (3-D TMA box of the class map)
// and every row of every cell of the tile is produced in that one pass:
//     * A rows of air / domain-face cells ........ EC3D.f90:528-654
//     * A rows of conductor cells ................ EC3D.f90:656-710 (convection, 2C/dt, grad U:
//                                                  central or one-sided, chosen by the class byte)
//     * U rows of conductor cells ................ EC3D.f90:766-922 (interior 13-entry rows and the
//                                                  26 corner / edge / face cases incl. the :803-807
//                                                  sign anomaly)
// The per-cell class byte (k_build_cls2) is 0 for non-conductor cells, else 0x40 | sx | sy<<2 | sz<<4
// with s_axis bit0 = '-' neighbour is not a conductor, bit1 = '+' neighbour is not: everything the
// row rules branch on, so geoPHYS_C is not read here at all.  One-sided gradients reach two cells:
// U(i+-2) is inside the tile (2 halo columns), U(j+-2) and U(k+-2) are read straight from the dense U
// box (only surface cells do that).
// Rows are summed in ascending column order with unfused mul/add starting from 0.0, i.e. exactly the
// sequential CSR row sum of the reference's sprsAx (solvers.f90:54-61) on the assembled matrix.
// The dot products of BiCGSTABwr that follow the SpMV are fused in (MODE_AP / MODE_AS / MODE_INIT).
#pragma once
#include <cuda.h>

#include "ec3d_kernels.cuh"

namespace tma {

constexpr int TX = 64, TY = 8;              // cells per plane tile (2 cells per thread, 256 threads)
constexpr int BW = TX + 4, BH = TY + 2;     // halo box: 2 extra columns each side keep the own pair 16-byte aligned
constexpr int TILE_D = BW * BH;             // 680 doubles
constexpr int TILE_BYTES = TILE_D * 8;      // 5440
constexpr int XS_BYTES = 3 * TILE_BYTES;    // 16320: one TMA box 68 x 10 x 1 x 3
constexpr int CLS_BYTES = TX * TY;          // 512: class-byte tile (64 x 8 bytes)

// U values of one cell along one axis at offsets -2 .. +2
struct U5 {
    double m2, m1, c0, p1, p2;
};

// A row (component / axis AX) of a conductor cell: 7 A columns (EC3D.f90:649-663) then the grad-U
// columns (:667-710) -- backward one-sided when the '+' neighbour is missing, forward one-sided when
// only the '-' neighbour is, central otherwise.  sa = that axis' 2-bit state.
template <int AX>
__device__ __forceinline__ double cond_a(const MatCoef &mc, const int sa, const double zm, const double ym, const double xm,
                                         const double cc, const double xp, const double yp, const double zp, const U5 &u)
{
    double y = DADD(0.0, DMUL(mc.cm[2], zm));
    y = DADD(y, DMUL(mc.cm[1], ym));
    y = DADD(y, DMUL(mc.cm[0], xm));
    y = DADD(y, DMUL(mc.diag, cc));
    y = DADD(y, DMUL(mc.cp[0], xp));
    y = DADD(y, DMUL(mc.cp[1], yp));
    y = DADD(y, DMUL(mc.cp[2], zp));
    if (sa & 2) {
        y = DADD(y, DMUL(-mc.g1[AX], u.m2));
        y = DADD(y, DMUL(mc.g4[AX], u.m1));
        y = DADD(y, DMUL(-mc.g3[AX], u.c0));
    } else if (sa & 1) {
        y = DADD(y, DMUL(mc.g3[AX], u.c0));
        y = DADD(y, DMUL(-mc.g4[AX], u.p1));
        y = DADD(y, DMUL(mc.g1[AX], u.p2));
    } else {
        y = DADD(y, DMUL(mc.g1[AX], u.m1));
        y = DADD(y, DMUL(-mc.g1[AX], u.p1));
    }
    return y;
}

// A-column part of a U row, contribution of component AX, added to the running sum in ascending
// column order (interior: A(-1), A(+1) of that axis, EC3D.f90:917-922; surface: the same cell's A
// with -+2/(dt*d), only for axes with a missing neighbour, :773-916, anomaly at :803-807).
template <int AX>
__device__ __forceinline__ double urow_a(const Coef &cf, const int st, const double am, const double ac, const double ap,
                                         double s)
{
    const int sx = st & 3, sy = (st >> 2) & 3, sz = (st >> 4) & 3;
    if ((st & 63) == 0) {
        s = DADD(s, DMUL(cf.ua_p[AX], am));
        s = DADD(s, DMUL(cf.ua_m[AX], ap));
        return s;
    }
    const bool anomaly = (sx == 1 && sy == 2 && sz == 2);
    const int sa = (AX == 0) ? sx : (AX == 1) ? sy : sz;
    if (sa) {
        double coef = (sa == 1) ? cf.uc_m[AX] : cf.uc_p[AX];
        if (anomaly && AX == 0) coef = cf.uc_p[0];
        if (anomaly && AX == 1) coef = cf.uc_m[1];
        s = DADD(s, DMUL(coef, ac));
    }
    return s;
}

// U-column part of a U row: k-1, j-1, i-1, centre, i+1, j+1, k+1 (ascending U numbers).
__device__ __forceinline__ double urow_u(const Coef &cf, const int st, const double ukm, const double ujm, const double uim,
                                         const double uc, const double uip, const double ujp, const double ukp, double s)
{
    const int sx = st & 3, sy = (st >> 2) & 3, sz = (st >> 4) & 3;
    if (sz == 0) s = DADD(s, DMUL(cf.msz, ukm)); else if (sz == 2) s = DADD(s, DMUL(cf.m2s[2], ukm));
    if (sy == 0) s = DADD(s, DMUL(cf.msy, ujm)); else if (sy == 2) s = DADD(s, DMUL(cf.m2s[1], ujm));
    if (sx == 0) s = DADD(s, DMUL(cf.msx, uim)); else if (sx == 2) s = DADD(s, DMUL(cf.m2s[0], uim));
    s = DADD(s, DMUL(cf.diag_int, uc));
    if (sx == 0) s = DADD(s, DMUL(cf.msx, uip)); else if (sx == 1) s = DADD(s, DMUL(cf.m2s[0], uip));
    if (sy == 0) s = DADD(s, DMUL(cf.msy, ujp)); else if (sy == 1) s = DADD(s, DMUL(cf.m2s[1], ujp));
    if (sz == 0) s = DADD(s, DMUL(cf.msz, ukp)); else if (sz == 1) s = DADD(s, DMUL(cf.m2s[2], ukp));
    return s;
}

}  // namespace tma

// Class byte of the TMA SpMV for every owned cell (layout [k-k0][j][i], row pitch clsx = sdx rounded
// up to 16 so the map is a legal TMA tensor).
__global__ void k_build_cls2(const SlabGeom G, const int *__restrict__ geo, const int *__restrict__ cond_cells,
                             const int ncond, unsigned char *__restrict__ cls, const int clsx)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncond) return;
    const int cell0 = cond_cells[t];
    const int k = cell0 / G.kdz, rem = cell0 - k * G.kdz, j = rem / G.sdx, i = rem - j * G.sdx;
    GeoView gv{geo, G.sdx, G.sdy, G.kdz, G.k0 - 2, G.nzl + 4};
    const int sx = (gv.at(i - 1, j, k) == 0) | ((gv.at(i + 1, j, k) == 0) << 1);
    const int sy = (gv.at(i, j - 1, k) == 0) | ((gv.at(i, j + 1, k) == 0) << 1);
    const int sz = (gv.at(i, j, k - 1) == 0) | ((gv.at(i, j, k + 1) == 0) << 1);
    cls[((long long)(k - G.k0) * G.sdy + j) * clsx + i] = (unsigned char)(0x40 | sx | (sy << 2) | (sz << 4));
}

// Row epilogue of a cell pair: stores, and the fused BiCGSTABwr dot products.  The products of the
// dots are accumulated with fma (the reference's sequential dot_product order is not reproducible
// in parallel anyway; fewer roundings, half the FP64 instructions).  xa, xb = the SpMV input at the
// two cells (MODE_SAS: s = r - alpha*Ap, which this kernel also materialises in vs.S).
// Several ranks (px.on): rows of a slab-boundary plane that are inputs of a later SpMV (Ap; R and P of the
// initial residual) are also stored into the neighbour's halo slot (ec3d_comm.cuh) -- seg / o = owned
// segment and offset of the pair inside it.
struct PushCtx {
    const PeerTable *pt;
    const SlabGeom *G;
    bool on;                // this plane is a boundary plane of the slab (block uniform)
};
__device__ __forceinline__ void push_rows(const PushCtx &px, const int vi, const int seg, const long long o, const double a,
                                          const double b, const bool wa, const bool wb)
{
    if (wa && wb) peer_push2(*px.pt, *px.G, vi, seg, o, a, b);
    else {
        if (wa) peer_push1(*px.pt, *px.G, vi, seg, o, a);
        if (wb) peer_push1(*px.pt, *px.G, vi, seg, o + 1, b);
    }
}

// ja / jb: optional Jacobi scaling (ec3d_set_preconditioner, OFF by default): the rows of D^-1 A, i.e.
// every row result -- and the right-hand side in MODE_INIT -- times 1/diagonal; 0.0 = off (exact reference rows).
template <int MODE>
__device__ __forceinline__ void pair_out(double ya, double yb, const bool wa, const bool wb, const long long idx,
                                         const double xa, const double xb, double2 aux, const VecSet &vs,
                                         double &a0, double &a1, double &a2, const PushCtx &px, const int seg, const long long o,
                                         const double ja = 0.0, const double jb = 0.0)
{
    if (ja != 0.0) {
        ya = DMUL(ja, ya); yb = DMUL(jb, yb);
        if (MODE == MODE_INIT) { aux.x = DMUL(ja, aux.x); aux.y = DMUL(jb, aux.y); }
    }
    if (MODE == MODE_INIT) {
        const double ra = DSUB(aux.x, ya), rb = DSUB(aux.y, yb);
        if (wa && wb) { st2(vs.R + idx, ra, rb); st2(vs.R0 + idx, ra, rb); st2(vs.P + idx, ra, rb); }
        else {
            if (wa) { vs.R[idx] = ra; vs.R0[idx] = ra; vs.P[idx] = ra; }
            if (wb) { vs.R[idx + 1] = rb; vs.R0[idx + 1] = rb; vs.P[idx + 1] = rb; }
        }
        if (px.on) { push_rows(px, vs.vi_R, seg, o, ra, rb, wa, wb); push_rows(px, vs.vi_P, seg, o, ra, rb, wa, wb); }
        if (wa) { a0 = __fma_rn(aux.x, aux.x, a0); a1 = __fma_rn(ra, ra, a1); }
        if (wb) { a0 = __fma_rn(aux.y, aux.y, a0); a1 = __fma_rn(rb, rb, a1); }
        return;
    }
    if (wa && wb) {
        st2(vs.y + idx, ya, yb);
        if (MODE == MODE_SAS) st2(vs.S + idx, xa, xb);
    } else {
        if (wa) { vs.y[idx] = ya; if (MODE == MODE_SAS) vs.S[idx] = xa; }
        if (wb) { vs.y[idx + 1] = yb; if (MODE == MODE_SAS) vs.S[idx + 1] = xb; }
    }
    if (MODE == MODE_AP && px.on) push_rows(px, vs.vi_y, seg, o, ya, yb, wa, wb);
    if (MODE == MODE_AP) {
        if (wa) a0 = __fma_rn(ya, aux.x, a0);
        if (wb) a0 = __fma_rn(yb, aux.y, a0);
    } else if (MODE == MODE_AS || MODE == MODE_SAS) {
        if (wa) { a0 = __fma_rn(ya, xa, a0); a1 = __fma_rn(ya, ya, a1); if (MODE == MODE_SAS) a2 = __fma_rn(xa, xa, a2); }
        if (wb) { a0 = __fma_rn(yb, xb, a0); a1 = __fma_rn(yb, yb, a1); if (MODE == MODE_SAS) a2 = __fma_rn(xb, xb, a2); }
    }
}

// 7-point row with per-position coefficients.  Neighbours that do not exist (domain faces) arrive
// as +0.0 from the TMA zero fill / the zero halo planes, and coef * 0 = +-0 added to a partial sum
// that is never -0.0 (it starts from +0.0) leaves every bit unchanged -- so the row equals the
// reference's shorter boundary row (EC3D.f90:528-646) without a single branch.
__device__ __forceinline__ double row7(const double czm, const double cym, const double cxm, const double dg, const double cxp,
                                       const double cyp, const double czp, const double zm, const double ym,
                                       const double xm, const double cc, const double xp, const double yp,
                                       const double zp)
{
    double y = __fma_rn(czm, zm, 0.0);          // == 0.0 + czm*zm, one rounding either way
    y = DADD(y, DMUL(cym, ym));
    y = DADD(y, DMUL(cxm, xm));
    y = DADD(y, DMUL(dg, cc));
    y = DADD(y, DMUL(cxp, xp));
    y = DADD(y, DMUL(cyp, yp));
    y = DADD(y, DMUL(czp, zp));
    return y;
}

// One work item of the TMA SpMV: a 64 x 8 tile column and a z range.  has_u != 0 when the item
// contains conductor cells (then U tiles and class bytes are staged and the generic row code runs);
// items without conductor cells run a lean 7-point loop.  The two kinds are separate kernel
// instantiations (different register budgets and ring shapes) launched back to back.
struct WorkItem {
    int x0, y0, kb, ke;     // tile origin, owned planes [kb, ke)
    int has_u, pad0, pad1, pad2;
};

// Shared-memory stage of the ring.  NIN input tiles (MODE_SAS stages r AND Ap and forms
// s = r - alpha*Ap on the fly -- solvers.f90:33 fused into the SpMV of :39), each 68 x 10 x 3
// doubles in a 16 KiB slot; conductor items add the class bytes and NIN U tiles.
template <int MODE, bool HAS_U>
struct Stage {
    static constexpr int NIN = (MODE == MODE_SAS) ? 2 : 1;
    static constexpr int A_SLOT = 16384;                                  // >= XS_BYTES, 128-byte multiple
    static constexpr int U_SLOT = 5504;                                   // >= TILE_BYTES, 128-byte multiple
    static constexpr int CLS_OFF = NIN * A_SLOT;
    static constexpr int US_OFF = CLS_OFF + (HAS_U ? tma::CLS_BYTES : 0);
    static constexpr int BYTES = US_OFF + (HAS_U ? NIN * U_SLOT : 0);
    static constexpr int A2 = A_SLOT / 8, U2 = U_SLOT / 8;                // second input, offset in doubles
};
template <int MODE, bool HAS_U>
constexpr int tma_smem_bytes(int nstage) { return nstage * Stage<MODE, HAS_U>::BYTES + 128; }

struct TmaMaps {            // tensor maps of one launch (by value in kernel parameter space)
    CUtensorMap xA, xU;     // SpMV input: A part (4-D), dense U box (3-D)
    CUtensorMap x2A, x2U;   // MODE_SAS: second input (Ap)
    CUtensorMap cls;        // class bytes
    CUtensorMap auxA, auxU; // r0 / b: halo-free boxes, L2 prefetch only
};

struct TmaCtx {             // per-thread constants of k_spmv_tma
    int tx, ty, lane, warp;
    int x0, y0, kb, ke, nload;
    int o_own, o_cls;
    bool active, inUxy;
    double cxpA, cxmB, cym, cyp, dgA0, dgA1, dgB0, dgB1;
    double alpha;           // MODE_SAS
};

// value(s) of the SpMV input at smem position p: the staged vector itself, or (MODE_SAS)
// s = r - alpha*Ap from the two staged tiles -- same expression, hence same bits, as the s this
// kernel stores for the owner of that cell
template <bool SAS>
__device__ __forceinline__ double in1(const double *p, const int d2, const double alpha)
{
    return SAS ? DSUB(p[0], DMUL(alpha, p[d2])) : p[0];
}
template <bool SAS>
__device__ __forceinline__ double2 in2(const double *p, const int d2, const double alpha)
{
    const double2 r = *reinterpret_cast<const double2 *>(p);
    if (!SAS) return r;
    const double2 q = *reinterpret_cast<const double2 *>(p + d2);
    return make_double2(DSUB(r.x, DMUL(alpha, q.x)), DSUB(r.y, DMUL(alpha, q.y)));
}

// Plane loop of one work item.  HAS_U = item contains conductor cells.
template <int MODE, int NSTAGE, bool HAS_U, bool JAC>
__device__ __forceinline__ void tma_plane_loop(const TmaMaps &tm, const SlabGeom &G, const Coef &cf, const MatCoef &mc,
                                               const VecSet &vs, const TmaCtx &t, unsigned char *smem,
                                               unsigned long long *full, unsigned *cnt, dd &acc0, dd &acc1, dd &acc2,
                                               const PeerTable *pt, const bool xf)
{
    using namespace tma;
    using ST = Stage<MODE, HAS_U>;
    constexpr bool SAS = (MODE == MODE_SAS);
    constexpr bool HAS_AUX = (MODE == MODE_AP || MODE == MODE_INIT);
    constexpr int A2 = ST::A2, U2 = ST::U2, USD = ST::US_OFF / 8;
    const int ukA = G.ub_k0 - 1, ukB = G.ub_k0 + G.ub_nz;   // U tiles are needed for planes [ukA, ukB]
    const int kb = t.kb, ke = t.ke, nload = t.nload, x0 = t.x0, y0 = t.y0;
    const int sdx = G.sdx, sdz = G.sdz, kdz = G.kdz;
    const int i0 = x0 + 2 * t.tx, j = y0 + t.ty;
    const int o_own = t.o_own;
    const double alpha = t.alpha;

    auto issue = [&](int q) {                               // q-th plane of this item: pl = kb-1+q
        const int pl = kb - 1 + q, s = q % NSTAGE;
        unsigned char *st = smem + s * ST::BYTES;
        const bool u_ = HAS_U && pl >= ukA && pl <= ukB;
        const bool c_ = pl >= kb && pl < ke;                // planes that are computed
        mbar_expect_tx(full + s, ST::NIN * (XS_BYTES + (u_ ? TILE_BYTES : 0)) + ((HAS_U && c_) ? CLS_BYTES : 0));
        load4d(st, &tm.xA, full + s, x0 - 2, y0 - 1, pl - G.k0 + 1, 0);
        if (SAS) load4d(st + ST::A_SLOT, &tm.x2A, full + s, x0 - 2, y0 - 1, pl - G.k0 + 1, 0);
        if (u_) {
            load3d(st + ST::US_OFF, &tm.xU, full + s, x0 - 2 - G.ub_i0, y0 - 1 - G.ub_j0, pl - G.ub_kl0);
            if (SAS) load3d(st + ST::US_OFF + ST::U_SLOT, &tm.x2U, full + s, x0 - 2 - G.ub_i0, y0 - 1 - G.ub_j0, pl - G.ub_kl0);
        }
        if (HAS_U && c_) load3d(st + ST::CLS_OFF, &tm.cls, full + s, x0, y0, pl - G.k0);
        if (HAS_AUX && c_) {                                // r0 / b of that plane: pull into L2 ahead of the LDGs
            prefetch4d(&tm.auxA, x0, y0, pl - G.k0 + 1, 0);     // 64 x 8 x 1 x 3 box, no halo
            if (u_ && pl > ukA && pl < ukB) prefetch3d(&tm.auxU, x0 - G.ub_i0, y0 - G.ub_j0, pl - G.ub_kl0);
        }
    };
    if (t.warp == 0 && t.lane == 0)
        for (int q = 0; q < min(NSTAGE, nload); ++q) issue(q);

    // A warp is done with the tiles of load ql: the last of the 8 warps to say so re-arms that stage
    // with load ql + NSTAGE.  No CTA-wide barrier: warps drift apart by up to the ring depth.  The
    // acq_rel counter orders the other warps' (generic-proxy) reads of the stage before the issuing
    // thread, and the proxy fence orders them before the copy engine's (async-proxy) refill.
    // Invariant: ONE work item per CTA -- barrier phases and counters are never reset.
    auto release = [&](int ql) {
        __syncwarp();
        if (t.lane == 0) {
            const int s = ql % NSTAGE;
            unsigned old;
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(cnt + s)) : "memory");
            if (old == 7u) {
                cnt[s] = 0u;
                if (ql + NSTAGE < nload) { fence_proxy_async(); issue(ql + NSTAGE); }
            }
        }
    };

    const double *__restrict__ auxp = (MODE == MODE_AP) ? vs.r0 : (MODE == MODE_INIT) ? vs.b : nullptr;
    // offsets of the pair at the plane computed next (k = kb first)
    long long pA = (long long)(kb - G.k0 + 1) * kdz + (long long)j * sdx + i0;
    long long pU = G.offU + (long long)(kb - G.ub_kl0) * G.ub_pl + (long long)(j - G.ub_j0) * G.ub_nx + (i0 - G.ub_i0);
    const double2 zero2 = make_double2(0.0, 0.0);
    double2 pl3[3][3];                                      // [slot][component]: planes k-1, k, k+1 rotate through the slots
    double2 ug3[3] = {zero2, zero2, zero2};                 // same for U
    int s = 0;                                              // stage of load q
    uint32_t ph = 0;

    auto own_pair = [&](int slot, int q) {                  // wait for load q, read the own pair into `slot`
        mbar_wait(full + s, ph);
        const double *xn = reinterpret_cast<const double *>(smem + s * ST::BYTES);
#pragma unroll
        for (int a = 0; a < 3; ++a) pl3[slot][a] = in2<SAS>(xn + a * TILE_D + o_own, A2, alpha);
        if (HAS_U) {
            const int pl = kb - 1 + q;
            ug3[slot] = (pl >= ukA && pl <= ukB) ? in2<SAS>(xn + USD + o_own, U2, alpha) : zero2;
        }
    };
    auto advance = [&]() { if (++s == NSTAGE) { s = 0; ph ^= 1u; } };
    // U of the dense box straight from global memory (one-sided gradients reach two cells)
    auto uglob = [&](long long p) -> double {
        return SAS ? DSUB(vs.x[p], DMUL(alpha, vs.x2[p])) : vs.x[p];
    };

    // loads 0 and 1: planes kb-1, kb
    own_pair(0, 0); advance();
    own_pair(1, 1); release(0); advance();

    const int nz = ke - kb;
    for (int it = 0; it < nz; it += 3) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            if (it + u < nz) {
                const int q = it + u + 2;
                const int k = kb + it + u;                   // plane computed now
                const int sm_ = u % 3, sc_ = (u + 1) % 3, sz_ = (u + 2) % 3;   // slots of planes k-1, k, k+1
                double2 aux[3] = {zero2, zero2, zero2}, auxU = zero2;
                if (HAS_AUX && t.active) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) aux[a] = ld2(auxp + a * G.segA + pA);
                    if (HAS_U && t.inUxy && k >= G.ub_k0 && k < ukB) auxU = ld2(auxp + pU);
                }
                const int sprev = (s == 0) ? NSTAGE - 1 : s - 1;
                // this thread's dot-product terms of plane k: summed in a fixed order here, accumulated
                // across planes in double-double (independent of how z is cut into items and slabs)
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                // fused halo push: planes k0 / k1-1 (A rows) and the two planes nearest each slab face (U rows)
                const bool pushes = xf && (MODE == MODE_AP || MODE == MODE_INIT);
                const PushCtx pxA{pt, &G, pushes && (k == G.k0 || k == G.k1 - 1)};
                const PushCtx pxU{pt, &G, pushes && (k < G.k0 + 2 || k >= G.k1 - 2)};
                const long long oA = pA - kdz, oU = pU - G.own_off[3];       // offsets inside the owned segments
                own_pair(sz_, q);
                if (t.active) {
                    const unsigned char *stp = smem + sprev * ST::BYTES;
                    const double *xs = reinterpret_cast<const double *>(stp);
                    const bool zl = (k == 0), zh = (k == sdz - 1);
                    const double czm = zh ? cf.bhi[2] : cf.msz, czp = zl ? cf.blo[2] : cf.msz;
                    const double dgA = (zl | zh) ? t.dgA1 : t.dgA0, dgB = (zl | zh) ? t.dgB1 : t.dgB0;
                    // Jacobi (off by default; the divisions cost nothing then): 1/diagonal of the air rows of
                    // the pair, of conductor A rows and of U rows
                    double jA = 0.0, jB = 0.0, jC = 0.0, jU = 0.0;
                    if (JAC) { jA = 1.0 / dgA; jB = 1.0 / dgB; jC = 1.0 / mc.diag; jU = 1.0 / cf.diag_int; }
                    const double2 *m = pl3[sm_], *c = pl3[sc_], *z1 = pl3[sz_];
                    int ca = 0, cb = 0;
                    if (HAS_U) {
                        const uchar2 cl = *reinterpret_cast<const uchar2 *>(stp + t.o_cls);
                        ca = cl.x; cb = cl.y;
                    }
                    if (!HAS_U || (ca | cb) == 0) {
                        // ---- both cells are non-conductor ----
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const double *tt = xs + a * TILE_D + o_own;
                            const double2 ym = in2<SAS>(tt - BW, A2, alpha);
                            const double2 yp = in2<SAS>(tt + BW, A2, alpha);
                            const double xm = in1<SAS>(tt - 1, A2, alpha), xp = in1<SAS>(tt + 2, A2, alpha);
                            const double ya = row7(czm, t.cym, cf.msx, dgA, t.cxpA, t.cyp, czp, m[a].x, ym.x, xm, c[a].x, c[a].y, yp.x, z1[a].x);
                            const double yb = row7(czm, t.cym, t.cxmB, dgB, cf.msx, t.cyp, czp, m[a].y, ym.y, c[a].x, c[a].y, xp, yp.y, z1[a].y);
                            pair_out<MODE>(ya, yb, true, true, a * G.segA + pA, c[a].x, c[a].y, aux[a], vs, a0, a1, a2, pxA, a, oA, jA, jB);
                        }
                    } else if (ca == 0x40 && cb == 0x40) {
                        // ---- both cells are interior conductor cells (all six neighbours conductor):
                        //      central grad U in the A rows (EC3D.f90:677-679), 13-entry U rows (:917-922) ----
                        const double2 ugm = ug3[sm_], ugc = ug3[sc_], ugp = ug3[sz_];
                        const double *tu = xs + USD + o_own;
                        const double uxm = in1<SAS>(tu - 1, U2, alpha), uxp = in1<SAS>(tu + 2, U2, alpha);
                        const double2 uym = in2<SAS>(tu - BW, U2, alpha);
                        const double2 uyp = in2<SAS>(tu + BW, U2, alpha);
                        double suA = 0.0, suB = 0.0;              // running A part of the U rows
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const double *tt = xs + a * TILE_D + o_own;
                            const double2 ym = in2<SAS>(tt - BW, A2, alpha);
                            const double2 yp = in2<SAS>(tt + BW, A2, alpha);
                            const double xm = in1<SAS>(tt - 1, A2, alpha), xp = in1<SAS>(tt + 2, A2, alpha);
                            double ya = row7(mc.cm[2], mc.cm[1], mc.cm[0], mc.diag, mc.cp[0], mc.cp[1], mc.cp[2], m[a].x, ym.x, xm,
                                             c[a].x, c[a].y, yp.x, z1[a].x);
                            double yb = row7(mc.cm[2], mc.cm[1], mc.cm[0], mc.diag, mc.cp[0], mc.cp[1], mc.cp[2], m[a].y, ym.y, c[a].x,
                                             c[a].y, xp, yp.y, z1[a].y);
                            // U(-1), U(+1) and A(-1), A(+1) along this component's axis
                            const double umA = (a == 0) ? uxm : (a == 1) ? uym.x : ugm.x;
                            const double upA = (a == 0) ? ugc.y : (a == 1) ? uyp.x : ugp.x;
                            const double umB = (a == 0) ? ugc.x : (a == 1) ? uym.y : ugm.y;
                            const double upB = (a == 0) ? uxp : (a == 1) ? uyp.y : ugp.y;
                            ya = DADD(ya, DMUL(mc.g1[a], umA)); ya = DADD(ya, DMUL(-mc.g1[a], upA));
                            yb = DADD(yb, DMUL(mc.g1[a], umB)); yb = DADD(yb, DMUL(-mc.g1[a], upB));
                            const double amA = (a == 0) ? xm : (a == 1) ? ym.x : m[a].x;
                            const double apA = (a == 0) ? c[a].y : (a == 1) ? yp.x : z1[a].x;
                            const double amB = (a == 0) ? c[a].x : (a == 1) ? ym.y : m[a].y;
                            const double apB = (a == 0) ? xp : (a == 1) ? yp.y : z1[a].y;
                            suA = DADD(suA, DMUL(cf.ua_p[a], amA)); suA = DADD(suA, DMUL(cf.ua_m[a], apA));
                            suB = DADD(suB, DMUL(cf.ua_p[a], amB)); suB = DADD(suB, DMUL(cf.ua_m[a], apB));
                            pair_out<MODE>(ya, yb, true, true, a * G.segA + pA, c[a].x, c[a].y, aux[a], vs, a0, a1, a2, pxA, a, oA, jC, jC);
                        }
                        // U columns k-1, j-1, i-1, centre, i+1, j+1, k+1
                        suA = DADD(suA, DMUL(cf.msz, ugm.x)); suB = DADD(suB, DMUL(cf.msz, ugm.y));
                        suA = DADD(suA, DMUL(cf.msy, uym.x)); suB = DADD(suB, DMUL(cf.msy, uym.y));
                        suA = DADD(suA, DMUL(cf.msx, uxm));   suB = DADD(suB, DMUL(cf.msx, ugc.x));
                        suA = DADD(suA, DMUL(cf.diag_int, ugc.x)); suB = DADD(suB, DMUL(cf.diag_int, ugc.y));
                        suA = DADD(suA, DMUL(cf.msx, ugc.y)); suB = DADD(suB, DMUL(cf.msx, uxp));
                        suA = DADD(suA, DMUL(cf.msy, uyp.x)); suB = DADD(suB, DMUL(cf.msy, uyp.y));
                        suA = DADD(suA, DMUL(cf.msz, ugp.x)); suB = DADD(suB, DMUL(cf.msz, ugp.y));
                        pair_out<MODE>(suA, suB, true, true, pU, ugc.x, ugc.y, auxU, vs, a0, a1, a2, pxU, 3, oU, jU, jU);
                    } else {
                        // ---- conductor-surface cells / mixed pairs (never on a domain face): generic per-cell rows ----
                        const double2 ugm = ug3[sm_], ugc = ug3[sc_], ugp = ug3[sz_];
                        const double *us = xs + USD;
                        U5 ua[3], ub[3];                          // U along x / y / z for cell a / b
                        {
                            const double *tt = us + o_own;
                            const double t_m2 = in1<SAS>(tt - 2, U2, alpha), t_m1 = in1<SAS>(tt - 1, U2, alpha);
                            const double t_p2 = in1<SAS>(tt + 2, U2, alpha), t_p3 = in1<SAS>(tt + 3, U2, alpha);
                            ua[0] = U5{t_m2, t_m1, ugc.x, ugc.y, t_p2};
                            ub[0] = U5{t_m1, ugc.x, ugc.y, t_p2, t_p3};
                            const double2 um = in2<SAS>(tt - BW, U2, alpha);
                            const double2 up = in2<SAS>(tt + BW, U2, alpha);
                            ua[1] = U5{0.0, um.x, ugc.x, up.x, 0.0};
                            ub[1] = U5{0.0, um.y, ugc.y, up.y, 0.0};
                            ua[2] = U5{0.0, ugm.x, ugc.x, ugp.x, 0.0};
                            ub[2] = U5{0.0, ugm.y, ugc.y, ugp.y, 0.0};
                            // one-sided gradients along y / z reach two cells: read those from the dense box
                            const int sya = (ca >> 2) & 3, syb = (cb >> 2) & 3, sza = (ca >> 4) & 3, szb = (cb >> 4) & 3;
                            if (sya & 2) ua[1].m2 = uglob(pU - 2 * G.ub_nx); else if (sya & 1) ua[1].p2 = uglob(pU + 2 * G.ub_nx);
                            if (syb & 2) ub[1].m2 = uglob(pU + 1 - 2 * G.ub_nx); else if (syb & 1) ub[1].p2 = uglob(pU + 1 + 2 * G.ub_nx);
                            if (sza & 2) ua[2].m2 = uglob(pU - 2 * G.ub_pl); else if (sza & 1) ua[2].p2 = uglob(pU + 2 * G.ub_pl);
                            if (szb & 2) ub[2].m2 = uglob(pU + 1 - 2 * G.ub_pl); else if (szb & 1) ub[2].p2 = uglob(pU + 1 + 2 * G.ub_pl);
                        }
                        double suA = 0.0, suB = 0.0;              // running A part of the U rows
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const double *tt = xs + a * TILE_D + o_own;
                            const double2 ym = in2<SAS>(tt - BW, A2, alpha);
                            const double2 yp = in2<SAS>(tt + BW, A2, alpha);
                            const double xm = in1<SAS>(tt - 1, A2, alpha), xp = in1<SAS>(tt + 2, A2, alpha);
                            double ya, yb;
                            if (ca) {
                                const int sa = (ca >> (2 * a)) & 3;
                                ya = (a == 0) ? cond_a<0>(mc, sa, m[a].x, ym.x, xm, c[a].x, c[a].y, yp.x, z1[a].x, ua[0])
                                   : (a == 1) ? cond_a<1>(mc, sa, m[a].x, ym.x, xm, c[a].x, c[a].y, yp.x, z1[a].x, ua[1])
                                              : cond_a<2>(mc, sa, m[a].x, ym.x, xm, c[a].x, c[a].y, yp.x, z1[a].x, ua[2]);
                                const double am = (a == 0) ? xm : (a == 1) ? ym.x : m[a].x;
                                const double ap = (a == 0) ? c[a].y : (a == 1) ? yp.x : z1[a].x;
                                suA = (a == 0) ? urow_a<0>(cf, ca, am, c[a].x, ap, suA)
                                    : (a == 1) ? urow_a<1>(cf, ca, am, c[a].x, ap, suA)
                                               : urow_a<2>(cf, ca, am, c[a].x, ap, suA);
                            } else {
                                ya = row7(czm, t.cym, cf.msx, dgA, t.cxpA, t.cyp, czp, m[a].x, ym.x, xm, c[a].x, c[a].y, yp.x, z1[a].x);
                            }
                            if (cb) {
                                const int sb = (cb >> (2 * a)) & 3;
                                yb = (a == 0) ? cond_a<0>(mc, sb, m[a].y, ym.y, c[a].x, c[a].y, xp, yp.y, z1[a].y, ub[0])
                                   : (a == 1) ? cond_a<1>(mc, sb, m[a].y, ym.y, c[a].x, c[a].y, xp, yp.y, z1[a].y, ub[1])
                                              : cond_a<2>(mc, sb, m[a].y, ym.y, c[a].x, c[a].y, xp, yp.y, z1[a].y, ub[2]);
                                const double am = (a == 0) ? c[a].x : (a == 1) ? ym.y : m[a].y;
                                const double ap = (a == 0) ? xp : (a == 1) ? yp.y : z1[a].y;
                                suB = (a == 0) ? urow_a<0>(cf, cb, am, c[a].y, ap, suB)
                                    : (a == 1) ? urow_a<1>(cf, cb, am, c[a].y, ap, suB)
                                               : urow_a<2>(cf, cb, am, c[a].y, ap, suB);
                            } else {
                                yb = row7(czm, t.cym, t.cxmB, dgB, cf.msx, t.cyp, czp, m[a].y, ym.y, c[a].x, c[a].y, xp, yp.y, z1[a].y);
                            }
                            pair_out<MODE>(ya, yb, true, true, a * G.segA + pA, c[a].x, c[a].y, aux[a], vs, a0, a1, a2, pxA, a, oA,
                                           ca ? jC : jA, cb ? jC : jB);
                        }
                        // U rows
                        double sa_ = 0.0, sb_ = 0.0;
                        if (ca) sa_ = urow_u(cf, ca, ua[2].m1, ua[1].m1, ua[0].m1, ugc.x, ua[0].p1, ua[1].p1, ua[2].p1, suA);
                        if (cb) sb_ = urow_u(cf, cb, ub[2].m1, ub[1].m1, ub[0].m1, ugc.y, ub[0].p1, ub[1].p1, ub[2].p1, suB);
                        pair_out<MODE>(sa_, sb_, ca != 0, cb != 0, pU, ugc.x, ugc.y, auxU, vs, a0, a1, a2, pxU, 3, oU, jU, jU);
                    }
                }
                if (MODE != MODE_PLAIN) {
                    dd_add_d(acc0, a0);
                    if (MODE != MODE_AP) dd_add_d(acc1, a1);
                    if (MODE == MODE_SAS) dd_add_d(acc2, a2);
                }
                pA += kdz; pU += G.ub_pl;
                release(q - 1);
                advance();
            }
        }
    }
}

// One launch handles the work items of ONE kind: HAS_U (conductor cells present: U tiles, class
// bytes, all row rules; CPS = 1 CTA per SM gives it up to 255 registers, no spills) or lean
// (7-point rows only, CPS >= 2).  Both launches of an SpMV share the reduction ticket: `expected`
// = items of both, partial index = pbase + blockIdx.x.
template <int MODE, int NSTAGE, bool HAS_U, int CPS, bool JAC>
__global__ void __launch_bounds__(256, CPS)
k_spmv_tma(const __grid_constant__ TmaMaps tm, const SlabGeom G, const Coef cf, const MatCoef mc,
           const WorkItem *__restrict__ items, const VecSet vs, const IterCtl ctl, double *partials, const int pstride,
           const int pbase, const unsigned expected, const __grid_constant__ PeerTable pt, CommLocal *cl, const int xf)
{
    using namespace tma;
    extern __shared__ unsigned char smem_raw[];
    __shared__ double sh[64];
    __shared__ __align__(8) unsigned long long full[NSTAGE];
    __shared__ unsigned cnt[NSTAGE];
    if (!spmv_guard<MODE>(ctl)) return;
    unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const WorkItem w = items[blockIdx.x];
    TmaCtx t;
    t.tx = threadIdx.x; t.ty = threadIdx.y; t.lane = threadIdx.x; t.warp = threadIdx.y;
    t.x0 = w.x0; t.y0 = w.y0; t.kb = w.kb; t.ke = w.ke; t.nload = (w.ke - w.kb) + 2;
    const int i0 = t.x0 + 2 * t.tx, j = t.y0 + t.ty;
    t.active = (i0 < G.sdx) && (j < G.sdy);
    const bool xlA = (i0 == 0), xhB = (i0 + 2 == G.sdx), yl = (j == 0), yh = (j == G.sdy - 1);
    // thread-constant coefficients of non-conductor rows (EC3D.f90:528-654): on a low face the '+'
    // neighbour carries BND(axis,2)*s, on a high face the '-' neighbour carries BND(axis,1)*s
    t.cxpA = xlA ? cf.blo[0] : cf.msx;                      // cell a's right neighbour (= cell b)
    t.cxmB = xhB ? cf.bhi[0] : cf.msx;                      // cell b's left neighbour (= cell a)
    t.cym = yh ? cf.bhi[1] : cf.msy; t.cyp = yl ? cf.blo[1] : cf.msy;
    const int bxyA = (int)xlA | ((int)(yl | yh) << 1), bxyB = (int)xhB | ((int)(yl | yh) << 1);
    t.dgA0 = cf.diag_b[bxyA]; t.dgA1 = cf.diag_b[bxyA | 4];   // diag_b[0] == diag_int
    t.dgB0 = cf.diag_b[bxyB]; t.dgB1 = cf.diag_b[bxyB | 4];
    t.o_own = (t.ty + 1) * BW + 2 + 2 * t.tx;               // own pair inside a halo tile (doubles)
    t.o_cls = Stage<MODE, HAS_U>::CLS_OFF + t.ty * TX + 2 * t.tx;   // own pair's class bytes inside a stage
    t.inUxy = t.active && i0 >= G.ub_i0 && i0 < G.ub_i0 + G.ub_nx && j >= G.ub_j0 && j < G.ub_j0 + G.ub_ny;
    t.alpha = 0.0;
    if (MODE == MODE_SAS) {
        // alpha = rr0/(AP,R0)   solvers.f90:31-32 (every thread derives it from the device scalars)
        const Scal *sc = ctl.sc;
        const int it = *ctl.iter_base + ctl.it_off;
        const double rr0 = (it == 1) ? sc->red[RED_RR_INIT] : sc->rr0[it & 1];
        t.alpha = rr0 / sc->red[RED_APR0];
        if (is_block0()) ctl.sc->alpha = t.alpha;
    }

    if (t.warp == 0 && t.lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full + s, 1); cnt[s] = 0u; }
        fence_barrier_init();
        fence_proxy_async();
        if (xf && (MODE == MODE_AP || MODE == MODE_SAS)) {
            // only the items that read halo planes (A: kb-1 / ke; U: two planes) wait for the neighbours'
            // fused push of this kernel's input vector(s); everything else starts right away
            const bool lo = w.kb < G.k0 + 2, hi = w.ke > G.k1 - 2;
            if (lo || hi) {
                if (MODE == MODE_AP) halo_wait(pt, cl, HALO_P, lo, hi);
                else { halo_wait(pt, cl, HALO_R, lo, hi); halo_wait(pt, cl, HALO_AP, lo, hi); }
                asm volatile("fence.proxy.async;" ::: "memory");   // peer-written halo data -> the TMA (async proxy) reads below
            }
        }
    }
    __syncthreads();

    dd a0 = dd_zero(), a1 = dd_zero(), a2 = dd_zero();
    tma_plane_loop<MODE, NSTAGE, HAS_U, JAC>(tm, G, cf, mc, vs, t, smem, full, cnt, a0, a1, a2, &pt, xf != 0);

    if (MODE != MODE_PLAIN) {
        const int pidx = pbase + blockIdx.x;
        if (MODE == MODE_AP)
            reduce_epilogue<1>(a0, dd_zero(), dd_zero(), partials, pstride, pidx, expected, ctl.sc, RED_APR0, RED_APR0, RED_APR0, sh,
                               xf ? XchgCtx{&pt, cl, 1 << HALO_AP} : no_xchg());
        else if (MODE == MODE_AS)
            reduce_epilogue<2>(a0, a1, dd_zero(), partials, pstride, pidx, expected, ctl.sc, RED_ASS, RED_ASAS, RED_ASAS, sh);
        else if (MODE == MODE_SAS)
            reduce_epilogue<3>(a0, a1, a2, partials, pstride, pidx, expected, ctl.sc, RED_ASS, RED_ASAS, RED_SS, sh,
                               xf ? XchgCtx{&pt, cl, 0} : no_xchg());
        else
            reduce_epilogue<2>(a0, a1, dd_zero(), partials, pstride, pidx, expected, ctl.sc, RED_BB, RED_RR_INIT, RED_RR_INIT, sh,
                               xf ? XchgCtx{&pt, cl, (1 << HALO_P) | (1 << HALO_R)} : no_xchg());
    }
}

// reference U numbering (k,j,i order over this rank's conductor cells) <-> dense U box
__global__ void k_u_unpack(const SlabGeom G, const int *__restrict__ cond_cells, const int ncond,
                           const double *__restrict__ compact, double *__restrict__ vec)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncond) return;
    const int cell0 = cond_cells[t];
    const int k = cell0 / G.kdz, rem = cell0 - k * G.kdz, j = rem / G.sdx, i = rem - j * G.sdx;
    vec[u_local(G, i, j, k)] = compact[t];
}

__global__ void k_u_pack(const SlabGeom G, const int *__restrict__ cond_cells, const int ncond,
                         const double *__restrict__ vec, double *__restrict__ compact)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncond) return;
    const int cell0 = cond_cells[t];
    const int k = cell0 / G.kdz, rem = cell0 - k * G.kdz, j = rem / G.sdx, i = rem - j * G.sdx;
    compact[t] = vec[u_local(G, i, j, k)];
}
