// ec3d_rows.cuh -- the row rules of the coupled A-U operator as visitor templates.
//
// One definition of "which (column, coefficient) pairs make up a row, in ascending column order"
// serves the CSR assembly kernels (count / fill), the conductor part of the matrix-free SpMV, the
// U-row right-hand side and the geometry validation, so those cannot disagree with each other.
// What they must agree WITH is the reference, gen_sparse_matrix (EC3D.f90:465-1049); the CPU
// oracle restates that independently (literal case cascade) and the parity tests compare.
//
// Visitor concept:
//   void a(int comp, long long cell0, double coef);   // A column: component comp, 0-based GLOBAL cell
//   void u(int g, double coef);                       // U column: geoPHYS_C value g (3*nC + m); g <= 0 is invalid
//                                                     // (with a dense GeoView: 1 + index into the dense U box)
// Entries are visited in ascending column order (the order full_sort produces, utilites.f90:477).
#pragma once
#include "ec3d_common.cuh"

// Local view of geoPHYS_C around a slab: planes [gz0, gz0+gnz), zero outside.
// With dense != 0 a conductor cell is reported as 1 + its offset inside the dense U box of the local
// vector layout (SlabGeom::ub_*), which is what the gathering visitors of the SpMV / RHS kernels
// index with; the assembly kernels use the plain view (real geoPHYS_C values = CSR columns).
struct GeoView {
    const int *g;
    int sdx, sdy, kdz;
    int gz0, gnz;
    int dense, ub_i0, ub_j0, ub_kl0, ub_nx;
    long long ub_pl;
    __host__ __device__ __forceinline__ int at(int i, int j, int k) const {
        int kk = k - gz0;
        if (i < 0 || i >= sdx || j < 0 || j >= sdy || kk < 0 || kk >= gnz) return 0;
        const int v = g[(long long)kk * kdz + (long long)j * sdx + i];
        if (!dense || v == 0) return v;
        return 1 + (int)((long long)(k - ub_kl0) * ub_pl + (long long)(j - ub_j0) * ub_nx + (i - ub_i0));
    }
};

__host__ __device__ __forceinline__ GeoView dense_view(const SlabGeom &G, const int *geo)
{
    return GeoView{geo, G.sdx, G.sdy, G.kdz, G.k0 - 2, G.nzl + 4, 1, G.ub_i0, G.ub_j0, G.ub_kl0, G.ub_nx, G.ub_pl};
}

// ---- A row of a non-conductor cell or of any cell on a domain face (EC3D.f90:528-654) ----
// Domain-face cells ignore conductor properties (no kFi branch there).
template <class V>
__host__ __device__ __forceinline__ void air_row(const Coef &cf, int sdx, int sdy, int sdz, int kdz,
                                                 int i, int j, int k, int comp, V &v)
{
    const long long nn = (long long)k * kdz + (long long)j * sdx + i;
    const bool xl = (i == 0), xh = (i == sdx - 1), yl = (j == 0), yh = (j == sdy - 1), zl = (k == 0),
               zh = (k == sdz - 1);
    const bool onb = xl | xh | yl | yh | zl | zh;
    if (!zl) v.a(comp, nn - kdz, zh ? cf.bhi[2] : cf.msz);
    if (!yl) v.a(comp, nn - sdx, yh ? cf.bhi[1] : cf.msy);
    if (!xl) v.a(comp, nn - 1, xh ? cf.bhi[0] : cf.msx);
    v.a(comp, nn, onb ? cf.diag_b[(int)(xl | xh) | ((int)(yl | yh) << 1) | ((int)(zl | zh) << 2)] : cf.diag_int);
    if (!xh) v.a(comp, nn + 1, xl ? cf.blo[0] : cf.msx);
    if (!yh) v.a(comp, nn + sdx, yl ? cf.blo[1] : cf.msy);
    if (!zh) v.a(comp, nn + kdz, zl ? cf.blo[2] : cf.msz);
}

// ---- A row (component comp) of an interior conductor cell (EC3D.f90:649-710) ----
// Returns the nAx/nAy/nAz flag of that component (one-sided gradient => cel_bndX/Y/Z, :758-760).
template <class V>
__host__ __device__ __forceinline__ int cond_a_row(const MatCoef &mc, const GeoView &gv, int i, int j, int k,
                                                   int comp, V &v)
{
    const int sdx = gv.sdx, kdz = gv.kdz;
    const long long nn = (long long)k * kdz + (long long)j * sdx + i;
    v.a(comp, nn - kdz, mc.cm[2]);
    v.a(comp, nn - sdx, mc.cm[1]);
    v.a(comp, nn - 1, mc.cm[0]);
    v.a(comp, nn, mc.diag);
    v.a(comp, nn + 1, mc.cp[0]);
    v.a(comp, nn + sdx, mc.cp[1]);
    v.a(comp, nn + kdz, mc.cp[2]);
    const int di = (comp == 0), dj = (comp == 1), dk = (comp == 2);
    const int gc = gv.at(i, j, k);
    const int gp = gv.at(i + di, j + dj, k + dk);
    const int gm = gv.at(i - di, j - dj, k - dk);
    if (gp == 0) {                       // :667-671 (and :682, :697): backward one-sided gradient
        v.u(gv.at(i - 2 * di, j - 2 * dj, k - 2 * dk), -mc.g1[comp]);
        v.u(gm, mc.g4[comp]);
        v.u(gc, -mc.g3[comp]);
        return 1;
    } else if (gm == 0) {                // :672-676: forward one-sided gradient
        v.u(gc, mc.g3[comp]);
        v.u(gp, -mc.g4[comp]);
        v.u(gv.at(i + 2 * di, j + 2 * dj, k + 2 * dk), mc.g1[comp]);
        return 1;
    }
    v.u(gm, mc.g1[comp]);                // :677-679: central difference
    v.u(gp, -mc.g1[comp]);
    return 0;
}

// ---- U row of a conductor cell (EC3D.f90:766-922) ----
// part: 1 = A columns only, 2 = U columns only, 3 = both.  Returns flags: bit a set when
// nFix/nFiy/nFiz is set (=> cel_bndUx/Uy/Uz, :938-940); bit 3 set when the geometry is invalid
// (both neighbours missing along an axis: the reference ends up with a column 0 and STOPs).
template <class V>
__host__ __device__ __forceinline__ int cond_u_row(const Coef &cf, const GeoView &gv, int i, int j, int k, int part,
                                                   V &v)
{
    const int sdx = gv.sdx, kdz = gv.kdz;
    const long long nn = (long long)k * kdz + (long long)j * sdx + i;
    const int nc = gv.at(i, j, k);
    const int nim = gv.at(i - 1, j, k), nip = gv.at(i + 1, j, k);
    const int njm = gv.at(i, j - 1, k), njp = gv.at(i, j + 1, k);
    const int nkm = gv.at(i, j, k - 1), nkp = gv.at(i, j, k + 1);
    // per-axis state: 0 both neighbours present, 1 '-' missing, 2 '+' missing, 3 both missing
    const int sx = (nim == 0) | ((nip == 0) << 1);
    const int sy = (njm == 0) | ((njp == 0) << 1);
    const int sz = (nkm == 0) | ((nkp == 0) << 1);
    int flags = 0;
    if (sx == 3 || sy == 3 || sz == 3) flags |= 8;
    const int nmiss = (sx != 0) + (sy != 0) + (sz != 0);
    if (part & 1) {
        if (nmiss == 0) {                // interior, 13 entries (:917-922)
            v.a(0, nn - 1, cf.ua_p[0]);
            v.a(0, nn + 1, cf.ua_m[0]);
            v.a(1, nn - sdx, cf.ua_p[1]);
            v.a(1, nn + sdx, cf.ua_m[1]);
            v.a(2, nn - kdz, cf.ua_p[2]);
            v.a(2, nn + kdz, cf.ua_m[2]);
        } else {
            // corners / edges / faces couple to A at the SAME cell, only for axes with a missing
            // neighbour: -2/(dt*d) when the '-' neighbour is missing, +2/(dt*d) when '+' is.
            // Anomaly (:803-807): the corner with i-1, j+1, k+1 missing has a=+, b=- as written.
            const bool anomaly = (sx == 1 && sy == 2 && sz == 2);
            if (sx) v.a(0, nn, anomaly ? cf.uc_p[0] : (sx == 1 ? cf.uc_m[0] : cf.uc_p[0]));
            if (sy) v.a(1, nn, anomaly ? cf.uc_m[1] : (sy == 1 ? cf.uc_m[1] : cf.uc_p[1]));
            if (sz) v.a(2, nn, sz == 1 ? cf.uc_m[2] : cf.uc_p[2]);
        }
    }
    if (part & 2) {
        // U columns ascending: k-1, j-1, i-1, centre, i+1, j+1, k+1 (numbering is k,j,i ordered)
        if (sz == 0) v.u(nkm, cf.msz); else if (sz == 2) v.u(nkm, cf.m2s[2]);
        if (sy == 0) v.u(njm, cf.msy); else if (sy == 2) v.u(njm, cf.m2s[1]);
        if (sx == 0) v.u(nim, cf.msx); else if (sx == 2) v.u(nim, cf.m2s[0]);
        v.u(nc, cf.diag_int);
        if (sx == 0) v.u(nip, cf.msx); else if (sx == 1) v.u(nip, cf.m2s[0]);
        if (sy == 0) v.u(njp, cf.msy); else if (sy == 1) v.u(njp, cf.m2s[1]);
        if (sz == 0) v.u(nkp, cf.msz); else if (sz == 1) v.u(nkp, cf.m2s[2]);
    }
    if (nmiss != 0) flags |= (sx != 0) | ((sy != 0) << 1) | ((sz != 0) << 2);
    return flags;
}
