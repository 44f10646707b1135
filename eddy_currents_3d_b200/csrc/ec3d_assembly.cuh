// ec3d_assembly.cuh -- K1: CSR assembly kernels (count -> scan -> fill) that emit the reference's
// irow/jcol/valA and boundary-cell lists (gen_sparse_matrix, EC3D.f90:465-1049) from the voxel
// maps, plus geometry validation.  Values come from the host-built coefficient tables, so the
// device does no floating-point arithmetic here and valA is bit-exact by construction.
#pragma once
#include "ec3d_common.cuh"
#include "ec3d_rows.cuh"

struct CountVisitor {
    int n;
    int bad;
    __device__ __forceinline__ void a(int, long long, double) { ++n; }
    __device__ __forceinline__ void u(int g, double) { ++n; if (g <= 0) bad = 1; }
};

struct EmitVisitor {
    int *jcol;
    double *val;
    long long nC;
    int n;
    __device__ __forceinline__ void a(int comp, long long cell0, double coef)
    {
        jcol[n] = (int)(comp * nC + cell0 + 1);
        val[n] = coef;
        ++n;
    }
    __device__ __forceinline__ void u(int g, double coef) { jcol[n] = g; val[n] = coef; ++n; }
};

// Per conductor cell: validity + flags.  flags bit0..2 = nAx,nAy,nAz; bit3..5 = nFix,nFiy,nFiz;
// bit 6 = invalid geometry (the reference would STOP or read out of bounds).
__global__ void k_classify_conductor(const SlabGeom G, const Coef cf, const int *__restrict__ geo,
                                     const signed char *__restrict__ mat, const int nmat, const int mat0,
                                     const int *__restrict__ cond_cells, const int ncond,
                                     unsigned char *__restrict__ flags, int *__restrict__ nbad)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncond) return;
    const int cell0 = cond_cells[t];
    const int k = cell0 / G.kdz, rem = cell0 - k * G.kdz, j = rem / G.sdx, i = rem - j * G.sdx;
    GeoView gv{geo, G.sdx, G.sdy, G.kdz, G.k0 - 2, G.nzl + 4};
    int f = 0;
    const bool onb = (i == 0 || j == 0 || k == 0 || i == G.sdx - 1 || j == G.sdy - 1 || k == G.sdz - 1);
    const int m = mat[(long long)cell0 - (long long)(G.k0 - 2) * G.kdz];
    if (onb || m < 1 || m > nmat || m != mat0) {
        f = 64;
    } else {
        MatCoef mc{};   // values are irrelevant for classification
        CountVisitor cv{0, 0};
        for (int comp = 0; comp < 3; ++comp) f |= cond_a_row(mc, gv, i, j, k, comp, cv) << comp;
        const int uf = cond_u_row(cf, gv, i, j, k, 3, cv);
        f |= (uf & 7) << 3;
        if ((uf & 8) || cv.bad) f |= 64;
    }
    flags[t] = (unsigned char)f;
    if (f & 64) atomicAdd(nbad, 1);
}

// Row lengths in the reference's row order: [Ax rows | Ay rows | Az rows | U rows].
__global__ void k_asm_count(const SlabGeom G, const Coef cf, const int *__restrict__ geo, int *__restrict__ rowlen)
{
    const long long cell0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (cell0 >= G.nC) return;
    const int k = (int)(cell0 / G.kdz), rem = (int)(cell0 - (long long)k * G.kdz), j = rem / G.sdx, i = rem - j * G.sdx;
    GeoView gv{geo, G.sdx, G.sdy, G.kdz, G.k0 - 2, G.nzl + 4};
    const int g = gv.at(i, j, k);
    const bool onb = (i == 0 || j == 0 || k == 0 || i == G.sdx - 1 || j == G.sdy - 1 || k == G.sdz - 1);
    MatCoef mc{};
    for (int comp = 0; comp < 3; ++comp) {
        CountVisitor cv{0, 0};
        if (g != 0 && !onb) cond_a_row(mc, gv, i, j, k, comp, cv);
        else air_row(cf, G.sdx, G.sdy, G.sdz, G.kdz, i, j, k, comp, cv);
        rowlen[comp * G.nC + cell0] = cv.n;
    }
    if (g != 0) {
        CountVisitor cv{0, 0};
        cond_u_row(cf, gv, i, j, k, 3, cv);
        rowlen[g - 1] = cv.n;
    }
}

__global__ void k_asm_fill(const SlabGeom G, const Coef cf, const MatCoef *__restrict__ mcs,
                           const int *__restrict__ geo, const signed char *__restrict__ mat,
                           const long long *__restrict__ off, int *__restrict__ jcol, double *__restrict__ valA)
{
    const long long cell0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (cell0 >= G.nC) return;
    const int k = (int)(cell0 / G.kdz), rem = (int)(cell0 - (long long)k * G.kdz), j = rem / G.sdx, i = rem - j * G.sdx;
    GeoView gv{geo, G.sdx, G.sdy, G.kdz, G.k0 - 2, G.nzl + 4};
    const int g = gv.at(i, j, k);
    const bool onb = (i == 0 || j == 0 || k == 0 || i == G.sdx - 1 || j == G.sdy - 1 || k == G.sdz - 1);
    for (int comp = 0; comp < 3; ++comp) {
        const long long o = off[comp * G.nC + cell0];
        EmitVisitor ev{jcol + o, valA + o, G.nC, 0};
        if (g != 0 && !onb) {
            const MatCoef mc = mcs[mat[cell0 - (long long)(G.k0 - 2) * G.kdz] - 1];
            cond_a_row(mc, gv, i, j, k, comp, ev);
        } else {
            air_row(cf, G.sdx, G.sdy, G.sdz, G.kdz, i, j, k, comp, ev);
        }
    }
    if (g != 0) {
        const long long o = off[g - 1];
        EmitVisitor ev{jcol + o, valA + o, G.nC, 0};
        cond_u_row(cf, gv, i, j, k, 3, ev);
    }
}

// ---- exclusive scan of int32 lengths into int64 offsets (3 passes) ----
#define SCAN_T 256
#define SCAN_I 16
__global__ void __launch_bounds__(SCAN_T)
k_scan_block(const int *__restrict__ in, long long *__restrict__ out, long long n, long long *__restrict__ bsum)
{
    __shared__ long long sh[SCAN_T];
    const long long base = (long long)blockIdx.x * SCAN_T * SCAN_I + (long long)threadIdx.x * SCAN_I;
    long long loc[SCAN_I];
    long long s = 0;
#pragma unroll
    for (int q = 0; q < SCAN_I; ++q) {
        const long long idx = base + q;
        loc[q] = s;
        s += (idx < n) ? in[idx] : 0;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < SCAN_T; o <<= 1) {
        long long v = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    const long long excl = sh[threadIdx.x] - s;
#pragma unroll
    for (int q = 0; q < SCAN_I; ++q) {
        const long long idx = base + q;
        if (idx < n) out[idx] = excl + loc[q];
    }
    if (threadIdx.x == SCAN_T - 1) bsum[blockIdx.x] = sh[threadIdx.x];
}

__global__ void k_scan_top(long long *bsum, int nb, long long *total)
{
    long long s = 0;
    for (int q = 0; q < nb; ++q) { long long v = bsum[q]; bsum[q] = s; s += v; }
    *total = s;
}

__global__ void __launch_bounds__(SCAN_T)
k_scan_add(long long *__restrict__ out, long long n, const long long *__restrict__ bsum)
{
    const long long base = (long long)blockIdx.x * SCAN_T * SCAN_I + (long long)threadIdx.x * SCAN_I;
    const long long add = bsum[blockIdx.x];
#pragma unroll
    for (int q = 0; q < SCAN_I; ++q) {
        const long long idx = base + q;
        if (idx < n) out[idx] += add;
    }
}

__global__ void k_off_to_irow(const long long *__restrict__ off, long long n, long long total, int *__restrict__ irow)
{
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q < n) irow[q] = (int)(off[q] + 1);
    if (q == n) irow[n] = (int)(total + 1);
}

__global__ void k_flag_bit(const unsigned char *__restrict__ flags, int n, int bit, int *__restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) out[q] = (flags[q] >> bit) & 1;
}

// list = cell number (+ comp*nC) for bits 0..2, geoPHYS_C value for bits 3..5
__global__ void k_compact_list(const SlabGeom G, const unsigned char *__restrict__ flags, int n, int bit,
                               const long long *__restrict__ pos, const int *__restrict__ cond_cells,
                               const int *__restrict__ geo, int *__restrict__ list)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n || !((flags[q] >> bit) & 1)) return;
    const int cell0 = cond_cells[q];
    int v;
    if (bit < 3) v = (int)(cell0 + 1 + bit * G.nC);
    else v = geo[(long long)cell0 - (long long)(G.k0 - 2) * G.kdz];
    list[pos[q]] = v;
}
