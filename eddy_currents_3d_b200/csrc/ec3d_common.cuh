// ec3d_common.cuh -- shared host/device data structures of libec3d_gpu.so.
//
// Data layout in HBM (per rank; a single GPU is the 1-rank case):
//   * z-slab: this rank owns planes k in [k0, k1) of the sdx*sdy*sdz grid (x fastest).
//   * every Krylov / field vector is ONE allocation of `ltot` doubles made of four segments
//       [ Ax : (nzl+2) planes | Ay : (nzl+2) planes | Az : (nzl+2) planes | U : nUlo+nUown+nUhi ]
//     Each A segment carries one halo plane below and above the owned planes.  The U segment is
//     DENSE over the conductor's bounding box (ub_i0.., ub_j0.., ub_k0..; x extent padded to even):
//     entry (i,j,k) sits at offU + (k-ub_kl0)*ub_pl + (j-ub_j0)*ub_nx + (i-ub_i0), planes
//     [ub_kl0, ub_kl1) = box planes within [k0-2, k1+2) (two halo planes: one-sided z-gradients reach
//     two cells).  Box cells that are not conductor cells are padding: they hold 0.0 in every
//     vector and no kernel writes them, so BLAS-1 kernels can stream over the whole segment and
//     the stencil reads U neighbours at fixed offsets (TMA tiles) instead of through geoPHYS_C.
//     The reference's compact U numbering (k,j,i order over conductor cells) exists only at the
//     C-ABI boundary (ec3d_get_fields / ec3d_set_fields pack and unpack).  Halo entries are
//     refreshed by the exchange that precedes each SpMV.
//   * geoPHYS_C / geoPHYS are stored for planes [k0-2, k1+2) (zero outside the domain).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define EC3D_MAX_FUN 256

// Stencil coefficients that do not depend on the material.  Built on the host by
// build_coef() with the reference's expression shapes (EC3D.f90:496-501, 528-654, 773-922).
struct Coef {
    double msx, msy, msz;          // -sx, -sy, -sz
    double blo[3], bhi[3];         // low face: BND(a,2)*s_a on the '+' neighbour; high face: BND(a,1)*s_a on '-'
    double diag_int;               // 2*(sx+sy+sz)
    double diag_b[8];              // domain-boundary diagonal, index = onx | ony<<1 | onz<<2
    double m2s[3];                 // -2*s_a               (U rows, neighbour opposite a missing one)
    double ua_m[3], ua_p[3];       // (0.5/dt)*(-1/d_a) on the '+' A neighbour, (0.5/dt)*(1/d_a) on '-'
    double uc_m[3], uc_p[3];       // same-cell couplings -2/(dt*d_a), +2/(dt*d_a)
};

// Per-material coefficients of conductor cells (EC3D.f90:656-710).
struct MatCoef {
    double cm[3], cp[3];           // -s_a - V_a/(2 d_a),  -s_a + V_a/(2 d_a)
    double diag;                   // 2*(sx+sy+sz) + 2*C/dt
    double g1[3], g3[3], g4[3];    // (C)*ds_a, (3C)*ds_a, (4C)*ds_a
};

struct SlabGeom {
    int sdx, sdy, sdz, kdz;
    int k0, k1, nzl;               // owned planes [k0,k1)
    long long nC;                  // global cell count
    long long segA;                // doubles per A segment (even)
    long long offU;                // 3*segA
    long long ltot;                // total local doubles
    // dense U box (global cell coordinates, half open); ub_nx even, ub_i0 even
    int ub_i0, ub_j0, ub_k0, ub_nx, ub_ny, ub_nz;
    int ub_kl0, ub_kl1;            // box planes stored by this rank: [ub_kl0, ub_kl1)
    long long ub_pl;               // ub_nx * ub_ny
    long long nUlo, nUown, nUhi;   // entries (padding included) of the halo-below / owned / halo-above planes
    // owned ranges inside the local vector: seg s starts at own_off[s], has own_len[s] entries
    long long own_off[4], own_len[4], own_cum[5];
    long long n_own;
    // global index (0-based, reference layout) of the first owned entry of each segment
    long long glob_off[4];
    int vmap;                      // generic BLAS-1 kernels: 0 grid-stride, 1 one contiguous range per block
    int pad_;
};

// Local index of the U unknown of cell (i,j,k) (0-based global coordinates) in the dense U box.
__host__ __device__ __forceinline__ long long u_local(const SlabGeom &G, int i, int j, int k)
{
    return G.offU + (long long)(k - G.ub_kl0) * G.ub_pl + (long long)(j - G.ub_j0) * G.ub_nx + (i - G.ub_i0);
}

// Device-resident solver scalars (one per handle).
struct Scal {
    double red[8];                 // reduction results (rounded; with several ranks: high words until the exchange)
    double red_lo[8];              // low words of the per-rank double-double results (several ranks only)
    double rr0[2];                 // (R,R0), double-buffered by iteration parity
    double alpha, omega, beta;
    double tol;
    int itmax;
    int done;                      // 1 once the solve has exited
    int final_iter;
    int exit_kind;                 // 0 running, 1 ||s|| test, 2 ||r|| test, 3 iter>itmax, 4 ||b||==0
    int restarts;
    unsigned int counter;          // last-block ticket
    int multi;                     // != 0: several ranks, reductions stay double-double until the exchange
};

enum { RED_BB = 0, RED_RR_INIT = 1, RED_APR0 = 2, RED_SS = 3, RED_ASS = 4, RED_ASAS = 5, RED_RR = 6, RED_RR0N = 7 };

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ec3d_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return EC3D_ERR_CUDA;                                                           \
        }                                                                                   \
    } while (0)

void ec3d_set_error(const char *fmt, ...);
