// ec3d_p2p.cuh -- stand-alone exchange kernels over NVLink peer memory (one process per GPU).
//
// Inside the BiCGSTABwr iteration the exchanges are FUSED into the compute kernels (ec3d_comm.cuh: halo
// planes are pushed by the kernel that produces a vector, reductions are finished across ranks by the
// last block of the kernel that computes them).  The kernels here serve everything outside that loop --
// the halo of the initial guess / of Uaf before the right-hand side, ec3d_apply_operator, the output
// fields, barriers -- and the un-fused fallback sequence (odd grids, EC3D_XFUSE=0):
//
//   k_halo_push   copies the boundary planes of one or two vectors into the neighbours' halo slots, every
//                 block fences (system scope), the last block raises the neighbours' HALO_GEN flag
//   k_halo_wait   one thread waits until both neighbours' flags reached this rank's epoch
//   k_reduce_xchg stores this rank's partial result(s) into every rank's CommBlock (slot chosen by
//                 epoch parity), raises the flags, waits for all contributions and sums them in rank
//                 order in double-double -- bit-identical on every rank and, rounded, the same value a
//                 single GPU computes
#pragma once
#include "ec3d_common.cuh"
#include "ec3d_comm.cuh"
#include "ec3d_kernels.cuh"

// Copies this rank's boundary planes of local vector `vidx` (and `vidx2` when >= 0) into the neighbours'
// halo slots.  Work units are 16-byte pairs; per vector up to 8 segments: 3 A planes + U planes towards each side.
__global__ void __launch_bounds__(256)
k_halo_push(const SlabGeom G, const PeerTable pt, double *__restrict__ vecs, const int vidx, const int vidx2,
            const long long nU_send_lo, const long long nU_send_hi, const Scal *sc, const int check_done, CommLocal *cl)
{
    if (check_done && sc->done) return;
    const long long kdz = G.kdz;
    const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, tstep = (long long)gridDim.x * blockDim.x;
    for (int w = 0; w < 2; ++w) {
        const int vi = w == 0 ? vidx : vidx2;
        if (vi < 0) continue;
        const double *src = vecs + (long long)vi * G.ltot;
        // towards rank-1: my first owned plane -> its upper halo plane; my first two U planes -> its U halo above
        if (pt.vecs_lo) {
            double *dst = pt.vecs_lo + (long long)vi * pt.g_lo.ltot;
            for (int c = 0; c < 3; ++c) {
                const double2 *s2 = reinterpret_cast<const double2 *>(src + c * G.segA + kdz);
                double2 *d2 = reinterpret_cast<double2 *>(dst + c * pt.g_lo.segA + (long long)(pt.g_lo.nzl + 1) * kdz);
                for (long long q = t0; q < kdz / 2; q += tstep) d2[q] = s2[q];
            }
            const double *su = src + G.offU + G.nUlo;
            double *du = dst + pt.g_lo.offU + pt.g_lo.nUlo + pt.g_lo.nUown;
            for (long long q = t0; q < nU_send_lo; q += tstep) du[q] = su[q];
        }
        // towards rank+1: my last owned plane -> its lower halo plane; my last two U planes -> its U halo below
        if (pt.vecs_hi) {
            double *dst = pt.vecs_hi + (long long)vi * pt.g_hi.ltot;
            for (int c = 0; c < 3; ++c) {
                const double2 *s2 = reinterpret_cast<const double2 *>(src + c * G.segA + (long long)G.nzl * kdz);
                double2 *d2 = reinterpret_cast<double2 *>(dst + c * pt.g_hi.segA);
                for (long long q = t0; q < kdz / 2; q += tstep) d2[q] = s2[q];
            }
            const double *su = src + G.offU + G.nUlo + G.nUown - nU_send_hi;
            double *du = dst + pt.g_hi.offU;
            for (long long q = t0; q < nU_send_hi; q += tstep) du[q] = su[q];
        }
    }
    // every block: make its stores visible system wide, then take a ticket; the last block signals
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned last;
    if (threadIdx.x == 0) last = (atomicAdd(&cl->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence_system();
    cl->ticket = 0u;
    halo_raise(pt, cl, HALO_GEN);
}

__global__ void k_halo_wait(const PeerTable pt, const Scal *sc, const int check_done, CommLocal *cl)
{
    if (threadIdx.x != 0) return;
    if (check_done && sc->done) return;
    halo_wait(pt, cl, HALO_GEN, true, true);           // epoch already advanced by this rank's own push
    __threadfence_system();
}

// sc->red[slot .. slot+count) <- sum over ranks (count <= 3), in rank order on every rank
__global__ void k_reduce_xchg(const PeerTable pt, Scal *sc, const int slot, const int count, const int force, CommLocal *cl)
{
    if (!force && sc->done) return;              // count == 0, force != 0: a plain barrier over all ranks
    const int lane = threadIdx.x;
    unsigned long long e = 0;
    if (lane == 0) { e = cl->red_epoch + 1ull; cl->red_epoch = e; }
    e = __shfl_sync(0xffffffffu, e, 0);
    const int par = (int)(e & 1ull);
    if (lane < pt.nranks) {
        CommBlock *dst = pt.cb[lane];
        for (int q = 0; q < count; ++q) {
            dst->red_val[par][pt.rank][2 * q] = sc->red[slot + q];
            dst->red_val[par][pt.rank][2 * q + 1] = sc->red_lo[slot + q];
        }
        __threadfence_system();
        st_release_sys(&dst->red_flag[pt.rank], e);
    }
    CommBlock *me = pt.cb[pt.rank];
    bool ok = true;
    if (lane < pt.nranks) ok = wait_epoch(&me->red_flag[lane], e);
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        if (!ok) cl->error = 1;
        for (int q = 0; q < count; ++q) {
            dd t = dd_zero();                        // double-double sum in rank order, rounded once
            for (int r = 0; r < pt.nranks; ++r)
                dd_add_dd(t, dd{*(volatile double *)&me->red_val[par][r][2 * q], *(volatile double *)&me->red_val[par][r][2 * q + 1]});
            sc->red[slot + q] = dd_round(t);
        }
    }
}
