// ec3d_p2p.cuh -- halo exchange and scalar all-reduce over NVLink peer memory (one process per GPU).
//
// The BiCGSTABwr iteration has two nearest-neighbour exchanges (one plane of Ax, Ay, Az and two
// planes of the dense U box per neighbour, before each SpMV) and four reductions of 1-2 doubles.
// Both are latency, not bandwidth, problems, so they are done by small kernels of this library
// that store straight into the neighbour's memory (CUDA IPC mappings of every rank's vector
// allocation and of a small CommBlock) and signal with monotonically increasing epochs:
//
//   k_halo_push   copies the boundary planes of one vector into the neighbours' halo slots, every
//                 block fences (system scope), the last block raises the neighbours' halo flag
//   k_halo_wait   one thread waits until both neighbours' flags reached this rank's epoch
//   k_reduce_xchg stores this rank's partial result(s) into every rank's CommBlock (slot chosen by
//                 epoch parity), raises the flags, waits for all contributions and sums them in
//                 rank order in double-double -- bit-identical on every rank (all ranks take the
//                 same branches) and, rounded, the same value a single GPU computes
//
// All epochs live in device memory and are advanced by the kernels themselves, so a whole chunk
// of iterations (compute + exchange) is one CUDA graph.  Two pushes into the same halo slot are
// always separated by a reduction the reader takes part in after its SpMV, and a reduction slot is
// reused only two epochs later, so there are no write-after-read hazards.  Waits give up after
// ~30 s and set CommLocal::error instead of hanging the GPU.
#pragma once
#include "ec3d_common.cuh"
#include "ec3d_kernels.cuh"

#define EC3D_MAX_RANKS 16

struct CommBlock {                                   // written by peers
    unsigned long long halo_flag[2];                 // [0] from rank-1, [1] from rank+1: epoch of their last push
    unsigned long long red_flag[EC3D_MAX_RANKS];     // epoch of rank r's last contribution
    double red_val[2][EC3D_MAX_RANKS][8];            // [epoch parity][rank][hi0, lo0, hi1, lo1, hi2, lo2, -, -]
};

struct CommLocal {                                   // this rank only
    unsigned long long halo_epoch, red_epoch;
    unsigned int ticket;
    int error;                                       // 1: a wait timed out
};

struct PeerGeom {                                    // what a rank needs to know about a neighbour's layout
    long long segA, offU, nUlo, nUown, ltot;
    int nzl, pad;
};

struct PeerTable {
    int nranks, rank;
    CommBlock *cb[EC3D_MAX_RANKS];                   // every rank's CommBlock (own: local pointer)
    double *vecs_lo, *vecs_hi;                       // vector allocations of rank-1 / rank+1 (or null)
    PeerGeom g_lo, g_hi;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// spins until *p >= want; false on timeout
__device__ __forceinline__ bool wait_epoch(const unsigned long long *p, unsigned long long want)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < want) {
        if (clock64() - t0 > 60000000000LL) return false;      // ~30 s: ranks may be skewed by host work
        __nanosleep(64);
    }
    return true;
}

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
    const unsigned long long e = cl->halo_epoch + 1ull;
    cl->halo_epoch = e;
    cl->ticket = 0u;
    if (pt.rank > 0) st_release_sys(&pt.cb[pt.rank - 1]->halo_flag[1], e);            // I am its upper neighbour
    if (pt.rank < pt.nranks - 1) st_release_sys(&pt.cb[pt.rank + 1]->halo_flag[0], e);  // I am its lower neighbour
}

__global__ void k_halo_wait(const PeerTable pt, const Scal *sc, const int check_done, CommLocal *cl)
{
    if (threadIdx.x != 0) return;
    if (check_done && sc->done) return;
    const unsigned long long e = cl->halo_epoch;       // already advanced by this rank's own push
    CommBlock *me = pt.cb[pt.rank];
    bool ok = true;
    if (pt.rank > 0) ok = wait_epoch(&me->halo_flag[0], e) && ok;
    if (pt.rank < pt.nranks - 1) ok = wait_epoch(&me->halo_flag[1], e) && ok;
    if (!ok) cl->error = 1;
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
