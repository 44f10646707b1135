// ec3d_comm.cuh -- peer-memory (NVLink, CUDA IPC) communication state of a rank and the device-side
// primitives the compute kernels use to exchange with their z-neighbours WITHOUT extra launches.
//
// The BiCGSTABwr iteration has two kinds of exchange:
//   * nearest-neighbour halos of the SpMV inputs: one plane of Ax, Ay, Az and the two nearest planes
//     of the dense U box per neighbour.  The kernel that PRODUCES a vector stores its boundary planes
//     straight into the neighbour's halo slots (peer_push*), the kernel's last block raises a per-kind
//     epoch flag in the neighbour's CommBlock, and only the SpMV work items that touch a halo plane
//     wait for that flag (halo_wait) -- interior items never wait;
//   * reductions of 1-3 scalars: the last block of the kernel that finishes a local dot product
//     writes the rank's double-double partial(s) into every rank's CommBlock, waits for all
//     contributions and sums them in rank order (xchg_reduce) -- identical bits on every rank.
// Epochs are monotonically increasing counters kept in device memory, so a whole chunk of iterations
// (compute + exchange) is one CUDA graph per rank.  Hazards: a halo slot of kind K is rewritten only
// after a reduction that the reader joined AFTER its last read of that slot; reduction slots alternate
// by epoch parity.  Waits give up after ~30 s and set CommLocal::error instead of hanging the GPU.
#pragma once
#include "ec3d_common.cuh"

#define EC3D_MAX_RANKS 16

enum { HALO_GEN = 0, HALO_P = 1, HALO_R = 2, HALO_AP = 3, HALO_KINDS = 4 };

struct CommBlock {                                   // written by peers
    unsigned long long halo_flag[HALO_KINDS][2];     // [kind][0: from rank-1, 1: from rank+1]: epoch of their last push
    unsigned long long red_flag[EC3D_MAX_RANKS];     // epoch of rank r's last contribution
    double red_val[2][EC3D_MAX_RANKS][8];            // [epoch parity][rank][hi0, lo0, hi1, lo1, hi2, lo2, -, -]
};

struct CommLocal {                                   // this rank only
    unsigned long long halo_epoch[HALO_KINDS];       // pushes of each kind done by this rank
    unsigned long long red_epoch;
    unsigned int ticket;                             // k_halo_push
    unsigned int ticket_p;                           // fused push of the p-update kernel
    int error;                                       // 1: a wait timed out
    int pad;
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
    long long nU_send_lo, nU_send_hi;                // U entries of my first / last two planes (dense box)
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
        __nanosleep(32);
    }
    return true;
}

// ---- fused halo push -------------------------------------------------------------------------
// Owned entry pair at offset o (even) inside owned segment `seg` (0..2: A components, 3: U) of local
// vector `vi`: if it lies in a boundary plane, store it into the neighbour's halo slot as well.
__device__ __forceinline__ void peer_push2(const PeerTable &pt, const SlabGeom &G, const int vi, const int seg, const long long o,
                                           const double a, const double b)
{
    const double2 v = make_double2(a, b);
    if (seg < 3) {
        const long long kdz = G.kdz;
        if (pt.vecs_lo && o < kdz)               // my first owned plane -> rank-1's upper halo plane
            *reinterpret_cast<double2 *>(pt.vecs_lo + (long long)vi * pt.g_lo.ltot + seg * pt.g_lo.segA +
                                         (long long)(pt.g_lo.nzl + 1) * kdz + o) = v;
        const long long last = (long long)(G.nzl - 1) * kdz;
        if (pt.vecs_hi && o >= last)             // my last owned plane -> rank+1's lower halo plane
            *reinterpret_cast<double2 *>(pt.vecs_hi + (long long)vi * pt.g_hi.ltot + seg * pt.g_hi.segA + (o - last)) = v;
    } else {
        if (pt.vecs_lo && o < pt.nU_send_lo)     // my first two U planes -> rank-1's U halo above
            *reinterpret_cast<double2 *>(pt.vecs_lo + (long long)vi * pt.g_lo.ltot + pt.g_lo.offU + pt.g_lo.nUlo +
                                         pt.g_lo.nUown + o) = v;
        const long long first_hi = G.nUown - pt.nU_send_hi;
        if (pt.vecs_hi && o >= first_hi)         // my last two U planes -> rank+1's U halo below
            *reinterpret_cast<double2 *>(pt.vecs_hi + (long long)vi * pt.g_hi.ltot + pt.g_hi.offU + (o - first_hi)) = v;
    }
}
// single entry (pairs with only one existing row: conductor cells at the edge of the dense U box)
__device__ __forceinline__ void peer_push1(const PeerTable &pt, const SlabGeom &G, const int vi, const int seg, const long long o,
                                           const double a)
{
    if (seg < 3) {
        const long long kdz = G.kdz;
        if (pt.vecs_lo && o < kdz)
            pt.vecs_lo[(long long)vi * pt.g_lo.ltot + seg * pt.g_lo.segA + (long long)(pt.g_lo.nzl + 1) * kdz + o] = a;
        const long long last = (long long)(G.nzl - 1) * kdz;
        if (pt.vecs_hi && o >= last) pt.vecs_hi[(long long)vi * pt.g_hi.ltot + seg * pt.g_hi.segA + (o - last)] = a;
    } else {
        if (pt.vecs_lo && o < pt.nU_send_lo)
            pt.vecs_lo[(long long)vi * pt.g_lo.ltot + pt.g_lo.offU + pt.g_lo.nUlo + pt.g_lo.nUown + o] = a;
        const long long first_hi = G.nUown - pt.nU_send_hi;
        if (pt.vecs_hi && o >= first_hi) pt.vecs_hi[(long long)vi * pt.g_hi.ltot + pt.g_hi.offU + (o - first_hi)] = a;
    }
}

// Called by ONE thread after every block of the producing kernel has fenced its peer stores
// (__threadfence_system before the ticket): tells both neighbours that halo `kind` is complete.
__device__ __forceinline__ void halo_raise(const PeerTable &pt, CommLocal *cl, const int kind)
{
    const unsigned long long e = cl->halo_epoch[kind] + 1ull;
    cl->halo_epoch[kind] = e;
    if (pt.rank > 0) st_release_sys(&pt.cb[pt.rank - 1]->halo_flag[kind][1], e);            // I am its upper neighbour
    if (pt.rank < pt.nranks - 1) st_release_sys(&pt.cb[pt.rank + 1]->halo_flag[kind][0], e);  // I am its lower neighbour
}

// Called by ONE thread of an SpMV work item that reads halo planes: waits until the neighbour(s) have
// delivered halo `kind` for this rank's current epoch of that kind (every rank pushes each kind equally often).
__device__ __forceinline__ void halo_wait(const PeerTable &pt, CommLocal *cl, const int kind, const bool lo, const bool hi)
{
    const unsigned long long e = cl->halo_epoch[kind];
    CommBlock *me = pt.cb[pt.rank];
    bool ok = true;
    if (lo && pt.rank > 0) ok = wait_epoch(&me->halo_flag[kind][0], e) && ok;
    if (hi && pt.rank < pt.nranks - 1) ok = wait_epoch(&me->halo_flag[kind][1], e) && ok;
    if (!ok) cl->error = 1;
}
