// stream_probe.cu -- stand-alone HBM streaming probe for the BLAS-1 kernels of the BiCGSTABwr iteration.
//
// Question it answers: why do k_xr_update / k_p_update run at 0.95 of the measured
// copy bandwidth on plate(256) (0.42 GB per vector) but at 0.78-0.83 on plate(512) (3.4 GB per vector)?
// The probe runs the xr-update arithmetic (5 read streams, 2 write streams, two double-double dots)
// outside the solver at both vector sizes with several access schemes:
//   ls      grid-stride 16-byte loads (the r01 kernel)
//   ls_u2   same, two independent units per loop trip
//   ls_pf   same + prefetch.global.L2 of the lines `dist` bytes ahead in every stream
//   blk     one contiguous range per block
//   tma     cp.async.bulk (1-D TMA) ring of NST stages x 5 streams x CH doubles, mbarrier full/empty
// plus plain 1R+1W copies (kernel and cudaMemcpyAsync) as the per-size calibration of "peak".
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stream_probe stream_probe.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct dd { double hi, lo; };
__device__ __forceinline__ void dd_add_d(dd &a, const double b)
{
    const double s = __dadd_rn(a.hi, b);
    const double bb = __dsub_rn(s, a.hi);
    const double e = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(b, bb));
    a.hi = s;
    a.lo = __dadd_rn(a.lo, e);
}
__device__ __forceinline__ void unit(double2 &x, const double2 p, const double2 s, const double2 as, const double2 r0, double2 &r,
                                     const double alpha, const double omega, dd &a0, dd &a1)
{
    x.x = __dadd_rn(__dadd_rn(x.x, __dmul_rn(alpha, p.x)), __dmul_rn(omega, s.x));
    x.y = __dadd_rn(__dadd_rn(x.y, __dmul_rn(alpha, p.y)), __dmul_rn(omega, s.y));
    r.x = __dsub_rn(s.x, __dmul_rn(omega, as.x));
    r.y = __dsub_rn(s.y, __dmul_rn(omega, as.y));
    dd_add_d(a0, __fma_rn(r.y, r.y, __dmul_rn(r.x, r.x)));
    dd_add_d(a1, __fma_rn(r.y, r0.y, __dmul_rn(r.x, r0.x)));
}
__device__ __forceinline__ void finish(dd a0, dd a1, double *partials)
{
    // cheap stand-in for the block reduction: one atomic per warp keeps the dots alive
    double v = a0.hi + a0.lo + a1.hi + a1.lo;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(partials + (blockIdx.x & 63), v);
}

struct V7 { double *X; const double *P, *S, *AS; double *R; const double *R0; };

template <int UNROLL, bool PF>
__global__ void __launch_bounds__(256) k_ls(const long long units, const V7 v, const double alpha, const double omega,
                                            const long long pf_units, double *partials)
{
    dd a0{0, 0}, a1{0, 0};
    const long long step = (long long)gridDim.x * blockDim.x;
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; q + (UNROLL - 1) * step < units; q += UNROLL * step) {
        double2 x[UNROLL], p[UNROLL], s[UNROLL], as[UNROLL], r0[UNROLL], r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long l = 2 * (q + u * step);
            x[u] = *reinterpret_cast<const double2 *>(v.X + l);
            p[u] = *reinterpret_cast<const double2 *>(v.P + l);
            s[u] = *reinterpret_cast<const double2 *>(v.S + l);
            as[u] = *reinterpret_cast<const double2 *>(v.AS + l);
            r0[u] = *reinterpret_cast<const double2 *>(v.R0 + l);
        }
        if (PF && (threadIdx.x & 7) == 0) {           // one prefetch per 128-byte line
            const long long l = 2 * (q + pf_units);
            if (l < 2 * units) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(v.X + l));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(v.P + l));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(v.S + l));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(v.AS + l));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(v.R0 + l));
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long l = 2 * (q + u * step);
            unit(x[u], p[u], s[u], as[u], r0[u], r[u], alpha, omega, a0, a1);
            *reinterpret_cast<double2 *>(v.X + l) = x[u];
            *reinterpret_cast<double2 *>(v.R + l) = r[u];
        }
    }
    for (; q < units; q += step) {
        const long long l = 2 * q;
        double2 x = *reinterpret_cast<const double2 *>(v.X + l), r;
        unit(x, *reinterpret_cast<const double2 *>(v.P + l), *reinterpret_cast<const double2 *>(v.S + l),
             *reinterpret_cast<const double2 *>(v.AS + l), *reinterpret_cast<const double2 *>(v.R0 + l), r, alpha, omega, a0, a1);
        *reinterpret_cast<double2 *>(v.X + l) = x;
        *reinterpret_cast<double2 *>(v.R + l) = r;
    }
    finish(a0, a1, partials);
}

__global__ void __launch_bounds__(256) k_blk(const long long units, const V7 v, const double alpha, const double omega, double *partials)
{
    dd a0{0, 0}, a1{0, 0};
    const long long per = (units + gridDim.x - 1) / gridDim.x;
    const long long b = blockIdx.x * per, e = min(units, b + per);
    for (long long q = b + threadIdx.x; q < e; q += blockDim.x) {
        const long long l = 2 * q;
        double2 x = *reinterpret_cast<const double2 *>(v.X + l), r;
        unit(x, *reinterpret_cast<const double2 *>(v.P + l), *reinterpret_cast<const double2 *>(v.S + l),
             *reinterpret_cast<const double2 *>(v.AS + l), *reinterpret_cast<const double2 *>(v.R0 + l), r, alpha, omega, a0, a1);
        *reinterpret_cast<double2 *>(v.X + l) = x;
        *reinterpret_cast<double2 *>(v.R + l) = r;
    }
    finish(a0, a1, partials);
}

// ---- 1-D TMA ring ----
__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned long long *b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}

// chunk c (CH doubles of every stream) is processed by block c % gridDim.x
template <int NST, int CH>
__global__ void __launch_bounds__(256) k_tma(const long long nchunks, const V7 v, const double alpha, const double omega, double *partials)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full[NST], empty[NST];
    constexpr int STB = 5 * CH * 8;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const long long c0 = blockIdx.x, cs = gridDim.x;
    const long long mine = (nchunks > c0) ? (nchunks - c0 + cs - 1) / cs : 0;
    auto issue = [&](long long i) {
        const int s = (int)(i % NST);
        unsigned char *st = smem + s * STB;
        const long long off = (c0 + i * cs) * CH;
        mbar_expect(full + s, STB);
        bulk_load(st, v.X + off, CH * 8, full + s);
        bulk_load(st + CH * 8, v.P + off, CH * 8, full + s);
        bulk_load(st + 2 * CH * 8, v.S + off, CH * 8, full + s);
        bulk_load(st + 3 * CH * 8, v.AS + off, CH * 8, full + s);
        bulk_load(st + 4 * CH * 8, v.R0 + off, CH * 8, full + s);
    };
    if (tid == 0) for (long long i = 0; i < min((long long)NST, mine); ++i) issue(i);
    dd a0{0, 0}, a1{0, 0};
    for (long long i = 0; i < mine; ++i) {
        const int s = (int)(i % NST);
        const unsigned ph = (unsigned)((i / NST) & 1);
        mbar_wait(full + s, ph);
        const double *st = reinterpret_cast<const double *>(smem + s * STB);
        const long long off = (c0 + i * cs) * CH;
        constexpr int UPT = CH / 2 / 256;                  // units per thread
        double2 x[UPT], p[UPT], sv[UPT], as[UPT], r0[UPT], r[UPT];
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
            const int e = 2 * (tid + u * 256);
            x[u] = *reinterpret_cast<const double2 *>(st + e);
            p[u] = *reinterpret_cast<const double2 *>(st + CH + e);
            sv[u] = *reinterpret_cast<const double2 *>(st + 2 * CH + e);
            as[u] = *reinterpret_cast<const double2 *>(st + 3 * CH + e);
            r0[u] = *reinterpret_cast<const double2 *>(st + 4 * CH + e);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty + s);
        if (tid == 0 && i + NST < mine) {                  // re-arm the stage once all 8 warps have read it
            mbar_wait(empty + s, ph);
            issue(i + NST);
        }
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
            const int e = 2 * (tid + u * 256);
            unit(x[u], p[u], sv[u], as[u], r0[u], r[u], alpha, omega, a0, a1);
            *reinterpret_cast<double2 *>(v.X + off + e) = x[u];
            *reinterpret_cast<double2 *>(v.R + off + e) = r[u];
        }
    }
    finish(a0, a1, partials);
}

__global__ void __launch_bounds__(256) k_copy(const long long units, const double *__restrict__ a, double *__restrict__ b)
{
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < units; q += step)
        *reinterpret_cast<double2 *>(b + 2 * q) = *reinterpret_cast<const double2 *>(a + 2 * q);
}
__global__ void k_init(double *p, long long n, double v)
{
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) p[q] = v;
}

template <typename F>
static double timeit(F f, int warm = 2, int reps = 8)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < warm; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char **argv)
{
    const double peak = argc > 1 ? atof(argv[1]) : 6544.7;
    double *partials;
    CK(cudaMalloc(&partials, 64 * sizeof(double)));
    CK(cudaMemset(partials, 0, 64 * sizeof(double)));
    CK(cudaFuncSetAttribute(k_tma<4, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 5 * 512 * 8));
    CK(cudaFuncSetAttribute(k_tma<3, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 5 * 1024 * 8));
    CK(cudaFuncSetAttribute(k_tma<5, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 5 * 512 * 8));
    CK(cudaFuncSetAttribute(k_tma<2, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 5 * 1024 * 8));
    const long long sizes[2] = {52428800LL, 419430400LL};     // plate(256), plate(512) unknowns
    for (int si = 0; si < 2; ++si) {
        const long long n = sizes[si];
        for (int layout = 0; layout < 2; ++layout) {
            // layout 0: 7 vectors back to back (stride n + 2, as in the solver); layout 1: stride padded by 1 MiB + 4 KiB
            const long long stride = layout == 0 ? n + 2 : n + (1 << 17) + 512;
            double *base;
            CK(cudaMalloc(&base, (size_t)stride * 7 * sizeof(double)));
            k_init<<<148 * 8, 256>>>(base, stride * 7, 0.5);
            CK(cudaDeviceSynchronize());
            V7 v{base, base + stride, base + 2 * stride, base + 3 * stride, base + 4 * stride, base + 5 * stride};
            const long long units = n / 2;
            const double gb = 56.0 * n / 1e9, gbc = 16.0 * n / 1e9;
            auto rep = [&](const char *name, double ms, double g) {
                printf("n=%lld layout=%d %-28s %8.4f ms  %7.1f GB/s  %.3f of %.1f\n", n, layout, name, ms, g / ms * 1e3, g / ms * 1e3 / peak, peak);
                fflush(stdout);
            };
            rep("copy_kernel(1R1W)", timeit([&] { k_copy<<<148 * 8, 256>>>(units, v.P, base + 6 * stride); }), gbc);
            rep("cudaMemcpyAsync D2D", timeit([&] { CK(cudaMemcpyAsync(base + 6 * stride, v.P, n * 8, cudaMemcpyDeviceToDevice)); }), gbc);
            rep("ls grid 148x8", timeit([&] { k_ls<1, false><<<148 * 8, 256>>>(units, v, 1e-30, 1.0, 0, partials); }), gb);
            rep("ls grid 148x4", timeit([&] { k_ls<1, false><<<148 * 4, 256>>>(units, v, 1e-30, 1.0, 0, partials); }), gb);
            rep("ls_u2 grid 148x4", timeit([&] { k_ls<2, false><<<148 * 4, 256>>>(units, v, 1e-30, 1.0, 0, partials); }), gb);
            rep("ls_u2 grid 148x8", timeit([&] { k_ls<2, false><<<148 * 8, 256>>>(units, v, 1e-30, 1.0, 0, partials); }), gb);
            rep("ls_u4 grid 148x4", timeit([&] { k_ls<4, false><<<148 * 4, 256>>>(units, v, 1e-30, 1.0, 0, partials); }), gb);
            for (long long mb : {1LL, 4LL, 16LL}) {
                char nm[64];
                snprintf(nm, sizeof nm, "ls_pf %lld MiB ahead 148x8", mb);
                const long long pfu = mb * (1 << 20) / 16;
                rep(nm, timeit([&] { k_ls<1, true><<<148 * 8, 256>>>(units, v, 1e-30, 1.0, pfu, partials); }), gb);
            }
            rep("blk 148x8", timeit([&] { k_blk<<<148 * 8, 256>>>(units, v, 1e-30, 1.0, partials); }), gb);
            rep("tma NST4 CH512 148x2", timeit([&] { k_tma<4, 512><<<148 * 2, 256, 4 * 5 * 512 * 8>>>(n / 512, v, 1e-30, 1.0, partials); }), gb);
            rep("tma NST5 CH512 148x2", timeit([&] { k_tma<5, 512><<<148 * 2, 256, 5 * 5 * 512 * 8>>>(n / 512, v, 1e-30, 1.0, partials); }), gb);
            rep("tma NST3 CH1024 148x1", timeit([&] { k_tma<3, 1024><<<148, 256, 3 * 5 * 1024 * 8>>>(n / 1024, v, 1e-30, 1.0, partials); }), gb);
            rep("tma NST2 CH1024 148x2", timeit([&] { k_tma<2, 1024><<<148 * 2, 256, 2 * 5 * 1024 * 8>>>(n / 1024, v, 1e-30, 1.0, partials); }), gb);
            rep("tma NST4 CH512 148x3", timeit([&] { k_tma<4, 512><<<148 * 3, 256, 4 * 5 * 512 * 8>>>(n / 512, v, 1e-30, 1.0, partials); }), gb);
            CK(cudaFree(base));
        }
    }
    // clocks after the run (sustained-load context)
    return 0;
}
