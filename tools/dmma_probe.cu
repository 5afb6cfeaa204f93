// Probe: what limits DMMA issue on sm_100a?  (A) register-only DMMA with 32 accumulators at 1..4 warps
// per SM sub-partition; (B) the Gram kernel's MMA loop in isolation (fragments from shared memory,
// no barriers, no Theta build).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NA, int NB>
__global__ void __launch_bounds__(512) k_reg(double* out, int iters) {
    double a[NA], b[NB], acc[NA][NB][2];
    for (int i = 0; i < NA; ++i) a[i] = 1e-3 * (threadIdx.x + i);
    for (int j = 0; j < NB; ++j) b[j] = 1e-3 * (threadIdx.x - j);
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    double s = 0;
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[0] = s;
}

// (B) 8 warps, 64x32 patch each, 128x128 block, fragments from a [16][132] sample-major tile pair
template <int WARPS, int NA, int NB>
__global__ void __launch_bounds__(WARPS * 32) k_smem(double* out, int stages) {
    constexpr int LDT = 132, KS = 16;
    __shared__ double tile[1][2 * KS * LDT];
    for (int e = threadIdx.x; e < 2 * KS * LDT; e += blockDim.x) (&tile[0][0])[e] = 1e-3 * e;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int wm = warp & 1, wn = (warp >> 1) & 3;
    double acc[NA][NB][2];
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0;
    for (int st = 0; st < stages; ++st) {
        const double* tI = tile[0] + ((st & 1) ? 0 : 0);
        const double* tJ = tI + KS * LDT;
        const double* pa = tI + t * LDT + wm * (8 * NA) + g;
        const double* pb = tJ + t * LDT + wn * (8 * NB) + g;
#pragma unroll
        for (int kk = 0; kk < KS / 4; ++kk) {
            double a[NA], b[NB];
#pragma unroll
            for (int i = 0; i < NA; ++i) a[i] = pa[kk * 4 * LDT + 8 * i];
#pragma unroll
            for (int j = 0; j < NB; ++j) b[j] = pb[kk * 4 * LDT + 8 * j];
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[0] = s;
}


// (C) as (B) plus ND independent DMULs per k-step (MODE 0: spread, one group per k-step; MODE 1: all 4*ND at stage start)
template <int ND, int MODE>
__global__ void __launch_bounds__(256) k_mix(double* out, int stages, double mulv) {
    constexpr int LDT = 132, KS = 16, NA = 8, NB = 4;
    __shared__ double tile[2 * KS * LDT];
    for (int e = threadIdx.x; e < 2 * KS * LDT; e += blockDim.x) tile[e] = 1e-3 * e;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int wm = warp & 1, wn = (warp >> 1) & 3;
    double acc[NA][NB][2];
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0;
    double prod[ND > 0 ? ND : 1];
    for (int d = 0; d < ND; ++d) prod[d] = 1.0 + 1e-9 * (threadIdx.x + d);
    for (int st = 0; st < stages; ++st) {
        const double* tI = tile;
        const double* tJ = tI + KS * LDT;
        const double* pa = tI + t * LDT + wm * (8 * NA) + g;
        const double* pb = tJ + t * LDT + wn * (8 * NB) + g;
        if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int d = 0; d < ND; ++d) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(prod[d]) : "d"(mulv));
        }
#pragma unroll
        for (int kk = 0; kk < KS / 4; ++kk) {
            if (MODE == 0) {
#pragma unroll
                for (int d = 0; d < ND; ++d) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(prod[d]) : "d"(mulv));
            }
            double a[NA], b[NB];
#pragma unroll
            for (int i = 0; i < NA; ++i) a[i] = pa[kk * 4 * LDT + 8 * i];
#pragma unroll
            for (int j = 0; j < NB; ++j) b[j] = pb[kk * 4 * LDT + 8 * j];
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) s += acc[i][j][0] + acc[i][j][1];
    for (int d = 0; d < ND; ++d) s += prod[d];
    if (s == 12345.678) out[0] = s;
}


// (D) as (B) plus one CTA-wide synchronisation per stage: SYNC 1 = __syncthreads, 2 = mbarrier arrive (one lane per
// warp) + try_wait.parity, 3 = mbarrier arrive + test_wait spin
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int SYNC>
__global__ void __launch_bounds__(256) k_sync(double* out, int stages) {
    constexpr int LDT = 132, KS = 16, NA = 8, NB = 4;
    __shared__ double tile[2 * KS * LDT];
    __shared__ unsigned long long bar[2];
    for (int e = threadIdx.x; e < 2 * KS * LDT; e += blockDim.x) tile[e] = 1e-3 * e;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[0])), "r"(8));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[1])), "r"(8));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int wm = warp & 1, wn = (warp >> 1) & 3;
    double acc[NA][NB][2];
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0;
    for (int st = 0; st < stages; ++st) {
        const double* pa = tile + t * LDT + wm * (8 * NA) + g;
        const double* pb = tile + KS * LDT + t * LDT + wn * (8 * NB) + g;
#pragma unroll
        for (int kk = 0; kk < KS / 4; ++kk) {
            double a[NA], b[NB];
#pragma unroll
            for (int i = 0; i < NA; ++i) a[i] = pa[kk * 4 * LDT + 8 * i];
#pragma unroll
            for (int j = 0; j < NB; ++j) b[j] = pb[kk * 4 * LDT + 8 * j];
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (SYNC == 1) __syncthreads();
        if (SYNC == 2 || SYNC == 3) {
            unsigned long long* b = &bar[st & 1];
            const unsigned parity = (st >> 1) & 1;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
            unsigned done;
            do {
                if (SYNC == 2)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
                else
                    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
            } while (!done);
        }
    }
    double s = 0;
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[0] = s;
}


// (E) as (D, mbarrier) plus the Theta build traffic: per k-step NL LDS.64 from a raw region, NL/2 DMUL, NL/2 STS.64
// into a scratch tile.  MODE bit0: loads, bit1: muls, bit2: stores.
template <int NL, int MODE>
__global__ void __launch_bounds__(256) k_build(double* out, int stages, int n) {
    constexpr int LDT = 132, KS = 16, NA = 8, NB = 4;
    extern __shared__ double sm[];
    double* tile = sm;                       // 2 * KS * LDT
    double* scratch = tile + 2 * KS * LDT;   // 2 * KS * LDT
    double* raw = scratch + 2 * KS * LDT;    // KS * 2n
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(raw + KS * 2 * n);
    for (int e = threadIdx.x; e < 4 * KS * LDT + KS * 2 * n; e += blockDim.x) sm[e] = 1.0 + 1e-6 * e;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[0])), "r"(8));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[1])), "r"(8));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int wm = warp & 1, wn = (warp >> 1) & 3;
    const int col = threadIdx.x & 127, tl = threadIdx.x >> 7;
    const int o0 = col & 63, o1 = (col * 7) & 63;
    double acc[NA][NB][2];
    long long xacc = 0; double cmul = 1.0000001 + 1e-9 * n;
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0;
    for (int st = 0; st < stages; ++st) {
        const double* pa = tile + t * LDT + wm * (8 * NA) + g;
        const double* pb = tile + KS * LDT + t * LDT + wn * (8 * NB) + g;
        double* dst = scratch + tl * KS * LDT + col;
#pragma unroll
        for (int kk = 0; kk < KS / 4; ++kk) {
            double f0[NL / 2], f1[NL / 2];
#pragma unroll
            for (int q = 0; q < NL / 2; ++q) {
                const int sidx = kk * (NL / 2) + q;
                if (MODE & 32) {
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(f0[q]) : "r"(s32(raw + o0 + sidx * n)));
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(f1[q]) : "r"(s32(raw + o1 + sidx * n)));
                } else {
                f0[q] = (MODE & 1) ? raw[o0 + sidx * n] : 1.0 + q;
                f1[q] = (MODE & 1) ? raw[o1 + sidx * n] : 2.0 + q + st;
                }
            }
            double a[NA], b[NB];
#pragma unroll
            for (int i = 0; i < NA; ++i) a[i] = pa[kk * 4 * LDT + 8 * i];
#pragma unroll
            for (int j = 0; j < NB; ++j) b[j] = pb[kk * 4 * LDT + 8 * j];
            if (MODE & 128) {
                // all products of the k-step in ONE asm block (back-to-back DMULs), then the stores
                double p0, p1, p2, p3;
                asm volatile("mul.rn.f64 %0, %4, %8;\n\tmul.rn.f64 %1, %5, %9;\n\tmul.rn.f64 %2, %6, %10;\n\tmul.rn.f64 %3, %7, %11;"
                             : "=d"(p0), "=d"(p1), "=d"(p2), "=d"(p3)
                             : "d"(f0[0]), "d"(f0[1]), "d"(f0[2]), "d"(f0[3]), "d"(f1[0]), "d"(f1[1]), "d"(f1[2]), "d"(f1[3]));
                dst[(kk * 4 + 0) * LDT] = p0; dst[(kk * 4 + 1) * LDT] = p1; dst[(kk * 4 + 2) * LDT] = p2; dst[(kk * 4 + 3) * LDT] = p3;
            } else if (!(MODE & 8)) {
#pragma unroll
            for (int q = 0; q < NL / 2; ++q) {
                double p = f0[q];
                if (MODE & 64) {
                    // loads are consumed by integer ops only; the multiply works on register constants
                    xacc ^= __double_as_longlong(f0[q]) + __double_as_longlong(f1[q]);
                    p = 1.0 + q;
                    if (MODE & 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(p) : "d"(cmul));
                } else if (MODE & 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(p) : "d"(f1[q])); else p = f1[q];
                if (MODE & 4) dst[(kk * (NL / 2) + q) * LDT] = p; else if (p == 12345.678) out[1] = p;
            }
            }
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                    if ((MODE & 16) && j == 3 && (i & 1) == 1 && i / 2 < NL / 2) {  // sprinkle: one product every 8 DMMAs
                        const int q = i / 2;
                        double p = f0[q];
                        asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(p) : "d"(f1[q]));
                        dst[(kk * (NL / 2) + q) * LDT] = p;
                    }
                }
            if ((MODE & 8) && !(MODE & 16)) {
#pragma unroll
            for (int q = 0; q < NL / 2; ++q) {
                double p = f0[q];
                asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(p) : "d"(f1[q]));
                dst[(kk * (NL / 2) + q) * LDT] = p;
            }
            }
        }
        {
            unsigned long long* b = &bar[st & 1];
            const unsigned parity = (st >> 1) & 1;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
            unsigned done;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
            } while (!done);
        }
    }
    double s = 0;
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678 || xacc == 77) out[0] = s;
}

// (E16) 16 warps of 32x32: as (D, mbarrier) plus the Theta build traffic: per k-step NL LDS.64 from a raw region, NL/2 DMUL, NL/2 STS.64
// into a scratch tile.  MODE bit0: loads, bit1: muls, bit2: stores.
template <int NL, int MODE>
__global__ void __launch_bounds__(512) k_build16(double* out, int stages, int n) {
    constexpr int LDT = 132, KS = 16, NA = 4, NB = 4;
    extern __shared__ double sm[];
    double* tile = sm;                       // 2 * KS * LDT
    double* scratch = tile + 2 * KS * LDT;   // 2 * KS * LDT
    double* raw = scratch + 2 * KS * LDT;    // KS * 2n
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(raw + KS * 2 * n);
    for (int e = threadIdx.x; e < 4 * KS * LDT + KS * 2 * n; e += blockDim.x) sm[e] = 1.0 + 1e-6 * e;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[0])), "r"(16));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[1])), "r"(16));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = (warp >> 2) & 3;
    const int col = threadIdx.x & 127, tl = (threadIdx.x >> 7) & 1, half = threadIdx.x >> 8;
    const int o0 = col & 63, o1 = (col * 7) & 63;
    double acc[NA][NB][2];
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0;
    for (int st = 0; st < stages; ++st) {
        const double* pa = tile + t * LDT + wm * (8 * NA) + g;
        const double* pb = tile + KS * LDT + t * LDT + wn * (8 * NB) + g;
        double* dst = scratch + tl * KS * LDT + col;
#pragma unroll
        for (int kk = 0; kk < KS / 4; ++kk) {
            double f0[NL / 2], f1[NL / 2];
#pragma unroll
            for (int q = 0; q < NL / 2; ++q) {
                const int sidx = half * 8 + kk * (NL / 2) + q;
                if (MODE & 32) {
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(f0[q]) : "r"(s32(raw + o0 + sidx * n)));
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(f1[q]) : "r"(s32(raw + o1 + sidx * n)));
                } else {
                f0[q] = (MODE & 1) ? raw[o0 + sidx * n] : 1.0 + q;
                f1[q] = (MODE & 1) ? raw[o1 + sidx * n] : 2.0 + q + st;
                }
            }
            double a[NA], b[NB];
#pragma unroll
            for (int i = 0; i < NA; ++i) a[i] = pa[kk * 4 * LDT + 8 * i];
#pragma unroll
            for (int j = 0; j < NB; ++j) b[j] = pb[kk * 4 * LDT + 8 * j];
            if (!(MODE & 8)) {
#pragma unroll
            for (int q = 0; q < NL / 2; ++q) {
                double p = f0[q];
                if (MODE & 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(p) : "d"(f1[q])); else p = f1[q];
                if (MODE & 4) dst[(half * 8 + kk * (NL / 2) + q) * LDT] = p; else if (p == 12345.678) out[1] = p;
            }
            }
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                    if ((MODE & 16) && j == 3 && (i & 1) == 1 && i / 2 < NL / 2) {  // sprinkle: one product every 8 DMMAs
                        const int q = i / 2;
                        double p = f0[q];
                        asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(p) : "d"(f1[q]));
                        dst[(half * 8 + kk * (NL / 2) + q) * LDT] = p;
                    }
                }
            if ((MODE & 8) && !(MODE & 16)) {
#pragma unroll
            for (int q = 0; q < NL / 2; ++q) {
                double p = f0[q];
                asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(p) : "d"(f1[q]));
                dst[(half * 8 + kk * (NL / 2) + q) * LDT] = p;
            }
            }
        }
        {
            unsigned long long* b = &bar[st & 1];
            const unsigned parity = (st >> 1) & 1;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
            unsigned done;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
            } while (!done);
        }
    }
    double s = 0;
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[0] = s;
}



// (F) build phase for the whole stage FIRST (own basic block, closed by a warp barrier), then the 4 MMA k-steps
template <int CHUNK>
__global__ void __launch_bounds__(256) k_phase(double* out, int stages, int n) {
    constexpr int LDT = 132, KS = 16, NA = 8, NB = 4;
    extern __shared__ double sm[];
    double* tile = sm;
    double* scratch = tile + 2 * KS * LDT;
    double* raw = scratch + 2 * KS * LDT;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(raw + KS * 2 * n);
    for (int e = threadIdx.x; e < 4 * KS * LDT + KS * 2 * n; e += blockDim.x) sm[e] = 1.0 + 1e-6 * e;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[0])), "r"(8));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[1])), "r"(8));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int wm = warp & 1, wn = (warp >> 1) & 3;
    const int col = threadIdx.x & 127, tl = threadIdx.x >> 7;
    const int o0 = col & 63, o1 = (col * 7) & 63;
    double acc[NA][NB][2];
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0;
    for (int st = 0; st < stages; ++st) {
        const double* pa = tile + t * LDT + wm * (8 * NA) + g;
        const double* pb = tile + KS * LDT + t * LDT + wn * (8 * NB) + g;
        double* dst = scratch + tl * KS * LDT + col;
#pragma unroll
        for (int c0 = 0; c0 < KS; c0 += CHUNK) {
            double f0[CHUNK], f1[CHUNK];
#pragma unroll
            for (int q = 0; q < CHUNK; ++q) {
                f0[q] = raw[o0 + (c0 + q) * n];
                f1[q] = raw[o1 + (c0 + q) * n];
            }
#pragma unroll
            for (int q = 0; q < CHUNK; ++q) dst[(c0 + q) * LDT] = f0[q] * f1[q];
        }
        asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
#pragma unroll
        for (int kk = 0; kk < KS / 4; ++kk) {
            double a[NA], b[NB];
#pragma unroll
            for (int i = 0; i < NA; ++i) a[i] = pa[kk * 4 * LDT + 8 * i];
#pragma unroll
            for (int j = 0; j < NB; ++j) b[j] = pb[kk * 4 * LDT + 8 * j];
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        {
            unsigned long long* b = &bar[st & 1];
            const unsigned parity = (st >> 1) & 1;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
            unsigned done;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
            } while (!done);
        }
    }
    double s = 0;
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[0] = s;
}

// (G) Theta build on the tensor pipe itself: a product tile u_r(s) * v_j(s) (4 samples x 2 second factors x 8 first
// factors = 64 products) is ONE DMMA with a one-hot A operand (A[(s,r)][k] = u_r(s) if k == s else 0, B[k][j] = v_j(s_k),
// C = 0: every output is a single correctly rounded product).  NOPS ops per k-step replace the 4 DMULs of k_build.
// TAB 0: per-lane offsets in registers, 1: one LDS.128 per op from a per-warp table in shared memory.
template <int NOPS, int TAB>
__global__ void __launch_bounds__(256) k_buildmma(double* out, int stages, int n) {
    constexpr int LDT = 132, KS = 16, NA = 8, NB = 4;
    extern __shared__ double sm[];
    double* tile = sm;                       // 2 * KS * LDT
    double* scratch = tile + 2 * KS * LDT;   // 2 * KS * LDT
    double* raw = scratch + 2 * KS * LDT;    // KS * 2n
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(raw + KS * 2 * n);
    int4* tab = reinterpret_cast<int4*>(bar + 2);  // [8 warps][4 kk][NOPS][32 lanes]
    for (int e = threadIdx.x; e < 4 * KS * LDT + KS * 2 * n; e += blockDim.x) sm[e] = 1.0 + 1e-6 * e;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[0])), "r"(8));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[1])), "r"(8));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int wm = warp & 1, wn = (warp >> 1) & 3;
    const int srow = g >> 1, r = g & 1;  // row i = g of the A operand: sample srow, second factor r
    const bool hot = (t == srow);
    int offA[NOPS], offB[NOPS], dstc[NOPS];
#pragma unroll
    for (int o = 0; o < NOPS; ++o) {
        const int q = (warp * 5 + o * 11 + r * 3) & 63;           // second factor of row r
        const int p = (warp * 7 + o * 16 + g) & 63;               // first factor j = g (B operand column)
        offA[o] = q + srow * n;
        offB[o] = p + t * n;
        dstc[o] = (warp >> 2) * KS * LDT + ((warp & 3) * 32 + ((o & 1) * 16 + r * 8 + 2 * t)) + srow * LDT;
        if (TAB)
            for (int kk = 0; kk < 4; ++kk)
                tab[((warp * 4 + kk) * NOPS + o) * 32 + lane] =
                    make_int4((offA[o] + kk * 4 * n) * 8, (offB[o] + kk * 4 * n) * 8, (dstc[o] + kk * 4 * LDT) * 8, 0);
    }
    __syncthreads();
    const unsigned raw32 = s32(raw), scr32 = s32(scratch);
    double acc[NA][NB][2];
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0;
    for (int st = 0; st < stages; ++st) {
        const double* pa = tile + t * LDT + wm * (8 * NA) + g;
        const double* pb = tile + KS * LDT + t * LDT + wn * (8 * NB) + g;
#pragma unroll
        for (int kk = 0; kk < KS / 4; ++kk) {
            double ua[NOPS], vb[NOPS];
            unsigned dsta[NOPS];
#pragma unroll
            for (int o = 0; o < NOPS; ++o) {
                if (TAB) {
                    const int4 e = tab[((warp * 4 + kk) * NOPS + o) * 32 + lane];
                    ua[o] = 0.0;
                    if (hot) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(ua[o]) : "r"(raw32 + e.x));
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(vb[o]) : "r"(raw32 + e.y));
                    dsta[o] = scr32 + e.z;
                } else {
                    ua[o] = 0.0;
                    if (hot) ua[o] = raw[offA[o] + kk * 4 * n];
                    vb[o] = raw[offB[o] + kk * 4 * n];
                    dsta[o] = scr32 + (dstc[o] + kk * 4 * LDT) * 8;
                }
            }
            double a[NA], b[NB];
#pragma unroll
            for (int i = 0; i < NA; ++i) a[i] = pa[kk * 4 * LDT + 8 * i];
#pragma unroll
            for (int j = 0; j < NB; ++j) b[j] = pb[kk * 4 * LDT + 8 * j];
#pragma unroll
            for (int o = 0; o < NOPS; ++o) {
                double d0 = 0.0, d1 = 0.0;
                dmma884(d0, d1, ua[o], vb[o]);
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(dsta[o]), "d"(d0), "d"(d1) : "memory");
            }
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        {
            unsigned long long* b = &bar[st & 1];
            const unsigned parity = (st >> 1) & 1;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
            unsigned done;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
            } while (!done);
        }
    }
    double s = 0;
    for (int i = 0; i < NA; ++i) for (int j = 0; j < NB; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[0] = s;
}

template <class F> float time_ms(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 1024));
    printf("{");
    const int iters = 2000;
    for (int warps : {4, 8, 12, 16}) {
        float ms = time_ms([&] { k_reg<4, 4><<<sms, warps * 32>>>(out, iters); });
        printf("\"reg_4x4_w%d\": %.2f, ", warps, 512.0 * 16 * iters * sms * warps / ms * 1e-9);
    }
    for (int warps : {4, 8}) {
        float ms = time_ms([&] { k_reg<8, 4><<<sms, warps * 32>>>(out, iters); });
        printf("\"reg_8x4_w%d\": %.2f, ", warps, 512.0 * 32 * iters * sms * warps / ms * 1e-9);
    }
    const int stages = 4000;
    { float ms = time_ms([&] { k_smem<8, 8, 4><<<sms, 256>>>(out, stages); });
      printf("\"smem_8w_64x32\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
    { float ms = time_ms([&] { k_smem<16, 4, 4><<<sms, 512>>>(out, stages); });
      printf("\"smem_16w_32x32\": %.2f, ", 512.0 * 16 * 4 * stages * sms * 16 / ms * 1e-9); }
    { float ms = time_ms([&] { k_smem<4, 8, 4><<<sms, 128>>>(out, stages); });
      printf("\"smem_4w_64x32\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 4 / ms * 1e-9); }
    { float ms = time_ms([&] { k_smem<4, 8, 4><<<sms * 2, 128>>>(out, stages); });
      printf("\"smem_4w_64x32_2cta\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 2 * 4 / ms * 1e-9); }
    { float ms = time_ms([&] { k_mix<1, 0><<<sms, 256>>>(out, stages, 1.0000001); }); printf("\"mix_1_spread\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
    { float ms = time_ms([&] { k_mix<2, 0><<<sms, 256>>>(out, stages, 1.0000001); }); printf("\"mix_2_spread\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
    { float ms = time_ms([&] { k_mix<4, 0><<<sms, 256>>>(out, stages, 1.0000001); }); printf("\"mix_4_spread\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
    { float ms = time_ms([&] { k_mix<8, 0><<<sms, 256>>>(out, stages, 1.0000001); }); printf("\"mix_8_spread\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
    { float ms = time_ms([&] { k_mix<4, 1><<<sms, 256>>>(out, stages, 1.0000001); }); printf("\"mix_16_front\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
    { float ms = time_ms([&] { k_mix<8, 1><<<sms, 256>>>(out, stages, 1.0000001); }); printf("\"mix_32_front\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
    { float ms = time_ms([&] { k_sync<0><<<sms, 256>>>(out, stages); }); printf("\"sync_none\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
    { float ms = time_ms([&] { k_sync<1><<<sms, 256>>>(out, stages); }); printf("\"sync_bar\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
    { float ms = time_ms([&] { k_sync<2><<<sms, 256>>>(out, stages); }); printf("\"sync_mbar_try\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
    { float ms = time_ms([&] { k_sync<3><<<sms, 256>>>(out, stages); }); printf("\"sync_mbar_test\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
    {
        const int n = 64; const size_t smem = (4 * 16 * 132 + 16 * 2 * n) * 8 + 64;
#define RUNB(NL, MODE, name) { auto k = k_build<NL, MODE>; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        float ms = time_ms([&] { k<<<sms, 256, smem>>>(out, stages, n); }); printf("\"" name "\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
        RUNB(8, 0, "build_none") RUNB(8, 1, "build_lds") RUNB(8, 3, "build_lds_mul") RUNB(8, 4, "build_sts") RUNB(8, 6, "build_mul_sts") RUNB(8, 7, "build_all") RUNB(8, 15, "build_all_after") RUNB(8, 31, "build_all_sprinkle") { auto k = k_build16<4, 7>; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          float ms = time_ms([&] { k<<<sms, 512, smem>>>(out, stages, n); }); printf("\"build16_all\": %.2f, ", 512.0 * 16 * 4 * stages * sms * 16 / ms * 1e-9); }
        { auto k = k_build16<4, 0>; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          float ms = time_ms([&] { k<<<sms, 512, smem>>>(out, stages, n); }); printf("\"build16_none\": %.2f, ", 512.0 * 16 * 4 * stages * sms * 16 / ms * 1e-9); }
        RUNB(8, 71, "build_lds_int_mulconst_sts") RUNB(8, 69, "build_lds_int_sts") RUNB(8, 135, "build_all_grouped") { auto k = k_phase<8>; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          float ms = time_ms([&] { k<<<sms, 256, smem>>>(out, stages, n); }); printf("\"phase_chunk8\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
        { auto k = k_phase<16>; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          float ms = time_ms([&] { k<<<sms, 256, smem>>>(out, stages, n); }); printf("\"phase_chunk16\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
        RUNB(8, 47, "build_vol_after") RUNB(8, 63, "build_vol_sprinkle")
        {
            const size_t smem2 = smem + 8 * 4 * 3 * 32 * 16 + 64;
#define RUNM(NOPS, TAB, name) { auto k = k_buildmma<NOPS, TAB>; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
        float ms = time_ms([&] { k<<<sms, 256, smem2>>>(out, stages, n); }); printf("\"" name "\": %.2f, ", 512.0 * 32 * 4 * stages * sms * 8 / ms * 1e-9); }
            RUNM(2, 0, "buildmma_2ops_reg") RUNM(2, 1, "buildmma_2ops_tab") RUNM(3, 0, "buildmma_3ops_reg") RUNM(3, 1, "buildmma_3ops_tab")
            RUNM(1, 0, "buildmma_1op_reg")
        }
    }
    printf("\"sms\": %d}\n", sms);
    return 0;
}
