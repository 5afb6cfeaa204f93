// Single-warp latency / issue-rate probe for the instructions on the serial chains of the Cholesky kernels
// (csrc/chol.cu): FP64 FMA, 64-bit shuffle, shared-memory round trip, DMMA.8x8x4, reciprocal estimate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/lat_probe.cu -o tools/bin/lat_probe
//   tools/bin/lat_probe          -> one JSON object, cycles per operation (clock64 around 4096 operations, one warp)
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e = (x);                                                        \
        if (e != cudaSuccess) {                                                     \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                 \
            return 1;                                                               \
        }                                                                           \
    } while (0)

constexpr int kN = 4096;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: dependent DFMA chain           1: NI independent DFMA chains
//      2: dependent 64-bit shuffle chain  3: NI independent 64-bit shuffles
//      4: dependent DMMA chain            5: NI independent DMMA chains
//      6: STS -> syncwarp -> LDS -> DFMA chain (a value travelling through shared memory)
//      7: rcp.approx.ftz.f64 + 3 FMA chain (fast_rcp)   8: dependent DMUL->DFMA pairs with a broadcast LDS operand
//      9: NI independent LDS (uniform address) + DFMA
template <int MODE, int NI>
__global__ void probe(double* out, long long* cycles, double seed) {
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = seed + lane, sm[32 + lane] = 1e-9 * lane;
    __syncwarp();
    double acc[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i] = seed + i + lane;
    double c1[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) c1[i] = 0.0;
    const double m = 1.0 + 1e-12 * seed, a = 1e-9;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kN / NI / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                if (MODE == 0 || MODE == 1) acc[i] = fma(acc[i], m, a);
                if (MODE == 2 || MODE == 3) acc[i] = __shfl_sync(0xffffffffu, acc[i], (lane + 1) & 31);
                if (MODE == 4 || MODE == 5) dmma884(acc[i], c1[i], m, a);
                if (MODE == 6) {
                    sm[lane] = acc[i];
                    __syncwarp();
                    acc[i] = fma(sm[(lane + 1) & 31], m, a);
                    __syncwarp();
                }
                if (MODE == 7) {
                    double y;
                    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(acc[i]));
                    const double e = fma(-acc[i], y, 1.0);
                    const double e2 = fma(e, e, e);
                    acc[i] = fma(y, e2, y) + 1.5;
                }
                if (MODE == 8) {
                    acc[i] *= m;
                    acc[i] = fma(-acc[i], sm[32 + ((u * NI + i) & 31)], a);
                }
                if (MODE == 9) acc[i] = fma(acc[i], sm[32 + ((u * NI + i) & 31)], a);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NI; ++i) s += acc[i] + c1[i];
    out[lane] = s;
    if (lane == 0) *cycles = t1 - t0;
}

template <int MODE, int NI>
static int run(const char* name, double* dout, long long* dcyc, int ops_per_iter) {
    long long best = 1LL << 60;
    for (int r = 0; r < 5; ++r) {
        probe<MODE, NI><<<1, 32>>>(dout, dcyc, 1.0 + r);
        long long c;
        CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
        if (c < best) best = c;
    }
    printf("\"%s\": %.2f, ", name, (double)best / (kN * (double)ops_per_iter));
    return 0;
}

int main() {
    double* dout;
    long long* dcyc;
    CK(cudaMalloc(&dout, 256));
    CK(cudaMalloc(&dcyc, 8));
    printf("{");
    if (run<0, 1>("dfma_dependent", dout, dcyc, 1)) return 1;
    if (run<1, 2>("dfma_2_chains", dout, dcyc, 1)) return 1;
    if (run<1, 4>("dfma_4_chains", dout, dcyc, 1)) return 1;
    if (run<1, 8>("dfma_8_chains", dout, dcyc, 1)) return 1;
    if (run<1, 16>("dfma_16_chains", dout, dcyc, 1)) return 1;
    if (run<2, 1>("shfl64_dependent", dout, dcyc, 1)) return 1;
    if (run<3, 8>("shfl64_8_independent", dout, dcyc, 1)) return 1;
    if (run<3, 16>("shfl64_16_independent", dout, dcyc, 1)) return 1;
    if (run<4, 1>("dmma_dependent", dout, dcyc, 1)) return 1;
    if (run<5, 2>("dmma_2_chains", dout, dcyc, 1)) return 1;
    if (run<5, 4>("dmma_4_chains", dout, dcyc, 1)) return 1;
    if (run<5, 8>("dmma_8_chains", dout, dcyc, 1)) return 1;
    if (run<6, 1>("sts_lds_dfma_roundtrip", dout, dcyc, 1)) return 1;
    if (run<7, 1>("fast_rcp_plus_add_dependent", dout, dcyc, 1)) return 1;
    if (run<8, 1>("dmul_dfma_lds_dependent_pair", dout, dcyc, 1)) return 1;
    if (run<9, 8>("lds_dfma_8_chains", dout, dcyc, 1)) return 1;
    if (run<9, 16>("lds_dfma_16_chains", dout, dcyc, 1)) return 1;
    printf("\"unit\": \"cycles per operation, one warp\"}\n");
    return 0;
}
