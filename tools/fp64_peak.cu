// FP64 calibration for the Gram-stage roofline denominator ("FP64 tensor
// peak is not in MEASURED_PEAKS.json; must be measured").
//
// Measures, on the current device:
//   * DFMA   : register-resident fma.rn.f64 chains (the non-tensor FP64 pipe)
//   * DMMA   : mma.sync.aligned.{m8n8k4,m16n8k4,m16n8k8,m16n8k16}.f64 with independent accumulators
//   * cuBLAS : cublasDgemm N^3 (N = 8192 by default) and cublasDsyrk
// Prints one JSON object on stdout.  Build: see tools/Makefile.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int ACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

// shape: 0 = m8n8k4, 1 = m16n8k4, 2 = m16n8k8, 3 = m16n8k16
template <int SHAPE, int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters) {
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1e-3 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (threadIdx.x - i);
    double c[NACC][4];
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            if constexpr (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a[0]), "d"(b[0]));
            } else if constexpr (SHAPE == 1) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                             : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            } else if constexpr (SHAPE == 2) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            } else {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (s == 12345.678) out[0] = s;
}

static float time_ms(void (*launch)(void*), void* arg, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(arg); launch(arg);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch(arg);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

struct Cfg { double* out; int iters; int grid; int block; };

template <int ACC> static void l_dfma(void* p) { Cfg* c = (Cfg*)p; k_dfma<ACC><<<c->grid, c->block>>>(c->out, c->iters, 1.0000001, 1e-9); }
template <int S, int N> static void l_dmma(void* p) { Cfg* c = (Cfg*)p; k_dmma<S, N><<<c->grid, c->block>>>(c->out, c->iters); }

int main(int argc, char** argv) {
    int N = argc > 1 ? atoi(argv[1]) : 8192;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 1024));
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);

    // DFMA: 8 accumulators / thread, various warps/SM
    for (int wps : {8, 16, 32}) {
        Cfg c{out, 20000, sms * (wps / 8), 256};
        float ms = time_ms(l_dfma<8>, &c, 5);
        double flops = 2.0 * 8 * c.iters * (double)c.grid * c.block;
        printf(", \"dfma_tflops_w%d\": %.3f", wps, flops / ms * 1e-9);
    }
    // DMMA shapes, 8 warps per CTA, CTAs/SM in {1,2,4}
    const char* names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
    const double fl[4] = {2.0 * 8 * 8 * 4, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16};
    for (int cps : {1, 2, 4}) {
        Cfg c{out, 4000, sms * cps, 256};
        float ms[4];
        ms[0] = time_ms(l_dmma<0, 8>, &c, 5);
        ms[1] = time_ms(l_dmma<1, 8>, &c, 5);
        ms[2] = time_ms(l_dmma<2, 8>, &c, 5);
        ms[3] = time_ms(l_dmma<3, 8>, &c, 5);
        for (int s = 0; s < 4; ++s) {
            double flops = fl[s] * 8 * c.iters * (double)c.grid * (c.block / 32);
            printf(", \"dmma_%s_tflops_w%d\": %.3f", names[s], cps * 8, flops / ms[s] * 1e-9);
        }
    }
    // few accumulators (latency-exposed) to learn the dependent-issue latency: 1 accumulator, 1 warp per SMSP
    {
        Cfg c{out, 4000, sms, 128};
        float ms1 = time_ms(l_dmma<0, 1>, &c, 3);
        float ms3 = time_ms(l_dmma<3, 1>, &c, 3);
        printf(", \"dmma_m8n8k4_dep_ns\": %.2f, \"dmma_m16n8k16_dep_ns\": %.2f", ms1 * 1e6 / c.iters, ms3 * 1e6 / c.iters);
    }

    // cuBLAS
    cublasHandle_t h; cublasCreate(&h);
    double *A, *B, *C;
    size_t bytes = (size_t)N * N * sizeof(double);
    CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
    CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
    double one = 1.0, zero = 0.0;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, N, N, N, &one, A, N, B, N, &zero, C, N);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, N, N, N, &one, A, N, B, N, &zero, C, N);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf(", \"cublas_dgemm_n\": %d, \"cublas_dgemm_tflops\": %.3f", N, 2.0 * N * N * (double)N / best * 1e-9);
    // sustained: back-to-back for ~3 s
    {
        int reps = (int)(3000.0f / best) + 1;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, N, N, N, &one, A, N, B, N, &zero, C, N);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(", \"cublas_dgemm_tflops_sustained\": %.3f", 2.0 * N * N * (double)N * reps / ms * 1e-9);
    }
    for (int w = 0; w < 2; ++w) cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, N, N, &one, A, N, &zero, C, N);
    CK(cudaDeviceSynchronize());
    best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, N, N, &one, A, N, &zero, C, N);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf(", \"cublas_dsyrk_tflops_syrkcount\": %.3f", (double)N * (N + 1) * (double)N / best * 1e-9);
    printf("}\n");
    return 0;
}
