// Stand-alone timing / accuracy harness of the blocked Cholesky kernels (csrc/chol.cu is included as is).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/chol_bench.cu -o tools/bin/chol_bench
//   tools/bin/chol_bench [k = 2145] [nrhs = 64] [batch = 1] [reps = 20]
// Builds a random SPD matrix (Gram of a random k x 2k matrix, equilibrated), factorises it with the extra rows, runs
// the backward kernel, reports CUDA-event times per kernel and the residual |A x - b| / (|A| |x|).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../datadrivendiffeq.jl_b200/csrc/chol.cu"

namespace ddb {
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}
int DevBuf::reserve(size_t) { return 0; }
void DevBuf::release() {}
}  // namespace ddb

using namespace ddb;

#define CK(x)                                                                     \
    do {                                                                          \
        cudaError_t e = (x);                                                      \
        if (e != cudaSuccess) {                                                   \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));               \
            exit(1);                                                              \
        }                                                                         \
    } while (0)

int main(int argc, char** argv) {
    const int k = argc > 1 ? atoi(argv[1]) : 2145, nrhs = argc > 2 ? atoi(argv[2]) : 64, batch = argc > 3 ? atoi(argv[3]) : 1;
    const int reps = argc > 4 ? atoi(argv[4]) : 20;
    const size_t kk = (size_t)k * k, stride = kk + (size_t)nrhs * k;
    std::mt19937_64 rng(1);
    std::normal_distribution<double> nd;
    // a random symmetric, strictly diagonally dominant (hence positive definite) matrix with unit-order diagonal
    std::vector<double> A(kk), rhs((size_t)nrhs * k);
    std::uniform_real_distribution<double> ud(-1.0, 1.0);
    for (int i = 0; i < k; ++i) {
        for (int j = 0; j < i; ++j) A[(size_t)i * k + j] = A[(size_t)j * k + i] = ud(rng) / k;
        A[(size_t)i * k + i] = 1.5 + 0.5 * ud(rng);
    }
    for (auto& v : rhs) v = nd(rng);
    ddb_ctx ctx;
    CK(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
    ctx.sm_count = 148;
    double *dA0, *dA, *dX;
    int* dfail;
    CK(cudaMalloc(&dA0, stride * 8));
    CK(cudaMalloc(&dA, (size_t)batch * stride * 8));
    CK(cudaMalloc(&dX, (size_t)batch * nrhs * k * 8));
    CK(cudaMalloc(&dfail, batch * 4));
    CK(cudaMemset(dfail, 0, batch * 4));
    CK(cudaMemcpy(dA0, A.data(), kk * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA0 + kk, rhs.data(), (size_t)nrhs * k * 8, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventCreate(&e2));
    double tf = 0, tb = 0, tcls[3] = {0, 0, 0};
    for (int r = 0; r < reps + 2; ++r) {
        for (int b = 0; b < batch; ++b) CK(cudaMemcpyAsync(dA + b * stride, dA0, stride * 8, cudaMemcpyDeviceToDevice, ctx.stream));
        CK(cudaEventRecord(e0, ctx.stream));
        if (chol128_batched(&ctx, dA, stride, k, nullptr, k, batch, dfail, nrhs, k, nullptr)) return 1;
        CK(cudaEventRecord(e1, ctx.stream));
        if (chol_back_batched(&ctx, dA, stride, k, nullptr, k, batch, k, nrhs, dX, (size_t)nrhs * k, (size_t)k, nullptr, nullptr)) return 1;
        CK(cudaEventRecord(e2, ctx.stream));
        CK(cudaStreamSynchronize(ctx.stream));
        float a, b2;
        CK(cudaEventElapsedTime(&a, e0, e1));
        CK(cudaEventElapsedTime(&b2, e1, e2));
        if (r >= 2) tf += a, tb += b2;
    }
    {   // the same factorisation with an event pair around every launch: time per kernel class
        long long* dtrace;
        CK(cudaMalloc(&dtrace, 128));
        CK(cudaMemset(dtrace, 0, 128));
        CholBatch cb{dA, stride, k, nullptr, k, dfail, nrhs, k, dtrace};
        const size_t dsmem = (size_t)kCB * kCP * 8, psmem = (size_t)(kCB + 64) * kCP * 8, gsmem = 2 * (size_t)kCStage * 8;
        std::vector<cudaEvent_t> ev(4 * ((k + 127) / 128) + 4);
        for (auto& evt : ev) CK(cudaEventCreate(&evt));
        for (int r = 0; r < 5; ++r) {
            for (int b = 0; b < batch; ++b) CK(cudaMemcpyAsync(dA + b * stride, dA0, stride * 8, cudaMemcpyDeviceToDevice, ctx.stream));
            int ne = 0;
            std::vector<int> cls;
            CK(cudaEventRecord(ev[ne++], ctx.stream));
            for (int j0 = 0; j0 < k; j0 += kCB) {
                chol_diag128_kernel<<<batch, 256, dsmem, ctx.stream>>>(cb, j0);
                CK(cudaEventRecord(ev[ne++], ctx.stream));
                cls.push_back(0);
                const int rem = std::max(k - j0 - kCB, 0);
                if (rem + nrhs <= 0) break;
                const int Tr = (rem + nrhs + kCB - 1) / kCB, Tc = (rem + kCB - 1) / kCB;
                if ((long long)Tr * batch < 148) {
                    const size_t p32smem = ((size_t)kCB * kPT + 2 * 32 * kCP) * 8;
                    chol_panel32_kernel<<<dim3((rem + nrhs + kPR - 1) / kPR, batch), 128, p32smem, ctx.stream>>>(cb, j0);
                } else
                    chol_panel128_kernel<<<dim3(Tr, batch), 256, psmem, ctx.stream>>>(cb, j0);
                CK(cudaEventRecord(ev[ne++], ctx.stream));
                cls.push_back(1);
                if (Tc > 0) {
                    chol_syrk128_kernel<<<dim3(Tc, Tr, batch), 256, gsmem, ctx.stream>>>(cb, j0);
                    CK(cudaEventRecord(ev[ne++], ctx.stream));
                    cls.push_back(2);
                }
            }
            CK(cudaStreamSynchronize(ctx.stream));
            if (r >= 2)
                for (int i = 0; i + 1 < ne; ++i) {
                    float ms;
                    CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
                    tcls[cls[i]] += ms / 3.0;
                }
        }
        long long tr[16];
        CK(cudaMemcpy(tr, dtrace, 128, cudaMemcpyDeviceToHost));
        if (tr[12] > 0)
            printf("panel kernel, CTA 0, cycles per call: load/wait %lld, product %lld, substitution %lld, store %lld (%lld calls)\n",
                   tr[8] / tr[12], tr[9] / tr[12], tr[10] / tr[12], tr[11] / tr[12], tr[12]);
        if (tr[5] > 0)
            printf("diag kernel, CTA 0, cycles per call: load %lld, pivot %lld, rows below %lld, in-block update %lld, store %lld (%lld calls)\n",
                   tr[0] / tr[5], tr[1] / tr[5], tr[2] / tr[5], tr[3] / tr[5], tr[4] / tr[5], tr[5]);
        printf("per class (event pair around every launch): diag %.3f ms, panel %.3f ms, syrk %.3f ms\n", tcls[0], tcls[1], tcls[2]);
        if (chol_back_batched(&ctx, dA, stride, k, nullptr, k, batch, k, nrhs, dX, (size_t)nrhs * k, (size_t)k, nullptr, nullptr)) return 1;
        CK(cudaStreamSynchronize(ctx.stream));
    }
    std::vector<double> X((size_t)nrhs * k);
    CK(cudaMemcpy(X.data(), dX + (size_t)(batch - 1) * nrhs * k, X.size() * 8, cudaMemcpyDeviceToHost));
    int fail = 0;
    CK(cudaMemcpy(&fail, dfail, 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int q = 0; q < std::min(nrhs, 4); ++q) {
        double num = 0, den = 0;
        for (int i = 0; i < k; ++i) {
            double s = 0, sa = 0;
            for (int j = 0; j < k; ++j) s += A[(size_t)i * k + j] * X[(size_t)q * k + j], sa += fabs(A[(size_t)i * k + j] * X[(size_t)q * k + j]);
            num = std::max(num, fabs(s - rhs[(size_t)q * k + i]));
            den = std::max(den, sa);
        }
        worst = std::max(worst, num / den);
    }
    printf("{\"k\": %d, \"nrhs\": %d, \"batch\": %d, \"factor_ms\": %.4f, \"back_ms\": %.4f, \"factor_gflops\": %.1f, \"residual\": %.3e, \"fail\": %d}\n",
           k, nrhs, batch, tf / reps, tb / reps, batch * (double)k * k * k / 3.0 / (tf / reps * 1e-3) * 1e-9, worst, fail);
    return 0;
}
