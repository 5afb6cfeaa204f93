// After the Gram blocks of the DataDrivenDMD path: the exact residual of the fitted operator and the dense
// eigen-decompositions of the algorithms, all behind the C ABI.
//
// * ddb_dmd_residual: rss = || Z - M X ||_F^2 sample by sample (reference lib/DataDrivenDMD/src/result.jl:46-59 ->
//   `sum(abs2, Z .- C * (K * X + B * U))`), a DMMA product whose prediction tile never leaves the registers
//   (trsm.cu: gemm_nt_resid_kernel).  The Gram identity tr(S) - 2 tr(M Q') + tr(M P M') would cancel for a good fit.
// * ddb_dmd_eig_sym / ddb_dmd_eig: `eigen(Symmetric(P))` for the truncated SVD through P = U S^2 U'
//   (algorithms.jl:2-20) and `eigen(K)` / `eigen(Atilde)` (algorithms.jl:81, 157).  Dense n x n LAPACK-type work with
//   no structure to exploit: cuSOLVER (Xsyevd / Xgeev), bound at run time, as SURVEY.md H8 plans.
#include "common.cuh"
#include <cusolverDn.h>
#include <dlfcn.h>
#include <mutex>
#include <vector>

namespace ddb {

int gemm_nt_residual(ddb_ctx* ctx, const double* Z, size_t ldz, const double* A, size_t lda, const double* B, size_t ldb, int64_t rows,
                     int cols, int K, double* d_partials, double* d_out);

// out (rows x cols row-major) = in (rows x cols column-major)
__global__ void colmajor_to_rowmajor_kernel(const double* __restrict__ in, int rows, int cols, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int q = ty; q < 32; q += 8) {
        const int r = r0 + tx, c = c0 + q;
        tile[q][tx] = (r < rows && c < cols) ? in[(size_t)r + (size_t)rows * c] : 0.0;
    }
    __syncthreads();
    for (int q = ty; q < 32; q += 8) {
        const int r = r0 + q, c = c0 + tx;
        if (r < rows && c < cols) out[(size_t)r * cols + c] = tile[tx][q];
    }
}

struct SolverApi {
    void* handle = nullptr;
    cusolverStatus_t (*Create)(cusolverDnHandle_t*) = nullptr;
    cusolverStatus_t (*Destroy)(cusolverDnHandle_t) = nullptr;
    cusolverStatus_t (*SetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
    cusolverStatus_t (*CreateParams)(cusolverDnParams_t*) = nullptr;
    cusolverStatus_t (*DestroyParams)(cusolverDnParams_t) = nullptr;
    decltype(&cusolverDnXsyevd_bufferSize) SyevdBuf = nullptr;
    decltype(&cusolverDnXsyevd) Syevd = nullptr;
    decltype(&cusolverDnXgeev_bufferSize) GeevBuf = nullptr;
    decltype(&cusolverDnXgeev) Geev = nullptr;
};
static SolverApi g_sol;
static std::mutex g_sol_mutex;

static int solver_load() {
    std::lock_guard<std::mutex> lock(g_sol_mutex);
    if (g_sol.handle) return DDB_OK;
    const char* names[] = {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11",
                           "/usr/local/cuda/lib64/libcusolver.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("cuSOLVER is not available (dlopen libcusolver.so.11: %s); ddb_dmd_eig* need it", dlerror());
        return DDB_UNSUPPORTED;
    }
    SolverApi a;
    a.handle = h;
#define DDB_SOL_SYM(field, name)                                             \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));           \
    if (!a.field) {                                                          \
        set_error("cuSOLVER symbol %s is missing from the loaded library", name); \
        return DDB_UNSUPPORTED;                                              \
    }
    DDB_SOL_SYM(Create, "cusolverDnCreate")
    DDB_SOL_SYM(Destroy, "cusolverDnDestroy")
    DDB_SOL_SYM(SetStream, "cusolverDnSetStream")
    DDB_SOL_SYM(CreateParams, "cusolverDnCreateParams")
    DDB_SOL_SYM(DestroyParams, "cusolverDnDestroyParams")
    DDB_SOL_SYM(SyevdBuf, "cusolverDnXsyevd_bufferSize")
    DDB_SOL_SYM(Syevd, "cusolverDnXsyevd")
    DDB_SOL_SYM(GeevBuf, "cusolverDnXgeev_bufferSize")
    DDB_SOL_SYM(Geev, "cusolverDnXgeev")
#undef DDB_SOL_SYM
    g_sol = a;
    return DDB_OK;
}

#define DDB_SOL_TRY(expr)                                                                  \
    do {                                                                                   \
        cusolverStatus_t s__ = (expr);                                                     \
        if (s__ != CUSOLVER_STATUS_SUCCESS) {                                              \
            ::ddb::set_error("%s failed: cusolver status %d (%s:%d)", #expr, (int)s__, __FILE__, __LINE__); \
            status = DDB_CUDA_ERROR;                                                       \
            goto done;                                                                     \
        }                                                                                  \
    } while (0)

// jobs: 0 = symmetric (A lower, column-major; eigenvectors overwrite A, eigenvalues ascending to W),
//       1 = general (A overwritten; W = n complex eigenvalues interleaved; VR = right eigenvectors in LAPACK's packed
//           real form: a complex pair (j, j + 1) is stored as columns re, im)
static int eig_run(ddb_ctx* ctx, int job, double* dA, int n, double* dW, double* dVR) {
    DDB_TRY(solver_load());
    int status = DDB_OK;
    cusolverDnHandle_t h = nullptr;
    cusolverDnParams_t prm = nullptr;
    size_t wd = 0, wh = 0;
    void* dwork = nullptr;
    int* dinfo = nullptr;
    std::vector<char> hwork;
    int info = 0;
    DDB_SOL_TRY(g_sol.Create(&h));
    DDB_SOL_TRY(g_sol.SetStream(h, ctx->stream));
    DDB_SOL_TRY(g_sol.CreateParams(&prm));
    if (job == 0)
        DDB_SOL_TRY(g_sol.SyevdBuf(h, prm, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW,
                                   CUDA_R_64F, &wd, &wh));
    else
        DDB_SOL_TRY(g_sol.GeevBuf(h, prm, CUSOLVER_EIG_MODE_NOVECTOR, dVR ? CUSOLVER_EIG_MODE_VECTOR : CUSOLVER_EIG_MODE_NOVECTOR, n,
                                  CUDA_R_64F, dA, n, CUDA_C_64F, dW, CUDA_R_64F, nullptr, 1, CUDA_R_64F, dVR, n, CUDA_R_64F, &wd, &wh));
    if (cudaMalloc(&dwork, wd + 16) != cudaSuccess || cudaMalloc(&dinfo, sizeof(int)) != cudaSuccess) {
        set_error("ddb_dmd_eig: out of device memory for the cuSOLVER workspace (%zu bytes)", wd);
        status = DDB_OUT_OF_MEMORY;
        goto done;
    }
    hwork.resize(wh + 16);
    if (job == 0)
        DDB_SOL_TRY(g_sol.Syevd(h, prm, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW,
                                CUDA_R_64F, dwork, wd, hwork.data(), wh, dinfo));
    else
        DDB_SOL_TRY(g_sol.Geev(h, prm, CUSOLVER_EIG_MODE_NOVECTOR, dVR ? CUSOLVER_EIG_MODE_VECTOR : CUSOLVER_EIG_MODE_NOVECTOR, n,
                               CUDA_R_64F, dA, n, CUDA_C_64F, dW, CUDA_R_64F, nullptr, 1, CUDA_R_64F, dVR, n, CUDA_R_64F, dwork, wd,
                               hwork.data(), wh, dinfo));
    if (cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        set_error("ddb_dmd_eig: %s", cudaGetErrorString(cudaGetLastError()));
        status = DDB_CUDA_ERROR;
        goto done;
    }
    if (info != 0) {
        set_error("ddb_dmd_eig: the eigen-solver did not converge (info = %d)", info);
        status = DDB_NOT_POSITIVE_DEFINITE;
    }
done:
    if (dwork) cudaFree(dwork);
    if (dinfo) cudaFree(dinfo);
    if (prm) g_sol.DestroyParams(prm);
    if (h) g_sol.Destroy(h);
    return status;
}

// sweep.cu (minimum-norm steps of rank-deficient masked systems beyond the shared-memory Jacobi solver): eigenvectors
// of the symmetric n x n matrix dA (lower triangle read, column-major == row-major) overwrite it as columns,
// eigenvalues ascending to dW
int eig_sym_device(ddb_ctx* ctx, double* dA, int n, double* dW) { return eig_run(ctx, 0, dA, n, dW, nullptr); }

}  // namespace ddb

using namespace ddb;

extern "C" {

int ddb_dmd_residual(ddb_ctx* ctx, const double* dX, const double* dZ, const double* dM, int n_in, int n_out, int64_t m, double* rss) {
    if (!ctx || !dX || !dZ || !dM || !rss || n_in <= 0 || n_out <= 0 || m <= 0) {
        set_error("ddb_dmd_residual: null pointer or non-positive size");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t ntile = ((m + 127) / 128) * ((n_out + 127) / 128);
    DDB_TRY(ctx->d_stats.reserve(((size_t)n_in * n_out + (size_t)ntile + 2) * 8));
    double* Mrow = ctx->d_stats.as<double>();
    double* partials = Mrow + (size_t)n_in * n_out;
    double* d_out = partials + ntile;
    cudaStream_t st = ctx->stream;
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[4], st));
    colmajor_to_rowmajor_kernel<<<dim3((n_in + 31) / 32, (n_out + 31) / 32), 256, 0, st>>>(dM, n_out, n_in, Mrow);
    // prediction[s][o] = sum_i X[s][i] M[o][i]: samples are the rows (sample-contiguous storage = row-major m x n)
    DDB_TRY(gemm_nt_residual(ctx, dZ, (size_t)n_out, dX, (size_t)n_in, Mrow, (size_t)n_in, m, n_out, n_in, partials, d_out));
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[5], st));
    DDB_CUDA_TRY(cudaMemcpyAsync(rss, d_out, 8, cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    ctx->timings.rss_ms = ms;
    return DDB_OK;
}

int ddb_dmd_eig_sym(ddb_ctx* ctx, double* dA, int n, double* dW) {
    if (!ctx || !dA || !dW || n <= 0) {
        set_error("ddb_dmd_eig_sym: null pointer or non-positive size");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    return eig_run(ctx, 0, dA, n, dW, nullptr);
}

int ddb_dmd_eig(ddb_ctx* ctx, double* dA, int n, double* dW, double* dVR) {
    if (!ctx || !dA || !dW || n <= 0) {
        set_error("ddb_dmd_eig: null pointer or non-positive size");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    return eig_run(ctx, 1, dA, n, dW, dVR);
}

}  // extern "C"
