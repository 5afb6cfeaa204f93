// Per-target sums of the bound targets: what the StatsAPI statistics of a solution need besides rss.
//
// `nullloglikelihood(sol) = -n/2 log(sum(abs2, Y .- mean(Y, dims = 2)) / n)` (reference src/solution.jl:139-145) and
// through it `r2` (CoxSnell, :147-157) re-read all of Y on the CPU; here two HBM-bound passes over the resident
// targets give the per-target mean and centred sum of squares (two passes: sum y^2 - (sum y)^2 / m cancels).  On a
// sharded run both sums are all-reduced inside the library (collective call).
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace ddb {

int comm_allreduce(ddb_ctx* ctx, double* d, int64_t count);

// partials[blockIdx.x * n_t + e] = sum over this block's samples of (Y[e, s] - shift[e]) or its square.
// The tile of `ts` samples is staged through shared memory (coalesced loads); thread t sums target t % n_t over the
// samples t / n_t, t / n_t + groups, ...; the groups of a target are added in fixed order (deterministic).
template <bool SQ>
__global__ void __launch_bounds__(256) target_sums_kernel(const double* __restrict__ Y, int n_t, int64_t m,
                                                          const double* __restrict__ shift, int ts, double* __restrict__ partials) {
    extern __shared__ double tsm[];
    double* tile = tsm;                     // ts x n_t
    double* red = tsm + (size_t)ts * n_t;   // 256
    const int tid = threadIdx.x;
    const int64_t ntiles = (m + ts - 1) / ts;
    const int64_t t_lo = ntiles * blockIdx.x / gridDim.x, t_hi = ntiles * (blockIdx.x + 1) / gridDim.x;
    for (int e0 = 0; e0 < n_t; e0 += 256) {  // target chunks of 256 (one pass over the block's samples per chunk)
        const int ne = min(256, n_t - e0);
        const int groups = max(1, 256 / ne);
        const int e = tid % ne, g = tid / ne;
        const bool active = g < groups;
        const double sh = (shift && active) ? shift[e0 + e] : 0.0;
        double acc = 0.0;
        for (int64_t tl = t_lo; tl < t_hi; ++tl) {
            const int64_t s0 = tl * ts;
            const int valid = (int)min((int64_t)ts, m - s0);
            __syncthreads();
            for (int q = tid; q < valid * n_t; q += 256) tile[q] = Y[s0 * n_t + q];
            __syncthreads();
            if (active)
                for (int s = g; s < valid; s += groups) {
                    const double v = tile[(size_t)s * n_t + e0 + e] - sh;
                    acc += SQ ? v * v : v;
                }
        }
        __syncthreads();
        red[tid] = active ? acc : 0.0;
        __syncthreads();
        if (tid < ne) {
            double s = 0.0;
            for (int gg = 0; gg < groups; ++gg) s += red[tid + ne * gg];
            partials[(size_t)blockIdx.x * n_t + e0 + tid] = s;
        }
    }
}

__global__ void target_reduce_kernel(const double* __restrict__ partials, int nparts, int n_t, double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_t) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partials[(size_t)p * n_t + e];  // fixed order
    out[e] = s;
}

int target_stats_run(ddb_ctx* ctx, int64_t m_total, double* mean, double* css) {
    if (ctx->view_empty && ctx->n_t > 0 && mean && css) {  // this rank contributes zeros to both sums
        const int n_t = ctx->n_t;
        cudaStream_t st = ctx->stream;
        DDB_TRY(ctx->d_stats.reserve(2 * (size_t)n_t * 8));
        double* d_sum = ctx->d_stats.as<double>();
        std::vector<double> h(n_t);
        for (int pass = 0; pass < 2; ++pass) {
            DDB_CUDA_TRY(cudaMemsetAsync(d_sum, 0, (size_t)n_t * 8, st));
            DDB_TRY(comm_allreduce(ctx, d_sum, n_t));
            DDB_CUDA_TRY(cudaMemcpyAsync(h.data(), d_sum, (size_t)n_t * 8, cudaMemcpyDeviceToHost, st));
            DDB_CUDA_TRY(cudaStreamSynchronize(st));
            for (int e = 0; e < n_t; ++e) (pass == 0 ? mean : css)[e] = pass == 0 ? h[e] / (double)m_total : h[e];
        }
        return DDB_OK;
    }
    if (!ctx->dY || ctx->m <= 0 || ctx->n_t <= 0) {
        set_error("ddb_target_stats: no data bound");
        return DDB_BAD_STATE;
    }
    if (!mean || !css || m_total < ctx->m) {
        set_error("ddb_target_stats: mean, css must be non-null and m_total >= the local sample count");
        return DDB_INVALID_ARG;
    }
    const int n_t = ctx->n_t;
    const int64_t m = ctx->m;
    int ts = (int)std::max<int64_t>(32, std::min<int64_t>(4096 / n_t, 1024));
    while ((size_t)ts * n_t * 8 > 96 * 1024 && ts > 1) ts /= 2;
    const size_t smem = ((size_t)ts * n_t + 256) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("ddb_target_stats: %d targets do not fit the shared-memory tile", n_t);
        return DDB_UNSUPPORTED;
    }
    const int64_t ntiles = (m + ts - 1) / ts;
    const int nblk = (int)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * 8);
    cudaStream_t st = ctx->stream;
    DDB_TRY(ctx->d_stats.reserve(((size_t)nblk * n_t + 2 * (size_t)n_t) * 8));
    double* partials = ctx->d_stats.as<double>();
    double* d_sum = partials + (size_t)nblk * n_t;
    double* d_css = d_sum + n_t;
    DDB_CUDA_TRY(cudaFuncSetAttribute(target_sums_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DDB_CUDA_TRY(cudaFuncSetAttribute(target_sums_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    target_sums_kernel<false><<<nblk, 256, smem, st>>>(ctx->dY, n_t, m, nullptr, ts, partials);
    target_reduce_kernel<<<(n_t + 255) / 256, 256, 0, st>>>(partials, nblk, n_t, d_sum);
    DDB_TRY(comm_allreduce(ctx, d_sum, n_t));
    std::vector<double> h(n_t);
    DDB_CUDA_TRY(cudaMemcpyAsync(h.data(), d_sum, (size_t)n_t * 8, cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    for (int e = 0; e < n_t; ++e) mean[e] = h[e] / (double)m_total;
    DDB_CUDA_TRY(cudaMemcpyAsync(d_sum, mean, (size_t)n_t * 8, cudaMemcpyHostToDevice, st));
    target_sums_kernel<true><<<nblk, 256, smem, st>>>(ctx->dY, n_t, m, d_sum, ts, partials);
    target_reduce_kernel<<<(n_t + 255) / 256, 256, 0, st>>>(partials, nblk, n_t, d_css);
    DDB_TRY(comm_allreduce(ctx, d_css, n_t));
    DDB_CUDA_TRY(cudaMemcpyAsync(css, d_css, (size_t)n_t * 8, cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    return DDB_OK;
}

}  // namespace ddb

extern "C" int ddb_target_stats(ddb_ctx* ctx, int64_t m_total, double* mean, double* css) {
    if (!ctx) {
        ddb::set_error("ddb_target_stats: null context");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    return ddb::target_stats_run(ctx, m_total, mean, css);
}
