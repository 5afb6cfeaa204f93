// Exact residual sums of squares of every recorded candidate: second streaming pass over the samples.
//
// Replaces `rss(cache) = sum(abs2, X * A~ .- B~)` (reference lib/DataDrivenSparse/src/DataDrivenSparse.jl:
// 173-176), which the selector evaluates after every step (solver.jl:100).  The Gram identity
// y'y - 2 c'r + c'Gc cancels catastrophically for an exact fit, so the residual
// is formed sample by sample: candidate c of target e contributes (y_e(s) - sum_t c_t theta_t(x(s)))^2.
//
// Layout: each CTA owns a contiguous range of samples and walks it in tiles staged in shared memory
// (X tile and Y tile, sample-major exactly as in HBM -> coalesced, each byte of X/Y read once per pass);
// each THREAD owns candidates (tid, tid + 256, ...) and keeps their sums in registers over the CTA's
// whole range, so there is no cross-thread reduction and the result is bit-reproducible.  A second
// kernel adds the per-CTA partials in CTA order.  The candidates' sparse terms are expanded once into
// (coefficient, factor offsets) records so the inner loop is loads + multiplies only.
// HBM-bound when the candidates are sparse: algorithmic bytes = 8 (n + n_t) per sample.
#include "common.cuh"
#include "sweep.cuh"

namespace ddb {

struct RssTerm {
    double val;
    int16_t off[kMaxFactors];  // offset inside one sample's x row, -1 = no factor
};

__global__ void rss_expand_kernel(const int32_t* __restrict__ cand_idx, const double* __restrict__ cand_val,
                                  int64_t nterms, const uint16_t* __restrict__ factors, RssTerm* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nterms) return;
    RssTerm r;
    r.val = cand_val[t];
    const uint16_t* f = factors + (size_t)cand_idx[t] * kMaxFactors;
#pragma unroll
    for (int d = 0; d < kMaxFactors; ++d) r.off[d] = (f[d] == kNoFactor) ? (int16_t)-1 : (int16_t)f[d];
    out[t] = r;
}

constexpr int kRssTile = 32;  // samples per shared-memory tile
constexpr int kRssAcc = 4;    // candidates per thread per pass

struct RssArgs {
    const double* X;
    const double* Y;
    int n, n_t, nf;
    int64_t m;
    int ncand;
    const int64_t* cand_off;
    const int32_t* cand_nnz;
    const int32_t* cand_eq;
    const RssTerm* terms;
    double* partials;  // gridDim.x x ncand
};

__global__ void __launch_bounds__(256) rss_kernel(const RssArgs a) {
    extern __shared__ double sm[];
    double* xs = sm;                       // kRssTile x n
    double* ys = sm + (size_t)kRssTile * a.n;  // kRssTile x n_t
    const int tid = threadIdx.x;
    const int cbase = blockIdx.y * (256 * kRssAcc);
    // this CTA's sample range (whole tiles)
    const int64_t ntiles = (a.m + kRssTile - 1) / kRssTile;
    const int64_t t_lo = ntiles * blockIdx.x / gridDim.x, t_hi = ntiles * (blockIdx.x + 1) / gridDim.x;

    int64_t off[kRssAcc];
    int nnz[kRssAcc], eq[kRssAcc];
    double acc[kRssAcc];
#pragma unroll
    for (int q = 0; q < kRssAcc; ++q) {
        const int c = cbase + q * 256 + tid;
        acc[q] = 0.0;
        nnz[q] = -1;
        off[q] = 0;
        eq[q] = 0;
        if (c < a.ncand) {
            off[q] = a.cand_off[c];
            nnz[q] = a.cand_nnz[c];
            eq[q] = a.cand_eq[c];
        }
    }
    for (int64_t tile = t_lo; tile < t_hi; ++tile) {
        const int64_t s0 = tile * kRssTile;
        const int valid = (int)min((int64_t)kRssTile, a.m - s0);
        __syncthreads();
        for (int e = tid; e < valid * a.n; e += 256) xs[e] = a.X[s0 * a.n + e];
        for (int e = tid; e < valid * a.n_t; e += 256) ys[e] = a.Y[s0 * a.n_t + e];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kRssAcc; ++q) {
            if (nnz[q] < 0) continue;
            const RssTerm* tp = a.terms + off[q];
            for (int s = 0; s < valid; ++s) {
                const double* xr = xs + s * a.n;
                double fit = 0.0;
                for (int t = 0; t < nnz[q]; ++t) {
                    const RssTerm r = tp[t];
                    double p = r.val;
                    for (int d = 0; d < a.nf; ++d)
                        if (r.off[d] >= 0) p *= xr[r.off[d]];
                    fit += p;
                }
                const double res = fit - ys[s * a.n_t + eq[q]];
                acc[q] = fma(res, res, acc[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kRssAcc; ++q) {
        const int c = cbase + q * 256 + tid;
        if (c < a.ncand) a.partials[(size_t)blockIdx.x * a.ncand + c] = acc[q];
    }
}

__global__ void rss_reduce_kernel(const double* __restrict__ partials, int nparts, int ncand, double* __restrict__ rss) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partials[(size_t)p * ncand + c];  // fixed order
    rss[c] = s;
}

int rss_run(ddb_ctx* ctx, SweepState* S, double* dRss) {
    if (!ctx->dX || !ctx->dY || ctx->m <= 0) {
        set_error("ddb_rss: no data bound");
        return DDB_BAD_STATE;
    }
    if (ctx->factors_nt != ctx->n_t || !ctx->d_factors.p) {
        set_error("ddb_rss: the library factor table is missing (run ddb_gram on this context first)");
        return DDB_BAD_STATE;
    }
    const int ncand = S->ncand;
    cudaStream_t st = ctx->stream;
    if (!dRss) {
        DDB_TRY(S->rss.reserve((size_t)std::max(1, ncand) * 8));
        dRss = S->rss.as<double>();
    }
    const int64_t nterms = S->pool_used;
    DDB_TRY(S->cand_terms.reserve((size_t)std::max<int64_t>(1, nterms) * sizeof(RssTerm)));
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[4], st));
    if (nterms > 0)
        rss_expand_kernel<<<(unsigned)((nterms + 255) / 256), 256, 0, st>>>(S->cand_idx.as<int32_t>(), S->cand_val.as<double>(),
                                                                            nterms, ctx->d_factors.as<uint16_t>(),
                                                                            S->cand_terms.as<RssTerm>());
    const int passes = (ncand + 256 * kRssAcc - 1) / (256 * kRssAcc);
    const int64_t ntiles = (ctx->m + kRssTile - 1) / kRssTile;
    const int nparts = (int)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * 4);
    DDB_TRY(S->rss_partials.reserve((size_t)nparts * ncand * 8));
    RssArgs a;
    a.X = ctx->dX;
    a.Y = ctx->dY;
    a.n = ctx->n;
    a.n_t = ctx->n_t;
    a.nf = std::max(1, ctx->max_degree);
    a.m = ctx->m;
    a.ncand = ncand;
    a.cand_off = S->cand_off.as<int64_t>();
    a.cand_nnz = S->cand_nnz.as<int32_t>();
    a.cand_eq = S->cand_eq.as<int32_t>();
    a.terms = S->cand_terms.as<RssTerm>();
    a.partials = S->rss_partials.as<double>();
    const size_t smem = (size_t)kRssTile * (ctx->n + ctx->n_t) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("ddb_rss: n + n_t too large for the shared-memory sample tile");
        return DDB_UNSUPPORTED;
    }
    DDB_CUDA_TRY(cudaFuncSetAttribute(rss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rss_kernel<<<dim3(nparts, passes), 256, smem, st>>>(a);
    rss_reduce_kernel<<<(ncand + 255) / 256, 256, 0, st>>>(S->rss_partials.as<double>(), nparts, ncand, dRss);
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[5], st));
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    ctx->timings.rss_ms = ms;
    ctx->timings.sweep_launches += 3;
    S->rss_ptr = dRss;
    S->have_rss = true;
    return DDB_OK;
}

}  // namespace ddb
