// Exact residual sums of squares of every recorded candidate: second streaming pass over the samples.
//
// Replaces `rss(cache) = sum(abs2, X * A~ .- B~)` (reference lib/DataDrivenSparse/src/DataDrivenSparse.jl:
// 173-176), which the selector evaluates after every step (solver.jl:100).  The Gram identity
// y'y - 2 c'r + c'Gc cancels catastrophically for an exact fit, so the residual
// is formed sample by sample: candidate c of target e contributes (y_e(s) - sum_t c_t theta_t(x(s)))^2.
//
// Layout: each CTA owns a contiguous range of samples and walks it in tiles of 32 samples, double
// buffered with cp.async; a tile is staged TRANSPOSED ([state or target][sample], odd pitch) so that
// "lane = sample" reads are conflict-free while the global reads stay coalesced (each byte of X/Y is
// read once per pass).  Each WARP owns candidates (warp, warp + 8, ...); every lane keeps one partial
// sum per candidate in registers over the CTA's whole range, reduced at the end by a fixed shuffle
// tree; a second kernel adds the per-CTA partials in CTA order -> bit-reproducible, no atomics.  The
// candidates' sparse terms are expanded once into (coefficient, factor offsets) records, read through
// the read-only path (warp-uniform addresses), so the inner loop is loads + multiplies only.
// HBM-bound when the candidates are sparse: algorithmic bytes = 8 (n + n_t) per sample.
#include "common.cuh"
#include "sweep.cuh"

namespace ddb {

struct __align__(16) RssTerm {
    int16_t off[kMaxFactors];  // offset inside one sample's x row, -1 = no factor
    double val;
    double pad;
};

__global__ void rss_expand_kernel(const int32_t* __restrict__ cand_idx, const double* __restrict__ cand_val,
                                  int64_t nterms, const uint16_t* __restrict__ factors, RssTerm* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nterms) return;
    RssTerm r;
    r.val = cand_val[t];
    r.pad = 0.0;
    const uint16_t* f = factors + (size_t)cand_idx[t] * kMaxFactors;
#pragma unroll
    for (int d = 0; d < kMaxFactors; ++d) r.off[d] = (f[d] == kNoFactor) ? (int16_t)-1 : (int16_t)f[d];
    out[t] = r;
}

constexpr int kRssTile = 32;   // samples per shared-memory tile (one per lane)
constexpr int kRssAcc = 32;    // candidates per warp per pass (register accumulators per lane)
constexpr int kRssWarps = 8;
constexpr int kRssPitch = kRssTile + 1;  // doubles; transposed tile [state][sample], odd pitch -> conflict-free

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

__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}

// Stage one tile of kRssTile samples TRANSPOSED into shared memory ([row][sample], rows = n states then
// n_t targets) with 8-byte asynchronous copies: global reads stay coalesced (a sample's values are
// contiguous), the transposition makes "lane = sample" accesses conflict-free.
__device__ __forceinline__ void rss_stage_tile(const RssArgs& a, double* buf, int64_t s0, int valid) {
    const int rows = a.n + a.n_t;
    for (int e = threadIdx.x; e < kRssTile * a.n; e += blockDim.x) {
        const int s = e / a.n, i = e % a.n;
        if (s < valid) cp_async8(buf + i * kRssPitch + s, a.X + (s0 + s) * a.n + i);
    }
    for (int e = threadIdx.x; e < kRssTile * a.n_t; e += blockDim.x) {
        const int s = e / a.n_t, i = e % a.n_t;
        if (s < valid) cp_async8(buf + (a.n + i) * kRssPitch + s, a.Y + (s0 + s) * a.n_t + i);
    }
    (void)rows;
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kRssWarps * 32) rss_kernel(const RssArgs a) {
    extern __shared__ double sm[];
    const int rows = a.n + a.n_t;
    double* buf[2] = {sm, sm + (size_t)rows * kRssPitch};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cbase = blockIdx.y * (kRssWarps * kRssAcc);
    // this CTA's sample range (whole tiles)
    const int64_t ntiles = (a.m + kRssTile - 1) / kRssTile;
    const int64_t t_lo = ntiles * blockIdx.x / gridDim.x, t_hi = ntiles * (blockIdx.x + 1) / gridDim.x;

    // candidates of this warp: cbase + warp + 8 * slot
    int nslots = 0;
    for (int q = 0; q < kRssAcc; ++q)
        if (cbase + warp + kRssWarps * q < a.ncand) nslots = q + 1;
    double acc[kRssAcc];
#pragma unroll
    for (int q = 0; q < kRssAcc; ++q) acc[q] = 0.0;

    if (t_lo < t_hi) rss_stage_tile(a, buf[0], t_lo * kRssTile, (int)min((int64_t)kRssTile, a.m - t_lo * kRssTile));
    for (int64_t tile = t_lo; tile < t_hi; ++tile) {
        const int cur = (int)((tile - t_lo) & 1);
        if (tile + 1 < t_hi) {
            const int64_t s1 = (tile + 1) * kRssTile;
            rss_stage_tile(a, buf[cur ^ 1], s1, (int)min((int64_t)kRssTile, a.m - s1));
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int valid = (int)min((int64_t)kRssTile, a.m - tile * kRssTile);
        const double* xs = buf[cur] + lane;
        const bool live = lane < valid;
#pragma unroll
        for (int q = 0; q < kRssAcc; ++q) {
            if (q < nslots) {
                const int c = cbase + warp + kRssWarps * q;
                const int64_t off = __ldg(a.cand_off + c);
                const int nnz = __ldg(a.cand_nnz + c);
                const int eq = __ldg(a.cand_eq + c);
                double fit = 0.0;
                for (int t = 0; t < nnz; ++t) {
                    const RssTerm* r = a.terms + off + t;
                    double p = __ldg(&r->val);
                    const int4 o4 = __ldg(reinterpret_cast<const int4*>(r->off));  // 8 x int16
                    const int o[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
                    for (int d = 0; d < kMaxFactors; ++d) {
                        if (d < a.nf) {
                            const int od = (int)(short)((o[d >> 1] >> (16 * (d & 1))) & 0xFFFF);
                            if (od >= 0) p *= xs[od * kRssPitch];
                        }
                    }
                    fit += p;
                }
                const double res = live ? fit - xs[(a.n + eq) * kRssPitch] : 0.0;
                acc[q] = fma(res, res, acc[q]);
            }
        }
        __syncthreads();  // everyone is done with buf[cur] before it is refilled two iterations later
    }
    // lanes hold per-sample-slot partial sums: fixed shuffle tree, then one write per candidate
#pragma unroll
    for (int q = 0; q < kRssAcc; ++q) {
        if (q < nslots) {
            double v = acc[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) a.partials[(size_t)blockIdx.x * a.ncand + cbase + warp + kRssWarps * q] = v;
        }
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
    const int passes = (ncand + kRssWarps * kRssAcc - 1) / (kRssWarps * kRssAcc);
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
    const size_t smem = 2 * (size_t)kRssPitch * (ctx->n + ctx->n_t) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("ddb_rss: n + n_t too large for the shared-memory sample tile");
        return DDB_UNSUPPORTED;
    }
    DDB_CUDA_TRY(cudaFuncSetAttribute(rss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rss_kernel<<<dim3(nparts, passes), kRssWarps * 32, smem, st>>>(a);
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
