// Exact streaming rss of DENSE candidates on the FP64 tensor pipe.
//
// The reference scores every model the sweep records by evaluating it on the data (`rss` of the cache,
// lib/DataDrivenSparse/src/solver.jl:72-118 through `X' * Theta - Y`).  rss.cu streams sparse candidates term by term
// (a few instructions per term and sample); with noisy data the candidates of the small thresholds keep hundreds or
// thousands of terms and that pass becomes the whole step (C3 with derivative noise 1.0: 2600 candidates of ~1000
// terms, 6.1 s per 1e6 samples).  For candidates with many terms the fit is a matrix product,
//     F[c, s] = sum_j C[c, j] theta_j(s),        rss_c = sum_s (F[c, s] + shift_c - y_e(c)(s))^2,
// contracted over the TERMS on DMMA.8x8x4: a CTA owns 64 candidates and a range of samples; for every tile of 64
// samples it walks the library in blocks of 64 terms -- the coefficient block C[64 x 64] arrives by cp.async, the
// Theta block [64 samples x 64 terms] is multiplied out from the TMA-fetched raw samples exactly as the Gram kernel
// does (same products, same rounding), both double-buffered under the 1024 DMMAs of the previous block -- and squares
// the residuals in the epilogue.  Theta is never written to memory and nothing here goes through the Gram matrix
// (the quadratic form y'y - 2 c'r + c'Gc loses the digits the selectors compare when coefficients are large and
// cancel, which is exactly the dense, ill-conditioned case).  Per-candidate sums are reduced in a fixed order.
#include "common.cuh"
#include "mma_common.cuh"
#include <algorithm>

namespace ddb {

constexpr int kDC = 64;   // candidates per CTA
constexpr int kDS = 64;   // samples per tile
constexpr int kDT = 64;   // terms per block
constexpr int kDP = 68;   // shared-memory pitch of both operand tiles: 68 = 4 (mod 16), conflict-free fragments

struct DenseArgs {
    const double* X;
    const double* Y;
    int n, n_t;
    int64_t m;
    int k, kpad;              // library terms, padded to a multiple of kDT
    const uint16_t* factors;  // (k + n_t) x kMaxFactors
    const double* Cd;         // nd_pad x kpad coefficients, zero beyond a candidate's support and beyond k
    const int32_t* eq;        // nd_pad: target of the candidate (0 for padding rows)
    const double* shift;      // nd_pad constants added to the fit, or null
    int nd_pad;
    double* partials;         // gridDim.x x nd_pad
};

__global__ void dense_fill_kernel(const int32_t* __restrict__ ids, int nd, const int64_t* __restrict__ off,
                                  const int32_t* __restrict__ nnz, const int32_t* __restrict__ eq,
                                  const int32_t* __restrict__ idx, const double* __restrict__ val,
                                  const double* __restrict__ scale, const int32_t* __restrict__ map, int kpad,
                                  double* __restrict__ Cd, int32_t* __restrict__ eq_out) {
    const int r = blockIdx.x;  // row of Cd (rows beyond nd stay zero)
    double* row = Cd + (size_t)r * kpad;
    for (int j = threadIdx.x; j < kpad; j += blockDim.x) row[j] = 0.0;
    __syncthreads();
    if (r >= nd) {
        if (threadIdx.x == 0) eq_out[r] = 0;
        return;
    }
    const int c = ids[r];
    const int64_t o = off[c];
    for (int t = threadIdx.x; t < nnz[c]; t += blockDim.x) {
        const int j = idx[o + t];
        const double v = scale ? val[o + t] / scale[j] : val[o + t];
        row[map ? map[j] : j] = v;
    }
    if (threadIdx.x == 0) eq_out[r] = eq[c];
}

template <int NF>
__global__ void __launch_bounds__(256, 1) rss_dense_kernel(const DenseArgs a) {
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(dsm_raw);  // [2] raw tile landed
    const int n = a.n;
    const int rawd = kDS * n;
    double* raw = reinterpret_cast<double*>(dsm_raw + 128);  // [2][kDS x n]
    double* Cs = raw + 2 * rawd;                             // [2][kDC x kDP]
    double* Ts = Cs + 2 * kDC * kDP;                         // [2][kDS x kDP]
    __shared__ double red[4][kDC];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wc = warp & 1, ws = warp >> 1;  // 32 candidates x 16 samples per warp
    const int cb = blockIdx.y;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t ntiles = (a.m + kDS - 1) / kDS;
    const int64_t t_lo = ntiles * blockIdx.x / gridDim.x, t_hi = ntiles * (blockIdx.x + 1) / gridDim.x;
    const int64_t full_tiles = a.m / kDS;
    const bool bulk_ok = ((size_t)rawd * 8) % 16 == 0 && (((uintptr_t)a.X) & 15) == 0;
    auto request_raw = [&](int64_t tile, int slot) {  // whole CTA
        double* dst = raw + (size_t)slot * rawd;
        const int64_t s0 = tile * kDS;
        if (tile < full_tiles && bulk_ok) {
            if (tid == 0) {
                mbar_expect_tx(&full[slot], (uint32_t)(rawd * sizeof(double)));
                bulk_g2s(dst, a.X + s0 * n, (uint32_t)(rawd * sizeof(double)), &full[slot]);
            }
        } else {  // ragged or unaligned tile: plain loads, zero fill
            const int64_t left = a.m - s0;
            const int64_t valid = (left < kDS ? left : (int64_t)kDS) * n;
            for (int e = tid; e < rawd; e += 256) dst[e] = e < valid ? a.X[s0 * n + e] : 0.0;
            __syncthreads();
            if (tid == 0) mbar_arrive(&full[slot]);
        }
    };
    // builder: term (tid % 64) of the block, samples 16 (tid / 64) .. + 16
    const int bterm = tid & 63, bs0 = (tid >> 6) * 16;
    int foff[NF];  // state index of factor d, -1: missing (a constant 1)
    auto load_factors = [&](int blk) {
        const int term = blk * kDT + bterm;
        if (term < a.k) {
            const uint4 q = *reinterpret_cast<const uint4*>(a.factors + (size_t)term * kMaxFactors);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int d = 0; d < NF; ++d) {
                const uint32_t f = (w[d >> 1] >> (16 * (d & 1))) & 0xFFFFu;
                foff[d] = (f == kNoFactor) ? -1 : (int)f;
            }
        } else {  // padding term: coefficient column is zero; keep the product finite
#pragma unroll
            for (int d = 0; d < NF; ++d) foff[d] = -1;
        }
    };
    auto build = [&](const double* rw, double* T) {
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const int sm = bs0 + s;
            double p = foff[0] >= 0 ? rw[sm * n + foff[0]] : 1.0;
#pragma unroll
            for (int d = 1; d < NF; ++d) p *= foff[d] >= 0 ? rw[sm * n + foff[d]] : 1.0;
            T[sm * kDP + bterm] = p;
        }
    };
    auto fetch_c = [&](int blk, double* C) {  // 64 x 64 doubles = 2048 16-byte pieces
        const double* src = a.Cd + (size_t)cb * kDC * a.kpad + (size_t)blk * kDT;
        for (int e = tid; e < kDC * kDT / 2; e += 256) {
            const int r = e >> 5, j2 = (e & 31) * 2;
            const uint32_t d = smem_u32(C + r * kDP + j2);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + (size_t)r * a.kpad + j2) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int nblk = a.kpad / kDT;
    int eqc[4];
    double shc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = cb * kDC + wc * 32 + 8 * i + g;
        eqc[i] = a.eq[c];
        shc[i] = a.shift ? a.shift[c] : 0.0;
    }
    double rsum[4] = {0.0, 0.0, 0.0, 0.0};
    if (t_lo < t_hi) request_raw(t_lo, 0);
    for (int64_t tile = t_lo; tile < t_hi; ++tile) {
        const int it = (int)(tile - t_lo), slot = it & 1;
        if (tile + 1 < t_hi) request_raw(tile + 1, slot ^ 1);
        mbar_wait(&full[slot], (uint32_t)((it >> 1) & 1));
        const double* rw = raw + (size_t)slot * rawd;
        double acc[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        // block 0 of this tile
        load_factors(0);
        fetch_c(0, Cs);
        build(rw, Ts);
        if (nblk > 1) load_factors(1);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll 1
        for (int b = 0; b < nblk; ++b) {
            const double* Cc = Cs + (b & 1) * kDC * kDP;
            const double* Tc = Ts + (b & 1) * kDS * kDP;
            // The next block travels / is multiplied out underneath this one's DMMAs.  A warp issues in order, so its
            // build (loads feeding multiplies) stalls its own DMMAs: the two warps of a sub-partition (w and w + 4) take
            // the two jobs in OPPOSITE order, one builds while the other keeps the tensor pipe busy.
            auto prefetch = [&]() {
                if (b + 1 < nblk) {
                    fetch_c(b + 1, Cs + ((b + 1) & 1) * kDC * kDP);
                    build(rw, Ts + ((b + 1) & 1) * kDS * kDP);
                    if (b + 2 < nblk) load_factors(b + 2);
                }
            };
            if (warp < 4) prefetch();
            const double* pa = Cc + (wc * 32 + g) * kDP + t;
            const double* pb = Tc + (ws * 16 + g) * kDP + t;
#pragma unroll 4
            for (int kk = 0; kk < kDT; kk += 4) {
                double af[4], bf[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = pa[8 * i * kDP + kk];
#pragma unroll
                for (int j = 0; j < 2; ++j) bf[j] = pb[8 * j * kDP + kk];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            if (warp >= 4) prefetch();
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        }
        // residuals of this tile
        const int64_t sbase = tile * kDS + ws * 16 + 2 * t;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t s = sbase + 8 * j + h;
                    if (s < a.m) {
                        const double r = acc[i][j][h] + shc[i] - a.Y[s * a.n_t + eqc[i]];
                        rsum[i] = fma(r, r, rsum[i]);
                    }
                }
    }
    // per candidate: the four lanes of a row, then the four sample warps, in a fixed order
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double v = rsum[i];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (t == 0) red[ws][wc * 32 + 8 * i + g] = v;
    }
    __syncthreads();
    if (tid < kDC)
        a.partials[(size_t)blockIdx.x * a.nd_pad + cb * kDC + tid] = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
}

__global__ void dense_reduce_kernel(const double* __restrict__ partials, int nparts, int nd_pad, int nd, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nd) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partials[(size_t)p * nd_pad + c];  // fixed order
    out[c] = s;
}

// Is the dense pass applicable to the bound data / library?  (16-byte tiles, shared memory for two raw tiles)
bool rss_dense_ok(const ddb_ctx* ctx) {
    if (const char* e = getenv("DDB_RSS_DENSE"))
        if (atoi(e) == 0) return false;
    if (ctx->view_empty || !ctx->dX || !ctx->dY || ctx->m < kDS) return false;
    const size_t smem = 128 + (2 * (size_t)kDS * ctx->n + 2 * (size_t)kDC * kDP + 2 * (size_t)kDS * kDP) * sizeof(double);
    return smem <= 220 * 1024;
}

// rss of `nd` candidates (device list `ids` into the candidate arrays) -> out[0 .. nd) (device), streamed over the
// samples of the current view.  scale / map as in rss_core; shift: per-ROW constants (nd) or null.
int rss_dense_run(ddb_ctx* ctx, const int32_t* ids, int nd, const int64_t* off, const int32_t* nnz, const int32_t* eq,
                  const int32_t* idx, const double* val, const double* scale, const int32_t* map, const double* shift,
                  DevBuf& work, double* out, int* launches) {
    if (nd <= 0) return DDB_OK;
    cudaStream_t st = ctx->stream;
    const int k = ctx->k;
    const int kpad = (k + kDT - 1) / kDT * kDT;
    const int nd_pad = (nd + kDC - 1) / kDC * kDC;
    const int cbn = nd_pad / kDC;
    const int64_t ntiles = (ctx->m + kDS - 1) / kDS;
    // sample ranges: fill whole waves of one CTA per SM, at least 4 tiles per range
    int ranges = 1;
    {
        const int64_t rmax = std::max<int64_t>(1, std::min<int64_t>(ntiles / 4, 4096));
        double best = 1e30;
        for (int r = 1; r <= rmax && r <= 64 * ctx->sm_count; ++r) {
            const int64_t ctas = (int64_t)r * cbn;
            if (ctas < ctx->sm_count && r < rmax) continue;
            const double waves = (double)((ctas + ctx->sm_count - 1) / ctx->sm_count);
            const double cost = waves * ctx->sm_count / (double)ctas + 1e-4 * r;  // idle fraction, mild preference for few ranges
            if (cost < best) best = cost, ranges = r;
            if (ctas > 8LL * ctx->sm_count) break;
        }
    }
    const size_t b_cd = (size_t)nd_pad * kpad * 8, b_eq = ((size_t)nd_pad * 4 + 15) & ~size_t(15);
    const size_t b_part = (size_t)ranges * nd_pad * 8;
    DDB_TRY(work.reserve(b_cd + b_eq + b_part));
    unsigned char* wb = static_cast<unsigned char*>(work.p);
    double* Cd = reinterpret_cast<double*>(wb);
    int32_t* eq_d = reinterpret_cast<int32_t*>(wb + b_cd);
    double* partials = reinterpret_cast<double*>(wb + b_cd + b_eq);
    dense_fill_kernel<<<nd_pad, 256, 0, st>>>(ids, nd, off, nnz, eq, idx, val, scale, map, kpad, Cd, eq_d);
    DenseArgs a;
    a.X = ctx->dX;
    a.Y = ctx->dY;
    a.n = ctx->n;
    a.n_t = ctx->n_t;
    a.m = ctx->m;
    a.k = k;
    a.kpad = kpad;
    a.factors = ctx->d_factors.as<uint16_t>();
    a.Cd = Cd;
    a.eq = eq_d;
    a.shift = shift;
    a.nd_pad = nd_pad;
    a.partials = partials;
    const size_t smem = 128 + (2 * (size_t)kDS * ctx->n + 2 * (size_t)kDC * kDP + 2 * (size_t)kDS * kDP) * sizeof(double);
    auto kern = ctx->max_degree <= 2 ? rss_dense_kernel<2> : rss_dense_kernel<kMaxFactors>;
    DDB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(ranges, cbn), 256, smem, st>>>(a);
    dense_reduce_kernel<<<(nd + 255) / 256, 256, 0, st>>>(partials, ranges, nd_pad, nd, out);
    DDB_CUDA_TRY(cudaGetLastError());
    *launches += 3;
    return DDB_OK;
}

}  // namespace ddb
