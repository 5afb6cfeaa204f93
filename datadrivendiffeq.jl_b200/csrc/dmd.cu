// DataDrivenDMD Gram path on FP64 DMMA tensor cores.
//
// Replaces `Q = Y_ * X'`, `P = X * X'` (reference lib/DataDrivenDMD/src/solve.jl:128-129) and, through
// the normal equations, `K = Y / X` of DMDPINV (algorithms.jl:80-83: K = Q P^-1).  X and Y are the
// lifted snapshot matrices (n x m, column-major = snapshot-contiguous); for a discrete problem Y is X
// shifted by one column (solve.jl:36-65), so the caller passes dY = dX + n and one resident copy serves
// both.
//
// Kernel: the augmented rows z = [x; y] (2n of them) are tiled into 128-row blocks; P needs the lower
// triangle of the x-x block pairs, Q all y-x block pairs (y-y pairs are skipped).  The library is the
// identity, so a Theta tile IS a slab of the snapshot matrix: 16 snapshots x 128 rows, which the TMA unit
// copies straight into the fragment layout (16 bulk copies of 1 KB, pitch 132 doubles) -- no build step,
// no FP64 multiplies outside the tensor pipe.  Warp-specialised: one producer warp drives a 4-deep ring of
// tile pairs with full/empty mbarriers; eight MMA warps own 64 x 32 accumulator patches (DMMA.8x8x4), with
// the same diagonal-pair sub-tile skipping, persistent cost-balanced plan and fixed-order partial-tile
// reduction as gram.cu.
#include "common.cuh"
#include "mma_common.cuh"
#include "sweep.cuh"
#include <algorithm>
#include <cstring>

namespace ddb {

constexpr int kDmdBT = 128;
constexpr int kDmdPitch = kDmdBT + 4;
constexpr int kDmdRing = 4;
constexpr int kDmdStageDoubles = 2 * kKS * kDmdPitch;
constexpr int kDmdMmaWarps = 8;
constexpr int kDmdThreads = (kDmdMmaWarps + 1) * 32;

struct DmdArgs {
    const double* X;
    const double* Y;
    int n, nblk;  // nblk = blocks per matrix; augmented block a < nblk is X block a, else Y block a - nblk
    int64_t m;
    const GramSegment* segs;
    const int32_t* cta_seg_begin;
    double* partials;
};

template <uint32_t MASK>
__device__ __forceinline__ void dmd_mma_stage(double (&acc)[8][4][2], const double* __restrict__ pa,
                                              const double* __restrict__ pb) {
#pragma unroll
    for (int kk = 0; kk < kKS / 4; ++kk) {
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if ((MASK >> (4 * i)) & 0xFu) a[i] = pa[kk * 4 * kDmdPitch + 8 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((MASK >> j) & 0x11111111u) b[j] = pb[kk * 4 * kDmdPitch + 8 * j];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((MASK >> (i * 4 + j)) & 1u) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

__global__ void __launch_bounds__(kDmdThreads, 1) dmd_gram_kernel(const DmdArgs args) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // [ring] tile pair landed
    uint64_t* empty = full + kDmdRing;                       // [ring] all MMA warps done with the slot
    double* ring = reinterpret_cast<double*>(smem_raw + 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < kDmdRing; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kDmdMmaWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n = args.n;
    const int64_t nstages_total = (args.m + kKS - 1) / kKS;
    const int seg_begin = args.cta_seg_begin[blockIdx.x];
    const int seg_end = args.cta_seg_begin[blockIdx.x + 1];

    if (warp == kDmdMmaWarps) {
        // ================= producer warp: lane l copies snapshot row l % 16 of tile l / 16 =================
        int item = 0;
        for (int sg = seg_begin; sg < seg_end; ++sg) {
            const GramSegment seg = args.segs[sg];
            const bool same = (seg.bi == seg.bj);
            const int tile = lane >> 4, srow = lane & 15;
            const int blk = tile == 0 ? seg.bi : seg.bj;
            const double* src = (blk < args.nblk ? args.X : args.Y) + (size_t)(blk % args.nblk) * kDmdBT;
            const int width = min(kDmdBT, n - (blk % args.nblk) * kDmdBT);  // rows of this block
            const int wi = min(kDmdBT, n - (seg.bi % args.nblk) * kDmdBT), wj = min(kDmdBT, n - (seg.bj % args.nblk) * kDmdBT);
            for (int64_t st = seg.stage_begin; st < seg.stage_end; ++st, ++item) {
                const int slot = item & (kDmdRing - 1);
                if (item >= kDmdRing) mbar_wait(&empty[slot], (uint32_t)(((item / kDmdRing) & 1) ^ 1));
                double* dst = ring + (size_t)slot * kDmdStageDoubles + tile * kKS * kDmdPitch + srow * kDmdPitch;
                const int64_t s = st * kKS + srow;
                const int valid = (int)min((int64_t)kKS, args.m - st * kKS);
                if (valid < kKS && srow >= valid && (tile == 0 || !same)) {
                    for (int c = 0; c < width; ++c) dst[c] = 0.0;  // ragged last stage: zero the missing snapshots
                }
                __syncwarp();
                if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)(valid * (wi + (same ? 0 : wj)) * sizeof(double)));
                __syncwarp();
                if (srow < valid && (tile == 0 || !same)) bulk_g2s(dst, src + s * n, (uint32_t)(width * sizeof(double)), &full[slot]);
            }
        }
    } else {
        // ================= MMA warps =====================================================================
        const int wm = (0x4B >> warp) & 1;  // warps 0..7 -> (1,0)(1,1)(0,0)(1,2)(0,2)(0,3)(1,3)(0,1), see gram.cu
        const int wn = (0x7E84 >> (2 * warp)) & 3;
        const int g = lane >> 2, t = lane & 3;
        const int row0 = wm * 64, col0 = wn * 32;
        int item = 0;
        for (int sg = seg_begin; sg < seg_end; ++sg) {
            const GramSegment seg = args.segs[sg];
            const bool same = (seg.bi == seg.bj);
            int shape = 0;
            if (same) {
                const int d = (row0 - col0) / 8;
                shape = (d >= 3) ? 0 : (d == 0 ? 1 : (d == -4 ? 2 : 3));
            }
            double acc[8][4][2];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            for (int64_t st = seg.stage_begin; st < seg.stage_end; ++st, ++item) {
                const int slot = item & (kDmdRing - 1);
                mbar_wait(&full[slot], (uint32_t)((item / kDmdRing) & 1));
                const double* tI = ring + (size_t)slot * kDmdStageDoubles;
                const double* tJ = same ? tI : tI + kKS * kDmdPitch;
                const double* pa = tI + t * kDmdPitch + row0 + g;
                const double* pb = tJ + t * kDmdPitch + col0 + g;
                switch (shape) {
                    case 0: dmd_mma_stage<kShapeFull>(acc, pa, pb); break;
                    case 1: dmd_mma_stage<kShapeDiag0>(acc, pa, pb); break;
                    case 2: dmd_mma_stage<kShapeDiagM4>(acc, pa, pb); break;
                    default: break;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
            }
            const uint32_t mask = shape == 0 ? kShapeFull : shape == 1 ? kShapeDiag0 : shape == 2 ? kShapeDiagM4 : 0u;
            double* out = args.partials + (size_t)seg.slot * kDmdBT * kDmdBT;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = row0 + 8 * i + g;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = col0 + 8 * j + 2 * t;
                    if ((mask >> (i * 4 + j)) & 1)
                        *reinterpret_cast<double2*>(out + (size_t)row * kDmdBT + c) = make_double2(acc[i][j][0], acc[i][j][1]);
                }
            }
        }
    }
}

// adds the partial tiles of each block pair in slot order and scatters into P (symmetric) / Q
__global__ void __launch_bounds__(256) dmd_reduce_kernel(const double* __restrict__ partials,
                                                         const int2* __restrict__ pairs,
                                                         const int32_t* __restrict__ pair_slot_begin, int n, int nblk,
                                                         double* __restrict__ P, double* __restrict__ Q) {
    const int p = blockIdx.x;
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    if (e >= kDmdBT * kDmdBT) return;
    const int i = e / kDmdBT, j = e % kDmdBT;
    const int bi = pairs[p].x, bj = pairs[p].y;
    const int gj = bj * kDmdBT + j;
    const bool isq = bi >= nblk;
    const int gi = (isq ? bi - nblk : bi) * kDmdBT + i;
    if (gi >= n || gj >= n) return;
    if (!isq && gj > gi) return;
    double s = 0.0;
    for (int sl = pair_slot_begin[p]; sl < pair_slot_begin[p + 1]; ++sl) s += partials[(size_t)sl * kDmdBT * kDmdBT + e];
    if (isq) {
        Q[(size_t)gi + (size_t)n * gj] = s;  // Q = Y X'
    } else {
        P[(size_t)gi + (size_t)n * gj] = s;
        P[(size_t)gj + (size_t)n * gi] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// K = Q P^-1 through the equilibrated Cholesky of P (kernels shared with the sweep, sweep.cu)
// ---------------------------------------------------------------------------------------------
int chol_factor_shared(ddb_ctx* ctx, double* A, int k, int* d_fail, int extra, int* launches);
int chol_solve_many(ddb_ctx* ctx, const double* L, int k, double* Z, int nrhs, int* launches);

__global__ void dmd_diag_kernel(const double* __restrict__ P, int n, double* __restrict__ dinv) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double a = P[(size_t)j * n + j];
    dinv[j] = (a > 0.0) ? 1.0 / sqrt(a) : 0.0;
}
__global__ void dmd_scale_kernel(const double* __restrict__ P, int n, const double* __restrict__ dinv,
                                 double* __restrict__ Ps) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    const int i = (int)(e / n), j = (int)(e % n);
    double a = P[e] * dinv[i] * dinv[j];
    if (i == j && dinv[i] == 0.0) a = 1.0;
    Ps[e] = a;
}
// Z[i][j] = Q[i + n j] * dinv[j]   (row i of Q, scaled, stored contiguously)
__global__ void dmd_rhs_kernel(const double* __restrict__ Q, int n, const double* __restrict__ dinv,
                               double* __restrict__ Z) {
    __shared__ double tile[32][33];
    const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + tx, j = j0 + r;
        tile[r][tx] = (i < n && j < n) ? Q[(size_t)i + (size_t)n * j] * dinv[j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + tx;
        if (i < n && j < n) Z[(size_t)i * n + j] = tile[tx][r];
    }
}
// K[i + n j] = W[i][j] * dinv[j]
__global__ void dmd_out_kernel(const double* __restrict__ W, int n, const double* __restrict__ dinv,
                               double* __restrict__ K) {
    __shared__ double tile[32][33];
    const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + tx;
        tile[r][tx] = (i < n && j < n) ? W[(size_t)i * n + j] * dinv[j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + tx, j = j0 + r;
        if (i < n && j < n) K[(size_t)i + (size_t)n * j] = tile[tx][r];
    }
}

struct DmdState {
    DevBuf plan, partials, Ps, dinv, Z, fail, hX, hPQK;
    GramPlan gp;
    int plan_n = -1;
    int64_t plan_m = -1;
};

}  // namespace ddb

using namespace ddb;

extern "C" {

int ddb_dmd_gram(ddb_ctx* ctx, const double* dX, const double* dY, int n, int64_t m, double* dP, double* dQ) {
    if (!ctx || !dX || !dY || !dP || !dQ || n <= 0 || m <= 0) {
        set_error("ddb_dmd_gram: null pointer or non-positive size");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    if ((n & 1) || ((uintptr_t)dX & 15) || ((uintptr_t)dY & 15)) {
        set_error("ddb_dmd_gram: n must be even and the snapshot pointers 16-byte aligned (TMA bulk copies)");
        return DDB_UNSUPPORTED;
    }
    if (!ctx->dmd) ctx->dmd = new DmdState();
    DmdState* D = ctx->dmd;
    const int nblk = (n + kDmdBT - 1) / kDmdBT;
    if (D->plan_n != n || D->plan_m != m) {
        std::vector<std::pair<int, int>> pairs;
        std::vector<double> cost;
        for (int bi = 0; bi < nblk; ++bi)
            for (int bj = 0; bj <= bi; ++bj) {
                pairs.push_back({bi, bj});
                cost.push_back(bi == bj ? 0.58 : 1.0);
            }
        for (int bi = 0; bi < nblk; ++bi)
            for (int bj = 0; bj < nblk; ++bj) {
                pairs.push_back({nblk + bi, bj});
                cost.push_back(1.0);
            }
        build_plan_from_pairs(D->gp, pairs, cost, m, ctx->sm_count, kDmdBT, kKS);
        const GramPlan& pl = D->gp;
        std::vector<int2> p2;
        for (auto& pr : pl.pairs) p2.push_back(make_int2(pr.first, pr.second));
        size_t off_cta = pl.segments.size() * sizeof(GramSegment);
        size_t off_pairs = (off_cta + pl.cta_seg_begin.size() * sizeof(int32_t) + 15) & ~size_t(15);
        size_t off_psb = off_pairs + p2.size() * sizeof(int2);
        size_t total = off_psb + pl.pair_slot_begin.size() * sizeof(int32_t);
        std::vector<unsigned char> host(total);
        memcpy(host.data(), pl.segments.data(), pl.segments.size() * sizeof(GramSegment));
        memcpy(host.data() + off_cta, pl.cta_seg_begin.data(), pl.cta_seg_begin.size() * sizeof(int32_t));
        memcpy(host.data() + off_pairs, p2.data(), p2.size() * sizeof(int2));
        memcpy(host.data() + off_psb, pl.pair_slot_begin.data(), pl.pair_slot_begin.size() * sizeof(int32_t));
        DDB_TRY(D->plan.reserve(total));
        DDB_CUDA_TRY(cudaMemcpyAsync(D->plan.p, host.data(), total, cudaMemcpyHostToDevice, ctx->stream));
        DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        DDB_TRY(D->partials.reserve((size_t)pl.nslots * kDmdBT * kDmdBT * sizeof(double)));
        D->plan_n = n;
        D->plan_m = m;
    }
    const GramPlan& pl = D->gp;
    unsigned char* base = static_cast<unsigned char*>(D->plan.p);
    size_t off_cta = pl.segments.size() * sizeof(GramSegment);
    size_t off_pairs = (off_cta + pl.cta_seg_begin.size() * sizeof(int32_t) + 15) & ~size_t(15);
    size_t off_psb = off_pairs + (size_t)pl.npairs * sizeof(int2);
    DmdArgs a;
    a.X = dX;
    a.Y = dY;
    a.n = n;
    a.nblk = nblk;
    a.m = m;
    a.segs = reinterpret_cast<const GramSegment*>(base);
    a.cta_seg_begin = reinterpret_cast<const int32_t*>(base + off_cta);
    a.partials = D->partials.as<double>();
    const size_t smem = 128 + (size_t)kDmdRing * kDmdStageDoubles * sizeof(double);
    DDB_CUDA_TRY(cudaFuncSetAttribute(dmd_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    dmd_gram_kernel<<<pl.nctas, kDmdThreads, smem, ctx->stream>>>(a);
    dim3 rgrid(pl.npairs, (kDmdBT * kDmdBT + 255) / 256);
    dmd_reduce_kernel<<<rgrid, 256, 0, ctx->stream>>>(D->partials.as<double>(), reinterpret_cast<const int2*>(base + off_pairs),
                                                      reinterpret_cast<const int32_t*>(base + off_psb), n, nblk, dP, dQ);
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->timings.gram_ms = ms;
    ctx->timings.gram_launches = 2;
    return DDB_OK;
}

int ddb_dmd_operator(ddb_ctx* ctx, const double* dP, const double* dQ, int n, double* dK) {
    if (!ctx || !dP || !dQ || !dK || n <= 0) {
        set_error("ddb_dmd_operator: null pointer or non-positive size");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (!ctx->dmd) ctx->dmd = new DmdState();
    DmdState* D = ctx->dmd;
    const size_t nn = (size_t)n * n;
    DDB_TRY(D->Ps.reserve(nn * 8));
    DDB_TRY(D->Z.reserve(nn * 8));
    DDB_TRY(D->dinv.reserve((size_t)n * 8));
    DDB_TRY(D->fail.reserve(16));
    cudaStream_t st = ctx->stream;
    DDB_CUDA_TRY(cudaMemsetAsync(D->fail.p, 0, 16, st));
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[4], st));
    dmd_diag_kernel<<<(n + 255) / 256, 256, 0, st>>>(dP, n, D->dinv.as<double>());
    dmd_scale_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(dP, n, D->dinv.as<double>(), D->Ps.as<double>());
    dim3 tg((n + 31) / 32, (n + 31) / 32);
    dmd_rhs_kernel<<<tg, 256, 0, st>>>(dQ, n, D->dinv.as<double>(), D->Z.as<double>());
    int launches = 3;
    DDB_TRY(chol_factor_shared(ctx, D->Ps.as<double>(), n, D->fail.as<int>(), 0, &launches));
    DDB_TRY(chol_solve_many(ctx, D->Ps.as<double>(), n, D->Z.as<double>(), n, &launches));
    dmd_out_kernel<<<tg, 256, 0, st>>>(D->Z.as<double>(), n, D->dinv.as<double>(), dK);
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[5], st));
    int h_fail = 0;
    DDB_CUDA_TRY(cudaMemcpyAsync(&h_fail, D->fail.p, 4, cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    ctx->timings.factor_ms = ms;
    ctx->timings.sweep_launches = launches + 2;
    if (h_fail) {
        set_error("ddb_dmd_operator: X X' is not positive definite (rank-deficient snapshot matrix); the reference's "
                  "pivoted QR handles this case, the Gram path does not");
        return DDB_NOT_POSITIVE_DEFINITE;
    }
    return DDB_OK;
}

// Host-buffer form of the exact-DMD Gram path (what a `ccall` from Julia binds): snapshots x (n x (m + 1),
// column-major, one snapshot per column) -> P = X X', Q = Y X' with X = x[:, 1:m], Y = x[:, 2:m+1]
// (lib/DataDrivenDMD/src/solve.jl:36-65, 128-129) and, when K is requested, K = Q P^-1 (DMDPINV,
// algorithms.jl:80-83).  P, Q, K: n x n column-major host buffers, each optional.
int ddb_dmd_solve(ddb_ctx* ctx, const double* x, int n, int64_t m_plus_1, double* P, double* Q, double* K) {
    if (!ctx || !x || n <= 0 || m_plus_1 < 2) {
        set_error("ddb_dmd_solve: null pointer, non-positive size or fewer than two snapshots");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (!ctx->dmd) ctx->dmd = new DmdState();
    DmdState* D = ctx->dmd;
    const int64_t m = m_plus_1 - 1;
    const size_t nn = (size_t)n * n;
    DDB_TRY(D->hX.reserve((size_t)n * m_plus_1 * 8));
    DDB_TRY(D->hPQK.reserve(3 * nn * 8));
    cudaStream_t st = ctx->stream;
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[2], st));
    DDB_CUDA_TRY(cudaMemcpyAsync(D->hX.p, x, (size_t)n * m_plus_1 * 8, cudaMemcpyHostToDevice, st));
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[3], st));
    double* dP = D->hPQK.as<double>();
    double* dQ = dP + nn;
    double* dK = dQ + nn;
    DDB_TRY(ddb_dmd_gram(ctx, D->hX.as<double>(), D->hX.as<double>() + n, n, m, dP, dQ));
    if (K) DDB_TRY(ddb_dmd_operator(ctx, dP, dQ, n, dK));
    if (P) DDB_CUDA_TRY(cudaMemcpyAsync(P, dP, nn * 8, cudaMemcpyDeviceToHost, st));
    if (Q) DDB_CUDA_TRY(cudaMemcpyAsync(Q, dQ, nn * 8, cudaMemcpyDeviceToHost, st));
    if (K) DDB_CUDA_TRY(cudaMemcpyAsync(K, dK, nn * 8, cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
    ctx->timings.h2d_ms = ms;
    return DDB_OK;
}

}  // extern "C"

namespace ddb {
void dmd_free(ddb_ctx* ctx) {
    if (ctx->dmd) {
        DmdState* D = ctx->dmd;
        D->plan.release();
        D->partials.release();
        D->Ps.release();
        D->dinv.release();
        D->Z.release();
        D->fail.release();
        D->hX.release();
        D->hPQK.release();
        delete D;
        ctx->dmd = nullptr;
    }
}
}  // namespace ddb
