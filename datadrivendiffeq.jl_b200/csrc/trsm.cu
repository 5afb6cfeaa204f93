// Many right-hand sides against ONE Cholesky factor: blocked triangular solves whose bulk is FP64 DMMA GEMM.
//
// chol_solve_kernel (sweep.cu) gives every right-hand side its own CTA, which streams the whole factor from
// L2: fine for the 64 targets of a sweep, hopeless for k right-hand sides (the inverse of the ADMM / SR3
// iterations) or n = 4096 of them (K = Q P^-1 of the DMD path: 1024 CTAs x 134 MB).  Here the factor is cut
// into 128-wide block columns; with Linv_j = inverse of the j-th 128 x 128 diagonal block (computed once by
// tri_inv128_kernel) both substitution sweeps consist of products only:
//     forward  (L y = b):   Z_j <- Z_j Linv_j',   Z_rest -= Z_j L[rest, j]'
//     backward (L' x = y):  Z_j <- Z_j Linv_j,    Z_before -= Z_j L[j, before]     (through Lt = L')
// Z holds the right-hand sides as ROWS (nrhs x k, row-major), so every product has the form D (op)= A B' with
// the contraction index contiguous in both operands -- one kernel, gemm_nt_kernel: 128 x 128 tiles, 8 warps
// x (64 x 32 patch = 32 DMMA.8x8x4 accumulators), K in chunks of 32 through a double-buffered cp.async ring.
#include "common.cuh"
#include "mma_common.cuh"
#include <algorithm>

namespace ddb {

constexpr int kGT = 128;          // tile edge
constexpr int kGK = 32;           // contraction chunk
constexpr int kGP = kGK + 4;      // row pitch of an operand chunk: 36 = 4 (mod 16) -> conflict-free fragment loads
constexpr int kGStage = 2 * kGT * kGP;  // doubles per stage (A chunk + B chunk)

__device__ __forceinline__ void cp8(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// acc (this warp's 64 x 32 patch of the 128 x 128 tile at (r0, c0)) = sum_k A[r][k] B[c][k],  r < rows, c < cols, k < K:
// K in chunks of 32 through a double-buffered cp.async ring, 32 DMMA.8x8x4 accumulators per warp.
__device__ __forceinline__ void gemm_nt_tile(double (&acc)[8][4][2], double* gsm, const double* A, size_t lda,
                                             const double* __restrict__ B, size_t ldb, int rows, int cols, int K, int r0, int c0) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = (warp >> 2) * 64, col0 = (warp & 3) * 32;
    const int nchunks = (K + kGK - 1) / kGK;

    auto stage_load = [&](int chunk, int buf) {
        double* sa = gsm + (size_t)buf * kGStage;
        double* sb = sa + kGT * kGP;
        const int k0 = chunk * kGK;
        // 128 rows x 32 doubles per operand: thread -> (row = e / 32, k = e % 32), 16 elements per operand
#pragma unroll 4
        for (int e = tid; e < kGT * kGK; e += 256) {
            const int r = e >> 5, kk = e & 31;
            double* da = sa + r * kGP + kk;
            if (r0 + r < rows && k0 + kk < K) cp8(da, A + (size_t)(r0 + r) * lda + k0 + kk);
            else *da = 0.0;
            double* db = sb + r * kGP + kk;
            if (c0 + r < cols && k0 + kk < K) cp8(db, B + (size_t)(c0 + r) * ldb + k0 + kk);
            else *db = 0.0;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    stage_load(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) {
            stage_load(ch + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* sa = gsm + (size_t)buf * kGStage + (row0 + g) * kGP + t;
        const double* sb = gsm + (size_t)buf * kGStage + kGT * kGP + (col0 + g) * kGP + t;
#pragma unroll
        for (int kk = 0; kk < kGK / 4; ++kk) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = sa[8 * i * kGP + 4 * kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sb[8 * j * kGP + 4 * kk];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();  // the buffer is refilled by the next iteration's stage_load
    }
}

// D[r][c] (ASSIGN ? = : -=) sum_k A[r][k] B[c][k],  r < rows, c < cols, k < K.  D may alias A when cols <= 128
// (every CTA reads its rows of A completely before it writes).
template <bool ASSIGN>
__global__ void __launch_bounds__(256, 1) gemm_nt_kernel(double* __restrict__ D, size_t ldd, const double* A, size_t lda,
                                                         const double* __restrict__ B, size_t ldb, int rows, int cols, int K) {
    extern __shared__ __align__(16) double gsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.y * kGT, c0 = blockIdx.x * kGT;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = (warp >> 2) * 64, col0 = (warp & 3) * 32;
    double acc[8][4][2];
    gemm_nt_tile(acc, gsm, A, lda, B, ldb, rows, cols, K, r0, c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = r0 + row0 + 8 * i + g;
        if (r >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = c0 + col0 + 8 * j + 2 * t + h;
                if (c < cols) {
                    double* p = D + (size_t)r * ldd + c;
                    // subtract: one thread per element and launch -> a reduction (RED.ADD.F64) gives exactly D - acc
                    // without the load -> subtract -> store round trip
                    if (ASSIGN) *p = acc[i][j][h];
                    else atomicAdd(p, -acc[i][j][h]);
                }
            }
        }
    }
}

// partial[blockIdx.y * gridDim.x + blockIdx.x] = sum over the tile of (Z[r][c] - sum_k A[r][k] B[c][k])^2: the exact
// residual sum of squares of a linear model, prediction never written to memory (fixed-order sums: reproducible).
__global__ void __launch_bounds__(256, 1) gemm_nt_resid_kernel(const double* __restrict__ Z, size_t ldz, const double* A, size_t lda,
                                                               const double* __restrict__ B, size_t ldb, int rows, int cols, int K,
                                                               double* __restrict__ partial) {
    extern __shared__ __align__(16) double gsm[];
    __shared__ double wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.y * kGT, c0 = blockIdx.x * kGT;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = (warp >> 2) * 64, col0 = (warp & 3) * 32;
    double acc[8][4][2];
    gemm_nt_tile(acc, gsm, A, lda, B, ldb, rows, cols, K, r0, c0);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = r0 + row0 + 8 * i + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = c0 + col0 + 8 * j + 2 * t + h;
                if (r < rows && c < cols) {
                    const double e = Z[(size_t)r * ldz + c] - acc[i][j][h];
                    s = fma(e, e, s);
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) wsum[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += wsum[w];
        partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
}

__global__ void sum_partials_kernel(const double* __restrict__ partial, int64_t count, double* __restrict__ out) {
    // one CTA, fixed order: thread t adds entries t, t + 1024, ...; then a fixed tree
    __shared__ double red[1024];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < count; i += 1024) s += partial[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

// Inverse of every 128-wide diagonal block of the lower factor L (block b: rows/cols 128 b ...): thread c
// substitutes column c.  Writes Linv (row-major 128 x 128 per block, zero above the diagonal and in the padding)
// and its transpose.
__global__ void __launch_bounds__(128) tri_inv128_kernel(const double* __restrict__ L, size_t ld, int k,
                                                         double* __restrict__ Linv, double* __restrict__ LinvT) {
    extern __shared__ double tsm[];
    double* Xs = tsm;  // 128 x 129, column c owned by thread c; the factor itself is read through L1 (all threads
                       // of the CTA read the same element: one broadcast load)
    const int b = blockIdx.x, j0 = b * kGT, nb = min(kGT, k - j0);
    const int c = threadIdx.x;
    for (int e = c; e < kGT * (kGT + 1); e += blockDim.x) Xs[e] = 0.0;
    __syncthreads();
    if (c < nb) {
        for (int i = c; i < nb; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            const double* lr = L + (size_t)(j0 + i) * ld + j0;
            for (int tt = c; tt < i; ++tt) s = fma(-__ldg(lr + tt), Xs[tt * (kGT + 1) + c], s);
            Xs[i * (kGT + 1) + c] = s / __ldg(lr + i);
        }
    }
    __syncthreads();
    double* o = Linv + (size_t)b * kGT * kGT;
    double* ot = LinvT + (size_t)b * kGT * kGT;
    for (int e = c; e < kGT * kGT; e += blockDim.x) {
        const int i = e / kGT, j = e % kGT;
        const double v = Xs[i * (kGT + 1) + j];
        o[e] = v;
        ot[j * kGT + i] = v;
    }
}

// Lt = L' (only the part above the diagonal of Lt is read later: Lt[i][j] = L[j][i], j > i)
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ L, size_t ld, int k, double* __restrict__ Lt,
                                                        size_t ldt) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    if (bx > by + 31) return;  // L is lower triangular: tiles strictly above the diagonal hold nothing
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int q = ty; q < 32; q += 8) {
        const int r = by + q, cidx = bx + tx;
        tile[q][tx] = (r < k && cidx < k && cidx <= r) ? L[(size_t)r * ld + cidx] : 0.0;
    }
    __syncthreads();
    for (int q = ty; q < 32; q += 8) {
        const int r = bx + q, cidx = by + tx;  // Lt[r][cidx] = L[cidx][r]
        if (r < k && cidx < k) Lt[(size_t)r * ldt + cidx] = tile[tx][q];
    }
}

struct TrsmState {
    DevBuf linv, linvT, lt;
    // chol_factor_shared (below) leaves L' above the diagonal of ITS factor: solves against that factor need no
    // transposition pass; the inverses of its diagonal blocks are computed on the first such solve
    DevBuf f_linv, f_linvT;
    const double* factor = nullptr;
    int factor_k = 0;
    bool factor_inv = false;
};

int chol128_batched(ddb_ctx* ctx, double* base, size_t stride, int ld, const int* d_na, int na_max, int batch, int* d_fail,
                    int extra, int xrow, int* launches);

template <bool ASSIGN>
static int launch_gemm(ddb_ctx* ctx, double* D, size_t ldd, const double* A, size_t lda, const double* B, size_t ldb,
                       int rows, int cols, int K) {
    if (rows <= 0 || cols <= 0 || K <= 0) return DDB_OK;
    const size_t smem = 2 * (size_t)kGStage * sizeof(double);
    // (per device and cheap: no caching, a process may drive several GPUs through several contexts)
    DDB_CUDA_TRY(cudaFuncSetAttribute(gemm_nt_kernel<ASSIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_nt_kernel<ASSIGN><<<dim3((cols + kGT - 1) / kGT, (rows + kGT - 1) / kGT), 256, smem, ctx->stream>>>(D, ldd, A, lda, B, ldb,
                                                                                                          rows, cols, K);
    return DDB_OK;
}

// Z (nrhs x k, row-major, leading dimension k) <- Z (L L')^-1  i.e. every ROW of Z is solved against the factor
// L (k x k lower, row-major, leading dimension ld).  mirrored: L' lives above the diagonal of the same buffer (as
// chol128_batched leaves it) -- no transposition pass; linv / linvT: where the inverses of the 128-wide diagonal
// blocks and their transposes are (or go); have_inv: they are already there.
static int trsm_gemm_run(ddb_ctx* ctx, const double* L, size_t ld, int k, double* Z, int nrhs, int* launches, bool mirrored,
                         double* linv, double* linvT, bool have_inv, DevBuf* lt_buf) {
    const int nblk = (k + kGT - 1) / kGT;
    cudaStream_t st = ctx->stream;
    int n = 0;
    if (!have_inv) {
        const size_t tsmem = (size_t)kGT * (kGT + 1) * sizeof(double);
        DDB_CUDA_TRY(cudaFuncSetAttribute(tri_inv128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        tri_inv128_kernel<<<nblk, 128, tsmem, st>>>(L, ld, k, linv, linvT);
        ++n;
    }
    const double* Lt = L;  // Lt[i][j] = L[j][i] for j > i
    size_t ldt = ld;
    if (!mirrored) {
        ldt = (size_t)k;
        DDB_TRY(lt_buf->reserve((size_t)k * ldt * 8));
        transpose_kernel<<<dim3((k + 31) / 32, (k + 31) / 32), 256, 0, st>>>(L, ld, k, lt_buf->as<double>(), ldt);
        ++n;
        Lt = lt_buf->as<double>();
    }
    const size_t zs = (size_t)k;
    for (int b = 0; b < nblk; ++b) {  // forward
        const int j0 = b * kGT, nb = std::min(kGT, k - j0), rest = k - j0 - nb;
        DDB_TRY(launch_gemm<true>(ctx, Z + j0, zs, Z + j0, zs, linv + (size_t)b * kGT * kGT, kGT, nrhs, nb, nb));
        DDB_TRY(launch_gemm<false>(ctx, Z + j0 + nb, zs, Z + j0, zs, L + (size_t)(j0 + nb) * ld + j0, ld, nrhs, rest, nb));
        n += 1 + (rest > 0);
    }
    for (int b = nblk - 1; b >= 0; --b) {  // backward
        const int j0 = b * kGT, nb = std::min(kGT, k - j0);
        DDB_TRY(launch_gemm<true>(ctx, Z + j0, zs, Z + j0, zs, linvT + (size_t)b * kGT * kGT, kGT, nrhs, nb, nb));
        DDB_TRY(launch_gemm<false>(ctx, Z, zs, Z + j0, zs, Lt + j0, ldt, nrhs, j0, nb));
        n += 1 + (j0 > 0);
    }
    DDB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += n;
    return DDB_OK;
}

// One matrix, many right-hand sides (the factor every target shares; P of the DMD operator): factorise A (k x k,
// row-major, both triangles valid on entry) in place; L' is left above the diagonal.  `extra` right-hand sides
// stored as rows k .. k + extra - 1 of the same buffer are forward-substituted on the way (they end up as L^-1 b;
// chol_back_batched finishes them).
int chol_factor_shared(ddb_ctx* ctx, double* A, int k, int* d_fail, int extra, int* launches) {
    if (!ctx->trsm) ctx->trsm = new TrsmState();
    TrsmState* T = ctx->trsm;
    T->factor = nullptr;
    DDB_TRY(chol128_batched(ctx, A, 0, k, nullptr, k, 1, d_fail, extra, k, launches));
    T->factor = A;
    T->factor_k = k;
    T->factor_inv = false;
    return DDB_OK;
}

// Z (nrhs x k row-major) <- Z (L L')^-1 for MANY right-hand sides (the inverse of the ADMM / SR3 iterations,
// K = Q P^-1 of the DMD path): blocked substitution whose bulk is DMMA GEMM, against a factor of chol_factor_shared
// (its mirrored L' is used, the block inverses are computed once per factor) or any other row-major lower factor.
int chol_solve_many(ddb_ctx* ctx, const double* L, int k, double* Z, int nrhs, int* launches) {
    if (!ctx->trsm) ctx->trsm = new TrsmState();
    TrsmState* T = ctx->trsm;
    const int nblk = (k + kGT - 1) / kGT;
    if (T->factor == L && T->factor_k == k) {
        DDB_TRY(T->f_linv.reserve((size_t)nblk * kGT * kGT * 8));
        DDB_TRY(T->f_linvT.reserve((size_t)nblk * kGT * kGT * 8));
        const bool have = T->factor_inv;
        T->factor_inv = true;
        return trsm_gemm_run(ctx, L, (size_t)k, k, Z, nrhs, launches, true, T->f_linv.as<double>(), T->f_linvT.as<double>(), have,
                             &T->lt);
    }
    DDB_TRY(T->linv.reserve((size_t)nblk * kGT * kGT * 8));
    DDB_TRY(T->linvT.reserve((size_t)nblk * kGT * kGT * 8));
    return trsm_gemm_run(ctx, L, (size_t)k, k, Z, nrhs, launches, false, T->linv.as<double>(), T->linvT.as<double>(), false, &T->lt);
}

// rss = || Z - A B' ||_F^2 for row-major A (rows x K), B (cols x K), Z (rows x cols); device scalar out
int gemm_nt_residual(ddb_ctx* ctx, const double* Z, size_t ldz, const double* A, size_t lda, const double* B, size_t ldb, int64_t rows,
                     int cols, int K, double* d_partials, double* d_out) {
    const size_t smem = 2 * (size_t)kGStage * sizeof(double);
    DDB_CUDA_TRY(cudaFuncSetAttribute(gemm_nt_resid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ty = (rows + kGT - 1) / kGT;
    const int tx = (cols + kGT - 1) / kGT;
    int64_t done = 0;
    // the grid's y extent is limited to 65535 tiles: slabs of rows, partial sums appended
    for (int64_t y0 = 0; y0 < ty; y0 += 65535) {
        const int64_t ny = std::min<int64_t>(65535, ty - y0);
        const int64_t r_lo = y0 * kGT, r_n = std::min<int64_t>(rows - r_lo, ny * kGT);
        gemm_nt_resid_kernel<<<dim3(tx, (unsigned)ny), 256, smem, ctx->stream>>>(Z + r_lo * ldz, ldz, A + r_lo * lda, lda, B, ldb,
                                                                                 (int)r_n, cols, K, d_partials + done);
        done += ny * tx;
    }
    sum_partials_kernel<<<1, 1024, 0, ctx->stream>>>(d_partials, done, d_out);
    DDB_CUDA_TRY(cudaGetLastError());
    return DDB_OK;
}

void trsm_free(ddb_ctx* ctx) {
    if (ctx->trsm) {
        ctx->trsm->linv.release();
        ctx->trsm->linvT.release();
        ctx->trsm->lt.release();
        ctx->trsm->f_linv.release();
        ctx->trsm->f_linvT.release();
        delete ctx->trsm;
        ctx->trsm = nullptr;
    }
}

}  // namespace ddb

extern "C" int ddb_gemm_nt(ddb_ctx* ctx, double* dD, int64_t ldd, const double* dA, int64_t lda, const double* dB, int64_t ldb,
                           int rows, int cols, int K) {
    using namespace ddb;
    if (!ctx || !dD || !dA || !dB || rows <= 0 || cols <= 0 || K <= 0 || ldd < cols || lda < K || ldb < K) {
        set_error("ddb_gemm_nt: null pointer, non-positive size or leading dimension smaller than the row length");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    DDB_TRY(launch_gemm<true>(ctx, dD, (size_t)ldd, dA, (size_t)lda, dB, (size_t)ldb, rows, cols, K));
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DDB_OK;
}
