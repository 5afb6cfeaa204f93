// Thresholded sweep of DataDrivenSparse on the normal-equation blocks, device resident.
//
// Replaces `SparseLinearSolver` (reference lib/DataDrivenSparse/src/solver.jl:72-118) and the
// `init_cache` / `step!` methods of STLSQ (algorithms/STLSQ.jl:81-140), ADMM (ADMM.jl:68-121, skinny
// branch) and SR3 (SR3.jl:98-141) for ALL targets at once, working on G = Theta Theta', R = Theta Y'
// instead of Theta itself:
//   * the system matrix A = G (+ rho I) is equilibrated, As = D^-1 A D^-1 with D = sqrt(diag A)
//     (mandatory: cond(G) = cond(Theta)^2); thresholds act on the un-scaled
//     coefficients;
//   * one blocked Cholesky of As is shared by every target (initial full least squares of STLSQ,
//     every iteration of ADMM / SR3);
//   * STLSQ's per-target active-set solves run in ONE persistent kernel, one CTA per target, looping
//     thresholds and inner iterations on the device: the masked system is gathered into shared
//     memory, factorised and solved there, the support is re-thresholded, convergence and the
//     candidate bookkeeping for the selector are done in place.  Supports with more than
//     kSmallMax active terms fall back to a batched blocked Cholesky in global memory;
//   * every distinct (support, coefficients) a target visits is recorded as a sparse candidate; a
//     second streaming pass over the samples (rss.cu) scores them exactly and `select` replays the
//     selector along the recorded step sequence (solver.jl:100-107, incl. `<=` and `j == 1`).
#include "common.cuh"
#include "sweep.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace ddb {

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* red) {  // blockDim.x multiple of 32, <= 1024
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += red[w];  // fixed order: deterministic
    return s;
}
__device__ __forceinline__ int block_sum_int(int v, int* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    int s = 0;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}

// ---------------------------------------------------------------------------------------------
// preparation: A = G + rho I, equilibrate, scaled right-hand sides
// ---------------------------------------------------------------------------------------------
// check_domain of the reference (src/problem/type.jl:418-420) for data that never passed through the host
// mirror: NaN / Inf samples end up in the normal-equation blocks, where one pass finds them
__global__ void nonfinite_check_kernel(const double* __restrict__ packed, size_t count, int* __restrict__ flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count && !isfinite(packed[i])) atomicExch(flag, 1);
}

__global__ void prep_diag_kernel(const double* __restrict__ G, int k, double rho, double* __restrict__ dinv) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const double a = G[(size_t)j * k + j] + rho;
    dinv[j] = (a > 0.0) ? 1.0 / sqrt(a) : 0.0;  // an identically-zero library term gets coefficient 0
}
__global__ void prep_scale_kernel(const double* __restrict__ G, int k, double rho, const double* __restrict__ dinv,
                                  double* __restrict__ As) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)k * k) return;
    const int i = (int)(e / k), j = (int)(e % k);
    double a = G[e] + (i == j ? rho : 0.0);
    a *= dinv[i] * dinv[j];
    if (i == j && dinv[i] == 0.0) a = 1.0;
    As[e] = a;
}
// rs[e][j] = dinv[j] * R[j + k e]
__global__ void prep_rhs_kernel(const double* __restrict__ R, int k, int n_t, const double* __restrict__ dinv,
                                double* __restrict__ rs) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)k * n_t) return;
    rs[e] = R[e] * dinv[e % k];
}

// ---------------------------------------------------------------------------------------------
// batched blocked Cholesky, row-major lower triangle in place (leading dimension ld)
//   matrix b lives at base + b * stride; its order is na[b] (or `na_all` when na == nullptr).
// ---------------------------------------------------------------------------------------------
constexpr int kNB = 32;

// Factor the nb x nb diagonal block at (j0, j0) in place (whole CTA, D is shared scratch).
__device__ void factor_diag_block(double* A, int ld, int j0, int nb, int* fail_flag, double (*D)[kNB + 1]) {
    const int tid = threadIdx.x;
    for (int e = tid; e < kNB * kNB; e += blockDim.x) {
        const int i = e / kNB, j = e % kNB;
        double v = (i == j) ? 1.0 : 0.0;
        if (i < nb && j < nb && j <= i) v = A[(size_t)(j0 + i) * ld + j0 + j];
        D[i][j] = v;
    }
    __syncthreads();
    for (int c = 0; c < nb; ++c) {
        if (tid == 0) {
            const double d = D[c][c];
            if (!(d > 0.0)) {
                atomicExch(fail_flag, 1);
                D[c][c] = 1.0;  // keep going with a harmless pivot; the flag invalidates the result
            } else {
                D[c][c] = sqrt(d);
            }
        }
        __syncthreads();
        if (tid > c && tid < nb) D[tid][c] /= D[c][c];
        __syncthreads();
        for (int e = tid; e < nb * nb; e += blockDim.x) {
            const int i = e / nb, j = e % nb;
            if (j > c && i >= j) D[i][j] -= D[i][c] * D[j][c];
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += blockDim.x) {
        const int i = e / nb, j = e % nb;
        if (j <= i) A[(size_t)(j0 + i) * ld + j0 + j] = D[i][j];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) chol_diag_kernel(double* base, size_t stride, int ld, const int* na_arr,
                                                        int na_all, int j0, int* fail) {
    const int b = blockIdx.x;
    const int na = na_arr ? na_arr[b] : na_all;
    if (j0 >= na) return;
    __shared__ double D[kNB][kNB + 1];
    factor_diag_block(base + (size_t)b * stride, ld, j0, min(kNB, na - j0), &fail[b], D);
}

// Panel below an already factored diagonal block: row i <- row i * L_jj^-T.  grid (row tiles of 256, batch);
// rows are staged through shared memory so that global accesses are coalesced (256 contiguous bytes per row).
__global__ void __launch_bounds__(256) chol_panel_kernel(double* base, size_t stride, int ld, const int* na_arr,
                                                         int na_all, int j0) {
    const int b = blockIdx.y;
    const int na = na_arr ? na_arr[b] : na_all;
    if (j0 >= na) return;
    double* A = base + (size_t)b * stride;
    const int nb = min(kNB, na - j0);
    const int r0 = j0 + nb + blockIdx.x * 256;
    if (r0 >= na) return;
    __shared__ double D[kNB][kNB + 1];
    extern __shared__ double rows_sm[];  // 256 x (kNB + 1)
    const int tid = threadIdx.x;
    for (int e = tid; e < kNB * kNB; e += blockDim.x) {
        const int i = e / kNB, j = e % kNB;
        D[i][j] = (i < nb && j < nb && j <= i) ? A[(size_t)(j0 + i) * ld + j0 + j] : (i == j ? 1.0 : 0.0);
    }
    const int nrows = min(256, na - r0);
    for (int e = tid; e < nrows * kNB; e += blockDim.x) {
        const int r = e / kNB, c = e % kNB;
        rows_sm[r * (kNB + 1) + c] = (c < nb) ? A[(size_t)(r0 + r) * ld + j0 + c] : 0.0;
    }
    __syncthreads();
    if (tid < nrows) {
        double a[kNB];
        double* row = rows_sm + tid * (kNB + 1);
#pragma unroll
        for (int c = 0; c < kNB; ++c) a[c] = row[c];
#pragma unroll
        for (int c = 0; c < kNB; ++c) {
            double sacc = a[c];
#pragma unroll
            for (int t = 0; t < c; ++t) sacc -= a[t] * D[c][t];
            a[c] = sacc / D[c][c];
        }
#pragma unroll
        for (int c = 0; c < kNB; ++c) row[c] = a[c];
    }
    __syncthreads();
    for (int e = tid; e < nrows * kNB; e += blockDim.x) {
        const int r = e / kNB, c = e % kNB;
        if (c < nb) A[(size_t)(r0 + r) * ld + j0 + c] = rows_sm[r * (kNB + 1) + c];
    }
}

// trailing update C[i][j] -= sum_t P[i][t] P[j][t] for i >= j >= j0 + kNB, 64x64 tiles, lower tiles only
__global__ void __launch_bounds__(256) chol_update_kernel(double* base, size_t stride, int ld, const int* na_arr,
                                                          int na_all, int j0, int* fail) {
    const int b = blockIdx.z;
    const int na = na_arr ? na_arr[b] : na_all;
    const int t0 = j0 + kNB;
    if (t0 >= na) return;
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj > ti) return;
    const int r0 = t0 + ti * 64, c0 = t0 + tj * 64;
    if (r0 >= na) return;
    double* A = base + (size_t)b * stride;
    __shared__ double Pi[64][kNB + 1];
    __shared__ double Pj[64][kNB + 1];
    const int tid = threadIdx.x;
    for (int e = tid; e < 64 * kNB; e += blockDim.x) {
        const int r = e / kNB, c = e % kNB;
        Pi[r][c] = (r0 + r < na) ? A[(size_t)(r0 + r) * ld + j0 + c] : 0.0;
        Pj[r][c] = (c0 + r < na) ? A[(size_t)(c0 + r) * ld + j0 + c] : 0.0;
    }
    __syncthreads();
    const int tr = (tid / 16) * 4, tc = (tid % 16) * 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 8
    for (int t = 0; t < kNB; ++t) {
        double a[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = Pi[tr + i][t];
#pragma unroll
        for (int j = 0; j < 4; ++j) bb[j] = Pj[tc + j][t];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + tr + i, c = c0 + tc + j;
            if (r < na && c <= r) A[(size_t)r * ld + c] -= acc[i][j];
        }
    // the CTA that just finished the top-left trailing tile also factors the NEXT diagonal block, so the
    // next panel kernel finds it ready (saves one launch per block column)
    if (ti == 0 && tj == 0) {
        __threadfence_block();
        __syncthreads();
        double(*D)[kNB + 1] = reinterpret_cast<double(*)[kNB + 1]>(&Pi[0][0]);
        factor_diag_block(A, ld, t0, min(kNB, na - t0), &fail[b], D);
    }
}

static int chol_batched(ddb_ctx* ctx, double* base, size_t stride, int ld, const int* d_na, int na_max, int batch,
                        int* d_fail, int* launches) {
    const size_t panel_smem = 256 * (kNB + 1) * sizeof(double);
    DDB_CUDA_TRY(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem));
    chol_diag_kernel<<<batch, 256, 0, ctx->stream>>>(base, stride, ld, d_na, na_max, 0, d_fail);
    ++*launches;
    for (int j0 = 0; j0 < na_max; j0 += kNB) {
        const int rem = na_max - j0 - kNB;
        if (rem <= 0) break;
        chol_panel_kernel<<<dim3((rem + 255) / 256, batch), 256, panel_smem, ctx->stream>>>(base, stride, ld, d_na, na_max, j0);
        const int T = (rem + 63) / 64;
        chol_update_kernel<<<dim3(T, T, batch), 256, 0, ctx->stream>>>(base, stride, ld, d_na, na_max, j0, d_fail);
        *launches += 2;
    }
    DDB_CUDA_TRY(cudaGetLastError());
    return DDB_OK;
}

// ---------------------------------------------------------------------------------------------
// triangular solves with a row-major lower factor: z <- L^-T L^-1 z, one CTA per (rhs, matrix)
//   z vector b, r lives at zbase + (b * nrhs + r) * zstride
// ---------------------------------------------------------------------------------------------
// RPC right-hand sides per CTA (RPC = 1 for the per-target solves of the sweep; 4 when one matrix serves many
// right-hand sides -- inverse, DMD operator -- so that every element of L read from L2 is used four times).
template <int RPC>
__global__ void __launch_bounds__(256) chol_solve_kernel_t(const double* Lbase, size_t lstride, int ld,
                                                           const int* na_arr, int na_all, double* zbase, size_t zstride,
                                                           int nrhs, const int* active_matrix, const int* zmap) {
    const int b = blockIdx.y, r0 = blockIdx.x * RPC;
    if (active_matrix && !active_matrix[b]) return;
    const int na = na_arr ? na_arr[b] : na_all;
    if (na <= 0) return;
    const double* L = Lbase + (size_t)b * lstride;
    double* zg = zbase + ((size_t)(zmap ? zmap[b] : b) * nrhs + r0) * zstride;
    const int nr = min(RPC, nrhs - r0);
    extern __shared__ double sm[];
    const int zp = ((na + 31) / 32) * 32;
    double* z = sm;                   // RPC x zp
    double* blk = sm + RPC * zp;      // 32 x 33
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int q = 0; q < RPC; ++q)
        for (int i = tid; i < na; i += blockDim.x) z[q * zp + i] = (q < nr) ? zg[(size_t)q * zstride + i] : 0.0;
    __syncthreads();
    // forward: L y = z
    for (int j0 = 0; j0 < na; j0 += 32) {
        const int nb = min(32, na - j0);
        for (int e = tid; e < 32 * 32; e += blockDim.x) {
            const int i = e / 32, j = e % 32;
            blk[i * 33 + j] = (i < nb && j <= i) ? L[(size_t)(j0 + i) * ld + j0 + j] : (i == j ? 1.0 : 0.0);
        }
        __syncthreads();
        if (warp < RPC) {  // warp q substitutes right-hand side q
            double* zq = z + warp * zp;
            double zi = (lane < nb) ? zq[j0 + lane] : 0.0;
            for (int c = 0; c < nb; ++c) {
                const double zc = __shfl_sync(0xffffffffu, zi, c) / blk[c * 33 + c];
                if (lane == c) zi = zc;
                if (lane > c) zi -= blk[lane * 33 + c] * zc;
            }
            if (lane < nb) zq[j0 + lane] = zi;
        }
        __syncthreads();
        for (int i = j0 + nb + tid; i < na; i += blockDim.x) {
            const double* row = L + (size_t)i * ld + j0;
            double s[RPC];
#pragma unroll
            for (int q = 0; q < RPC; ++q) s[q] = 0.0;
            for (int c = 0; c < nb; ++c) {
                const double l = row[c];
#pragma unroll
                for (int q = 0; q < RPC; ++q) s[q] = fma(l, z[q * zp + j0 + c], s[q]);
            }
#pragma unroll
            for (int q = 0; q < RPC; ++q) z[q * zp + i] -= s[q];
        }
        __syncthreads();
    }
    // backward: L' x = y
    for (int j0 = ((na - 1) / 32) * 32; j0 >= 0; j0 -= 32) {
        const int nb = min(32, na - j0);
        for (int e = tid; e < 32 * 32; e += blockDim.x) {
            const int i = e / 32, j = e % 32;
            blk[i * 33 + j] = (i < nb && j <= i) ? L[(size_t)(j0 + i) * ld + j0 + j] : (i == j ? 1.0 : 0.0);
        }
        __syncthreads();
        if (warp < RPC) {
            double* zq = z + warp * zp;
            double zi = (lane < nb) ? zq[j0 + lane] : 0.0;
            for (int c = nb - 1; c >= 0; --c) {
                const double zc = __shfl_sync(0xffffffffu, zi, c) / blk[c * 33 + c];
                if (lane == c) zi = zc;
                if (lane < c) zi -= blk[c * 33 + lane] * zc;
            }
            if (lane < nb) zq[j0 + lane] = zi;
        }
        __syncthreads();
        for (int j = tid; j < j0; j += blockDim.x) {
            double s[RPC];
#pragma unroll
            for (int q = 0; q < RPC; ++q) s[q] = 0.0;
            for (int c = 0; c < nb; ++c) {
                const double l = L[(size_t)(j0 + c) * ld + j];
#pragma unroll
                for (int q = 0; q < RPC; ++q) s[q] = fma(l, z[q * zp + j0 + c], s[q]);
            }
#pragma unroll
            for (int q = 0; q < RPC; ++q) z[q * zp + j] -= s[q];
        }
        __syncthreads();
    }
    for (int q = 0; q < nr; ++q)
        for (int i = tid; i < na; i += blockDim.x) zg[(size_t)q * zstride + i] = z[q * zp + i];
}

__global__ void __launch_bounds__(256) chol_solve_kernel(const double* Lbase, size_t lstride, int ld,
                                                         const int* na_arr, int na_all, double* zbase, size_t zstride,
                                                         int nrhs, const int* active_matrix, const int* zmap) {
    const int b = blockIdx.y, r = blockIdx.x;
    if (active_matrix && !active_matrix[b]) return;
    const int na = na_arr ? na_arr[b] : na_all;
    if (na <= 0) return;
    const double* L = Lbase + (size_t)b * lstride;
    double* zg = zbase + ((size_t)(zmap ? zmap[b] : b) * nrhs + r) * zstride;
    extern __shared__ double sm[];
    double* z = sm;                 // na
    double* blk = sm + ((na + 31) / 32) * 32;  // 32 x 33
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < na; i += blockDim.x) z[i] = zg[i];
    __syncthreads();
    // forward: L y = z
    for (int j0 = 0; j0 < na; j0 += 32) {
        const int nb = min(32, na - j0);
        for (int e = tid; e < 32 * 32; e += blockDim.x) {
            const int i = e / 32, j = e % 32;
            blk[i * 33 + j] = (i < nb && j <= i) ? L[(size_t)(j0 + i) * ld + j0 + j] : (i == j ? 1.0 : 0.0);
        }
        __syncthreads();
        if (warp == 0) {
            double zi = (lane < nb) ? z[j0 + lane] : 0.0;
            for (int c = 0; c < nb; ++c) {
                const double zc = __shfl_sync(0xffffffffu, zi, c) / blk[c * 33 + c];
                if (lane == c) zi = zc;
                if (lane > c) zi -= blk[lane * 33 + c] * zc;
            }
            if (lane < nb) z[j0 + lane] = zi;
        }
        __syncthreads();
        for (int i = j0 + nb + tid; i < na; i += blockDim.x) {
            const double* row = L + (size_t)i * ld + j0;
            double s = 0.0;
            for (int c = 0; c < nb; ++c) s = fma(row[c], z[j0 + c], s);
            z[i] -= s;
        }
        __syncthreads();
    }
    // backward: L' x = y
    for (int j0 = ((na - 1) / 32) * 32; j0 >= 0; j0 -= 32) {
        const int nb = min(32, na - j0);
        for (int e = tid; e < 32 * 32; e += blockDim.x) {
            const int i = e / 32, j = e % 32;
            blk[i * 33 + j] = (i < nb && j <= i) ? L[(size_t)(j0 + i) * ld + j0 + j] : (i == j ? 1.0 : 0.0);
        }
        __syncthreads();
        if (warp == 0) {
            double zi = (lane < nb) ? z[j0 + lane] : 0.0;
            for (int c = nb - 1; c >= 0; --c) {
                const double zc = __shfl_sync(0xffffffffu, zi, c) / blk[c * 33 + c];
                if (lane == c) zi = zc;
                if (lane < c) zi -= blk[c * 33 + lane] * zc;
            }
            if (lane < nb) z[j0 + lane] = zi;
        }
        __syncthreads();
        for (int j = tid; j < j0; j += blockDim.x) {
            double s = 0.0;
            for (int c = 0; c < nb; ++c) s = fma(L[(size_t)(j0 + c) * ld + j], z[j0 + c], s);
            z[j] -= s;
        }
        __syncthreads();
    }
    for (int i = tid; i < na; i += blockDim.x) zg[i] = z[i];
}

// chol.cu / trsm.cu: blocked Cholesky with DMMA trailing updates and the GEMM-based substitution
int chol128_batched(ddb_ctx* ctx, double* base, size_t stride, int ld, const int* d_na, int na_max, int batch, int* d_fail,
                    int extra, int xrow, int* launches);
int chol_back_batched(ddb_ctx* ctx, double* base, size_t stride, int ld, const int* d_na, int na_max, int batch, int yrow0,
                      int nrhs, double* xout, size_t xstride_b, size_t xstride_r, const int* xmap, int* launches);
int chol_factor_shared(ddb_ctx* ctx, double* A, int k, int* d_fail, int extra, int* launches);
int chol_solve_many(ddb_ctx* ctx, const double* L, int k, double* Z, int nrhs, int* launches);
void comm_collect_time(ddb_ctx* ctx);  // comm.cu: duration of the last queued all-reduce -> timings (after a sync)
int comm_allreduce(ddb_ctx* ctx, double* d, int64_t count);  // comm.cu: in-place sum over the ranks on the compute stream
int eig_sym_device(ddb_ctx* ctx, double* dA, int n, double* dW);  // dmd_post.cu: symmetric eigen-decomposition (cuSOLVER Xsyevd)

// ---------------------------------------------------------------------------------------------
// per-target state and the shared post-solve step logic
// ---------------------------------------------------------------------------------------------
struct SweepDev {
    // problem
    int k, n_t, L, W, maxiters, algorithm, proximal;
    double abstol, reltol, rho, cad_rho;
    const double* lambdas;  // L, sorted
    const double* As;       // k x k equilibrated system matrix (symmetric)
    const double* dinv;     // k
    const double* rs;       // n_t x k scaled right-hand sides
    const double* rraw;     // n_t x k unscaled right-hand sides (R, column e = target e)
    // implicit mode (n_t == k, target e is library row e regressed on the others): Ninv = inverse of the
    // equilibrated system matrix; a solve with the full matrix and entry e of the right-hand side zeroed,
    // followed by z -= Ninv[:, e] * z[e] / Ninv[e, e], is the solve with row and column e deleted.
    int implicit;
    const int* init_masked; // device flag (or null): when set, the initial full least squares runs as a masked solve through
                            // the step kernels (the shared factorisation broke down: rank-deficient library / fewer
                            // samples than terms); decided on the device, the host does not wait for the flag
    const double* Ninv;     // k x k (symmetric) or null
    // per target
    double* x;              // n_t x k
    double* xprev;          // n_t x k
    double* aux1;           // n_t x k  (ADMM alpha / SR3 W)
    double* aux2;           // n_t x k  (ADMM w)
    double* znew;           // n_t x k  solve results (scaled unknowns) for the global-memory paths
    const double* zsol;     // n_t x k  where relax_finish reads the solved systems (znew, or the product with Ninv)
    uint32_t* active;       // n_t x W
    EqState* eq;            // n_t
    // candidates
    int32_t* cand_idx;
    double* cand_val;
    int64_t* cand_off;
    int32_t* cand_nnz;
    int32_t* cand_dof;
    int32_t* cand_eq;
    int64_t pool_cap;
    int cand_cap;
    int* counters;  // [0] ncand, [1] overflow flag, [2] pool used (low), [3] need_big count, [4] running count
    unsigned long long* pool_used;
    // runs (steps with an unchanged candidate are merged)
    int32_t* run_cand;
    int32_t* run_jfirst;  // threshold index of the run's first step
    int32_t* run_j;       // threshold index of the run's last step
    int run_cap;
    // path
    uint32_t* path_support;  // n_t x L x W
    double* path_coef;       // n_t x L x k
    int32_t* path_iters;     // n_t x L
};

__device__ __forceinline__ double threshold_cut(const SweepDev& S, double lam) {
    // value an entry must exceed (strictly) to stay active
    if (S.algorithm == DDB_ALG_STLSQ) return lam;
    if (S.algorithm == DDB_ALG_ADMM) return lam / S.rho;
    if (S.proximal == DDB_PROX_HARD) return sqrt(2.0 * lam);
    if (S.proximal == DDB_PROX_CAD) return isnan(S.cad_rho) ? 5.0 * lam : S.cad_rho;
    return lam;
}

// Append / merge the candidate described by the CURRENT x of target e (sparse: entries != 0) with
// dof = number of active entries.  Consecutive identical candidates extend the current run.
// Called by all threads of the CTA; `red`/`redi` are shared scratch of >= 32 entries.
__device__ void record_candidate(const SweepDev& S, int e, int j, int dof, double* red, int* redi, int* s_tmp,
                                 const double* x) {
    const int k = S.k;
    EqState& st = S.eq[e];
    // number of nonzeros
    int cnt = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) cnt += (x[i] != 0.0);
    const int nnz = block_sum_int(cnt, redi);
    // compare with the last candidate of this target
    int same = 0;
    const int last = st.last_cand;
    if (last >= 0 && S.cand_nnz[last] == nnz && S.cand_dof[last] == dof) {
        const int64_t off = S.cand_off[last];
        int diff = 0;
        for (int t = threadIdx.x; t < nnz; t += blockDim.x) {
            const int idx = S.cand_idx[off + t];
            const double v = S.cand_val[off + t];
            diff |= (__double_as_longlong(x[idx]) != __double_as_longlong(v));
        }
        same = (block_sum_int(diff, redi) == 0);
    }
    if (same) {
        if (threadIdx.x == 0) S.run_j[(size_t)e * S.run_cap + st.nruns - 1] = j;
        __syncthreads();
        return;
    }
    // new candidate: reserve id and pool space
    if (threadIdx.x == 0) {
        const int id = atomicAdd(&S.counters[0], 1);
        const unsigned long long off = atomicAdd(S.pool_used, (unsigned long long)nnz);
        int ok = 1;
        if (id >= S.cand_cap || off + (unsigned long long)nnz > (unsigned long long)S.pool_cap || st.nruns >= S.run_cap) {
            atomicExch(&S.counters[1], 1);
            ok = 0;
        }
        s_tmp[0] = ok ? id : -1;
        if (ok) {
            S.cand_off[id] = (int64_t)off;
            S.cand_nnz[id] = nnz;
            S.cand_dof[id] = dof;
            S.cand_eq[id] = e;
            S.run_cand[(size_t)e * S.run_cap + st.nruns] = id;
            S.run_jfirst[(size_t)e * S.run_cap + st.nruns] = j;
            S.run_j[(size_t)e * S.run_cap + st.nruns] = j;
            st.nruns += 1;
            st.last_cand = id;
        }
    }
    __syncthreads();
    const int id = s_tmp[0];
    if (id < 0) return;
    const int64_t off = S.cand_off[id];
    // ordered compaction: warp-level prefix over chunks of blockDim.x entries
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < k; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const int flag = (i < k) && (x[i] != 0.0);
        // block-wide exclusive scan of flag
        const unsigned ball = __ballot_sync(0xffffffffu, flag);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        if (lane == 0) redi[warp] = __popc(ball);
        __syncthreads();
        int wbase = 0, tot = 0;
        for (int w = 0; w < nw; ++w) {
            if (w < warp) wbase += redi[w];
            tot += redi[w];
        }
        const int pos = s_base + wbase + __popc(ball & ((1u << lane) - 1u));
        if (flag) {
            S.cand_idx[off + pos] = i;
            S.cand_val[off + pos] = x[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += tot;
        __syncthreads();
    }
}

// After the solve of one step wrote the new coefficients into x (entries of the old active set),
// finish the step exactly as `step!` + the body of the solver loop do.  xprev must already hold the
// pre-step x.  Returns nothing; updates the target's state machine.
// x / xp: the target's coefficient vectors -- in global memory (S.x + e k, S.xprev + e k) or the persistent kernel's
// shared-memory copies
__device__ void finish_step(const SweepDev& S, int e, double* red, int* redi, int* s_tmp, double* x, const double* xp) {
    const int k = S.k, W = S.W;
    EqState& st = S.eq[e];
    uint32_t* act = S.active + (size_t)e * W;
    const int j = st.j;
    const double lam = S.lambdas[j];
    const double cut = threshold_cut(S, lam);
    // active = |x| > cut ; x[!active] = 0   (_apply_active_set!, proximals.jl:1-13)
    // SR3 keeps x as produced by the proximal operator and only recomputes the mask (SR3.jl:138-139)
    int cnt = 0;
    double d2 = 0.0, n2 = 0.0;
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int i = w * 32 + b;
            if (i < k) {
                double v = x[i];
                const bool a = fabs(v) > cut;
                if (!a && S.algorithm != DDB_ALG_SR3) {
                    v = 0.0;
                    x[i] = 0.0;
                }
                bits |= (a ? 1u : 0u) << b;
                cnt += a;
                const double d = v - xp[i];
                d2 += d * d;
                n2 += v * v;
            }
        }
        act[w] = bits;
    }
    const int na = block_sum_int(cnt, redi);
    const double delta = sqrt(block_sum(d2, red));
    const double nx = sqrt(block_sum(n2, red));
    __syncthreads();
    record_candidate(S, e, j, na, red, redi, s_tmp, x);
    // convergence (_is_converged, DataDrivenSparse.jl:160-168)
    bool conv = (na == 0) || (delta < S.abstol) || (delta / nx < S.reltol);
    if (threadIdx.x == 0) {
        st.na = na;
        st.counter += 1;
        st.iter += 1;
    }
    __syncthreads();
    if (conv || st.iter >= S.maxiters) {
        // end of this threshold's inner loop: record the path entry, advance to the next threshold
        uint32_t* ps = S.path_support + ((size_t)e * S.L + j) * W;
        for (int w = threadIdx.x; w < W; w += blockDim.x) ps[w] = act[w];
        if (S.path_coef) {
            double* pc = S.path_coef + ((size_t)e * S.L + j) * k;
            for (int i = threadIdx.x; i < k; i += blockDim.x) pc[i] = x[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            S.path_iters[(size_t)e * S.L + j] = st.iter;
            st.j = j + 1;
            st.iter = 0;
            if (st.j >= S.L) st.status = kEqDone;
        }
    }
    __syncthreads();
}

// implicit mode: Ninv <- identity (right-hand sides of the inverse), z[e][e] <- 0
__global__ void implicit_prep_kernel(double* __restrict__ Ninv, double* __restrict__ z, int k) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)k * k) return;
    const int e = (int)(t / k), i = (int)(t % k);
    if (Ninv) Ninv[t] = (e == i) ? 1.0 : 0.0;
    if (z && e == i) z[t] = 0.0;
}

// implicit view: packed' = [G[idx, idx] | G[idx, idx] | diag] from the full packed blocks
__global__ void implicit_gather_kernel(const double* __restrict__ G, int k, const int32_t* __restrict__ idx, int kk,
                                       double* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)kk * kk) return;
    const int j = (int)(t / kk), i = (int)(t % kk);
    const double v = G[(size_t)idx[i] + (size_t)k * idx[j]];
    out[t] = v;
    out[(size_t)kk * kk + t] = v;
    if (i == j) out[2 * (size_t)kk * kk + i] = v;
}

int implicit_select(ddb_ctx* ctx, const int32_t* idx, int kk) {
    if (kk == 0) {
        ctx->imp_k = 0;
        return DDB_OK;
    }
    if (!ctx->d_packed.p || !ctx->have_gram || ctx->k <= 0) {
        set_error("ddb_implicit_select: no normal-equation blocks in the context (run ddb_gram first)");
        return DDB_BAD_STATE;
    }
    if (!idx || kk < 2 || kk > ctx->k) {
        set_error("ddb_implicit_select: need 2 <= kk <= k selected library rows");
        return DDB_INVALID_ARG;
    }
    std::vector<char> seen(ctx->k, 0);
    for (int i = 0; i < kk; ++i) {
        if (idx[i] < 0 || idx[i] >= ctx->k || seen[idx[i]]) {
            set_error("ddb_implicit_select: idx[%d] = %d is out of range or repeated", i, idx[i]);
            return DDB_INVALID_ARG;
        }
        seen[idx[i]] = 1;
    }
    cudaStream_t st = ctx->stream;
    DDB_TRY(ctx->imp_idx.reserve((size_t)kk * 4));
    DDB_TRY(ctx->imp_packed.reserve((2 * (size_t)kk * kk + kk) * 8));
    DDB_CUDA_TRY(cudaMemcpyAsync(ctx->imp_idx.p, idx, (size_t)kk * 4, cudaMemcpyHostToDevice, st));
    implicit_gather_kernel<<<(unsigned)(((size_t)kk * kk + 255) / 256), 256, 0, st>>>(
        ctx->d_packed.as<double>(), ctx->k, ctx->imp_idx.as<int32_t>(), kk, ctx->imp_packed.as<double>());
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaStreamSynchronize(st));  // idx is the caller's buffer
    ctx->imp_k = kk;
    if (ctx->sweep) ctx->sweep->valid = false;
    return DDB_OK;
}

// ---------------------------------------------------------------------------------------------
// init: x0 from the shared full solve, initial mask, zero-model bookkeeping
// ---------------------------------------------------------------------------------------------
// Tail of init_cache for target e, given the initial full solution in x: initial mask (strict cut at the
// smallest threshold), state machine, zero-model candidate.  Called by all threads of the CTA.
__device__ void finish_init(const SweepDev& S, int e, int* redi, const double* x) {
    const int k = S.k, W = S.W;
    EqState& st = S.eq[e];
    uint32_t* act = S.active + (size_t)e * W;
    const double cut = threshold_cut(S, S.lambdas[0]);  // lambdas sorted: [0] is the minimum
    __syncthreads();
    int cnt = 0;
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int i = w * 32 + b;
            if (i < k && fabs(x[i]) > cut) bits |= 1u << b, ++cnt;
        }
        act[w] = bits;
    }
    const int na = block_sum_int(cnt, redi);
    if (threadIdx.x == 0) {
        st.j = 0;
        st.iter = 0;
        st.counter = 0;
        st.na = na;
        st.status = kEqRunning;
        st.nruns = 0;
        st.last_cand = -1;
        st.dof0 = na;  // dof of the zeroed best cache = mask of the initial solution (solver.jl:82-83)
        st.pad[0] = 0;  // initial solve done
        st.pad[1] = 0;  // not parked for a minimum-norm big step
        // candidate e is the zero model of target e
        S.cand_off[e] = 0;
        S.cand_nnz[e] = 0;
        S.cand_dof[e] = na;
        S.cand_eq[e] = e;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) sweep_init_kernel(SweepDev S) {
    const int e = blockIdx.x;
    const int k = S.k, W = S.W;
    __shared__ int redi[32];
    EqState& st = S.eq[e];
    double* x = S.x + (size_t)e * k;
    double* xp = S.xprev + (size_t)e * k;
    const double* z = S.znew + (size_t)e * k;
    if ((S.implicit || (S.init_masked && *S.init_masked)) && S.algorithm == DDB_ALG_STLSQ) {
        // Implicit STLSQ: the initial least squares of target e runs on the library WITHOUT row e.  With an
        // exact implicit relation in the data the full Gram matrix is singular (that is the point of the
        // method) while G without a row of the relation is not, so there is no shared factorisation here: the
        // initial solve is a masked solve (mask = all but e) through the same kernels as every later step.
        uint32_t* act = S.active + (size_t)e * W;
        for (int i = threadIdx.x; i < k; i += blockDim.x) x[i] = xp[i] = 0.0;
        for (int w = threadIdx.x; w < W; w += blockDim.x) {
            uint32_t bits = (w * 32 + 32 <= k) ? 0xFFFFFFFFu : ((1u << (k - w * 32)) - 1u);
            if (S.implicit && e / 32 == w) bits &= ~(1u << (e % 32));
            act[w] = bits;
        }
        if (threadIdx.x == 0) {
            st.j = 0;
            st.iter = 0;
            st.counter = 0;
            st.na = S.implicit ? k - 1 : k;
            st.status = kEqRunning;
            st.nruns = 0;
            st.last_cand = -1;
            st.dof0 = 0;
            st.flags = 0;
            st.pad[0] = 1;  // initial solve pending
            st.pad[1] = 0;
        }
        return;
    }
    // implicit ADMM / SR3: row/column e deleted through the rank-one correction (see SweepDev)
    const double corr = S.implicit ? z[e] / S.Ninv[(size_t)e * k + e] : 0.0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        double v = z[i] * S.dinv[i];
        if (S.implicit) v = (i == e) ? 0.0 : (z[i] - S.Ninv[(size_t)e * k + i] * corr) * S.dinv[i];
        x[i] = v;
        xp[i] = (S.algorithm == DDB_ALG_SR3) ? v : 0.0;  // SR3.jl:122 copies, STLSQ/ADMM start from zero
        if (S.aux1) S.aux1[(size_t)e * k + i] = 0.0;
        if (S.aux2) S.aux2[(size_t)e * k + i] = 0.0;
    }
    if (threadIdx.x == 0) st.flags = 0;
    finish_init(S, e, redi, x);
}

// ---------------------------------------------------------------------------------------------
// Rank-deficient masked systems (fewer samples than active terms, collinear terms): minimum-norm least squares
// ---------------------------------------------------------------------------------------------
// The reference solves every STLSQ step with `A[p, :]' \ b` (STLSQ.jl:93-99, 135-140): pivoted QR, which returns the
// MINIMUM-NORM least-squares solution when the system is rank deficient (its own "Skinny" test: 6 terms, 5 samples,
// lib/DataDrivenSparse/test/Core/sparse_linear_solve.jl:63-96).  Through the normal equations that solution is
// x = G_pp^+ r_p with the UN-scaled Gram block (the minimum norm is not invariant under column scaling).  When the
// Cholesky factorisation of a small masked system breaks down, warp 0 diagonalises the block by cyclic Jacobi
// rotations in shared memory (na <= kPinvMax) and applies the pseudo-inverse; eigenvalues below na * 16 eps * max
// count as zero (singular values below ~1e-7 of the largest: what the squared condition number leaves of them).
constexpr int kPinvMax = 64;

// M: na x na symmetric (both triangles), V: na x na scratch, b: right-hand side -> solution.  Called by warp 0 only.
__device__ void pinv_solve_warp(double* M, int ldm, double* V, int ldv, int na, double* b) {
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < na * na; e += 32) V[(e / na) * ldv + e % na] = (e / na == e % na) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int e = lane; e < na * na; e += 32) {
            const int i = e / na, j = e % na;
            const double v = M[i * ldm + j];
            if (i == j) dg += v * v;
            else off += v * v;
        }
        for (int o = 16; o > 0; o >>= 1) {
            off += __shfl_xor_sync(0xffffffffu, off, o);
            dg += __shfl_xor_sync(0xffffffffu, dg, o);
        }
        if (off <= 1e-30 * dg) break;
        for (int p = 0; p < na - 1; ++p)
            for (int q = p + 1; q < na; ++q) {
                const double apq = M[p * ldm + q];
                const double app = M[p * ldm + p], aqq = M[q * ldm + q];
                if (fabs(apq) <= 1e-300 || fabs(apq) <= 1e-18 * sqrt(fabs(app * aqq))) continue;  // uniform over the warp
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                __syncwarp();
                for (int i = lane; i < na; i += 32) {  // columns p, q
                    const double mip = M[i * ldm + p], miq = M[i * ldm + q];
                    M[i * ldm + p] = c * mip - sn * miq;
                    M[i * ldm + q] = sn * mip + c * miq;
                    const double vip = V[i * ldv + p], viq = V[i * ldv + q];
                    V[i * ldv + p] = c * vip - sn * viq;
                    V[i * ldv + q] = sn * vip + c * viq;
                }
                __syncwarp();
                for (int j = lane; j < na; j += 32) {  // rows p, q
                    const double mpj = M[p * ldm + j], mqj = M[q * ldm + j];
                    M[p * ldm + j] = c * mpj - sn * mqj;
                    M[q * ldm + j] = sn * mpj + c * mqj;
                }
                __syncwarp();
            }
    }
    double lmax = 0.0;
    for (int i = 0; i < na; ++i) lmax = fmax(lmax, M[i * ldm + i]);
    const double tol = (double)na * 16.0 * 2.220446049250313e-16 * lmax;
    // x = sum_i [lambda_i > tol] (v_i' b / lambda_i) v_i ; lane = component
    double x0 = 0.0, x1 = 0.0;
    for (int i = 0; i < na; ++i) {
        const double lam = M[i * ldm + i];
        if (!(lam > tol)) continue;
        double d = 0.0;
        for (int j = lane; j < na; j += 32) d += V[j * ldv + i] * b[j];
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        d /= lam;
        if (lane < na) x0 += d * V[lane * ldv + i];
        if (lane + 32 < na) x1 += d * V[(lane + 32) * ldv + i];
    }
    __syncwarp();
    if (lane < na) b[lane] = x0;
    if (lane + 32 < na) b[lane + 32] = x1;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// STLSQ: persistent small-support kernel (one CTA per target)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stlsq_small_kernel(SweepDev S, int small_max, int x_in_smem) {
    const int e = blockIdx.x;
    const int k = S.k, W = S.W;
    extern __shared__ double smem[];
    double* M = smem;                                    // small_max x (small_max + 1)
    double* rhs = M + (size_t)small_max * (small_max + 1);  // small_max
    int* idx = reinterpret_cast<int*>(rhs + small_max);  // small_max
    __shared__ double red[32];
    __shared__ int redi[32];
    __shared__ int s_tmp[4];
    EqState& st = S.eq[e];
    double* xg = S.x + (size_t)e * k;
    double* xpg = S.xprev + (size_t)e * k;
    // The target's coefficient vectors live in shared memory for the whole run of the kernel when they fit (every step
    // makes several passes over all k entries: previous iterate, threshold / difference norms, candidate comparison --
    // from global memory each pass costs an L2 round trip, and a noise-free sweep is ~80 such steps in a row)
    double* x = x_in_smem ? reinterpret_cast<double*>(idx + small_max + (small_max & 1)) : xg;
    double* xp = x_in_smem ? x + k : xpg;
    if (x_in_smem) {
        for (int i = threadIdx.x; i < k; i += blockDim.x) x[i] = xg[i], xp[i] = xpg[i];
    }
    const uint32_t* act = S.active + (size_t)e * W;
    const double* rs = S.rs + (size_t)e * k;
    const int ldm = small_max + 1;

    while (true) {
        __syncthreads();
        if (st.status != kEqRunning) {
            if (x_in_smem)  // the blocked path and the selector read the iterate from global memory
                for (int i = threadIdx.x; i < k; i += blockDim.x) xg[i] = x[i], xpg[i] = xp[i];
            break;
        }
        const int na = st.na;
        if (na > small_max || st.pad[1]) {  // pad[1]: a rank-deficient system beyond the Jacobi solver (see below)
            if (threadIdx.x == 0) {
                st.status = kEqNeedBig;
                atomicAdd(&S.counters[3], 1);
            }
            if (x_in_smem)
                for (int i = threadIdx.x; i < k; i += blockDim.x) xg[i] = x[i], xpg[i] = xp[i];
            break;
        }
        // X_prev .= X   (STLSQ.jl:120)
        for (int i = threadIdx.x; i < k; i += blockDim.x) xp[i] = x[i];
        // ordered list of active indices
        if (threadIdx.x == 0) s_tmp[1] = 0;
        __syncthreads();
        for (int w0 = 0; w0 < W; w0 += blockDim.x) {
            const int w = w0 + threadIdx.x;
            const uint32_t bits = (w < W) ? act[w] : 0u;
            const int c = __popc(bits);
            // block exclusive scan of c
            int v = c;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            if (lane == 31) redi[warp] = v;
            __syncthreads();
            int wbase = 0, tot = 0;
            for (int q = 0; q < nw; ++q) {
                if (q < warp) wbase += redi[q];
                tot += redi[q];
            }
            int pos = s_tmp[1] + wbase + v - c;
            uint32_t bb = bits;
            while (bb) {
                const int b = __ffs(bb) - 1;
                bb &= bb - 1;
                idx[pos++] = w * 32 + b;
            }
            __syncthreads();
            if (threadIdx.x == 0) s_tmp[1] += tot;
            __syncthreads();
        }
        // gather the masked, equilibrated system into shared memory (lower triangle) and the rhs
        for (int t = threadIdx.x; t < na * na; t += blockDim.x) {
            const int i = t / na, j = t % na;
            if (j <= i) M[i * ldm + j] = S.As[(size_t)idx[i] * k + idx[j]];
        }
        for (int i = threadIdx.x; i < na; i += blockDim.x) rhs[i] = rs[idx[i]];
        __syncthreads();
        // Cholesky in shared memory (right-looking, column at a time)
        // The gathered block has a unit diagonal (equilibrated).  A pivot <= 0 is a breakdown; a pivot that is not
        // clearly positive is SUSPICIOUS: the block is either numerically rank deficient (fewer samples than terms,
        // collinear terms: the reference's pivoted QR then returns the minimum-norm solution) or merely very ill
        // conditioned (exact implicit relations: condition 1e13, yet full rank).  The two are told apart below by the
        // residual of the minimum-norm candidate.
        const double piv_tol = (double)na * 16.0 * 2.220446049250313e-16;
        int bad = 0, suspicious = 0;
        for (int c = 0; c < na; ++c) {
            const double d = M[c * ldm + c];
            if (!(d > 0.0)) bad = 1;
            if (!(d > piv_tol)) suspicious = 1;
            const double piv = (d > 0.0) ? sqrt(d) : 1.0;
            __syncthreads();
            if (threadIdx.x == 0) M[c * ldm + c] = piv;
            for (int i = c + 1 + threadIdx.x; i < na; i += blockDim.x) M[i * ldm + c] /= piv;
            __syncthreads();
            const int rem = na - c - 1;
            for (int t = threadIdx.x; t < rem * rem; t += blockDim.x) {
                const int i = c + 1 + t / rem, j = c + 1 + t % rem;
                if (j <= i) M[i * ldm + j] -= M[i * ldm + c] * M[j * ldm + c];
            }
            __syncthreads();
        }
        const bool try_pinv = suspicious && na <= kPinvMax && S.algorithm == DDB_ALG_STLSQ;
        if (suspicious && !try_pinv && S.algorithm == DDB_ALG_STLSQ) {
            // possibly rank deficient with more than kPinvMax active terms: nothing of the step has been published yet
            // (X_prev = X is repeated by the big path's gather) -- park the target; the next big step factorises it
            // again and, when that breaks down (pivot <= kPivotTol of chol.cu), takes the minimum-norm step in global
            // memory (pinv_big_*)
            __syncthreads();
            if (threadIdx.x == 0) st.pad[1] = 1;
            continue;
        }
        if (bad && !try_pinv && threadIdx.x == 0) st.flags |= 1;
        // forward / backward substitution by warp 0 (na <= small_max: serial over columns, lanes over rows)
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            for (int c = 0; c < na; ++c) {
                const double zc = rhs[c] / M[c * ldm + c];
                __syncwarp();
                if (lane == 0) rhs[c] = zc;
                for (int i = c + 1 + lane; i < na; i += 32) rhs[i] -= M[i * ldm + c] * zc;
                __syncwarp();
            }
            for (int c = na - 1; c >= 0; --c) {
                const double zc = rhs[c] / M[c * ldm + c];
                __syncwarp();
                if (lane == 0) rhs[c] = zc;
                for (int i = lane; i < c; i += 32) rhs[i] -= M[c * ldm + i] * zc;
                __syncwarp();
            }
        }
        __syncthreads();
        if (try_pinv) {
            // minimum-norm candidate x_p = G_pp^+ r_p on the UN-scaled block; it replaces the Cholesky solution when
            // the factorisation broke down or when x_p solves the normal equations (residual ~ rounding): the block
            // is then rank deficient and x_p is what `A[p, :]' \ b` returns in the reference
            const double xc = (threadIdx.x < na) ? rhs[threadIdx.x] * S.dinv[idx[threadIdx.x]] : 0.0;
            __syncthreads();
            for (int t = threadIdx.x; t < na * na; t += blockDim.x) {
                const int i = t / na, j = t % na;
                const double di = S.dinv[idx[i]], dj = S.dinv[idx[j]];
                M[i * ldm + j] = (di > 0.0 && dj > 0.0) ? S.As[(size_t)idx[i] * k + idx[j]] / (di * dj) : 0.0;
            }
            for (int i = threadIdx.x; i < na; i += blockDim.x) {
                const double di = S.dinv[idx[i]];
                rhs[i] = di > 0.0 ? rs[idx[i]] / di : 0.0;
            }
            __syncthreads();
            if (threadIdx.x < 32) pinv_solve_warp(M, ldm, M + (size_t)kPinvMax * ldm, ldm, na, rhs);
            __syncthreads();
            double ri = 0.0, bi = 0.0;
            if (threadIdx.x < na) {
                const int i = threadIdx.x;
                const double di = S.dinv[idx[i]];
                bi = di > 0.0 ? rs[idx[i]] / di : 0.0;
                double acc = -bi;
                for (int j = 0; j < na; ++j) {
                    const double dj = S.dinv[idx[j]];
                    if (di > 0.0 && dj > 0.0) acc = fma(S.As[(size_t)idx[i] * k + idx[j]] / (di * dj), rhs[j], acc);
                }
                ri = acc;
            }
            const double res2 = block_sum(ri * ri, red);
            const double b2 = block_sum(bi * bi, red);
            const bool use_p = bad || res2 <= 1e-14 * b2;
            if (threadIdx.x < na) x[idx[threadIdx.x]] = use_p ? rhs[threadIdx.x] : xc;
            if (use_p && threadIdx.x == 0) st.flags |= 2;  // bit 1: a minimum-norm solution was used
            __syncthreads();
        } else {
            // X[p] = solution (un-scaled)   (STLSQ.jl:128-140)
            for (int i = threadIdx.x; i < na; i += blockDim.x) x[idx[i]] = rhs[i] * S.dinv[idx[i]];
            __syncthreads();
        }
        if (st.pad[0])
            finish_init(S, e, redi, x);  // that was the initial solve of an implicit target, not a step
        else
            finish_step(S, e, red, redi, s_tmp, x, xp);
    }
}

// big path, STLSQ: gather masked systems of the flagged targets into global workspaces
// grid (row splits, number of flagged targets); dynamic shared memory: k ints
// Sharded runs (nr ranks, this is rank rk): every rank rebuilds the index lists of ALL parked targets (the finish step
// needs them everywhere) but gathers -- and later factorises -- only the systems of slots rk, rk + nr, ...; local slot
// slot / nr of the workspace, its order / global slot recorded in na_loc / map_loc.
__global__ void __launch_bounds__(256) stlsq_big_gather_kernel(SweepDev S, const int* big_list, double* ws,
                                                               size_t ws_stride, int* idx_all, int* na_out, int rhs_row,
                                                               int nr, int rk, int* na_loc, int* map_loc) {
    const int slot = blockIdx.y;
    const bool mine = (slot % nr) == rk;
    if (!mine && blockIdx.x != 0) return;
    const int e = big_list[slot];
    const int k = S.k, W = S.W;
    const uint32_t* act = S.active + (size_t)e * W;
    extern __shared__ int sidx[];
    __shared__ int s_na;
    if (threadIdx.x == 0) {  // ordered list of active indices (every CTA of the slot rebuilds it)
        int pos = 0;
        for (int w = 0; w < W; ++w) {
            uint32_t bb = act[w];
            while (bb) {
                const int b = __ffs(bb) - 1;
                bb &= bb - 1;
                sidx[pos++] = w * 32 + b;
            }
        }
        s_na = pos;
    }
    __syncthreads();
    const int na = s_na;
    double* A = ws + (size_t)(slot / nr) * ws_stride;
    if (mine)
        for (int i = blockIdx.x; i < na; i += gridDim.x) {
            const double* src = S.As + (size_t)sidx[i] * k;
            for (int j = threadIdx.x; j <= i; j += blockDim.x) A[(size_t)i * k + j] = src[sidx[j]];
        }
    if (blockIdx.x == 0) {
        int* idx = idx_all + (size_t)slot * k;
        double* x = S.x + (size_t)e * k;
        double* xp = S.xprev + (size_t)e * k;
        double* z = S.znew + (size_t)e * k;
        const double* rs = S.rs + (size_t)e * k;
        if (threadIdx.x == 0) {
            na_out[slot] = na;
            if (mine && na_loc) na_loc[slot / nr] = na, map_loc[slot / nr] = slot;
        }
        for (int i = threadIdx.x; i < na; i += blockDim.x) idx[i] = sidx[i];
        for (int i = threadIdx.x; i < k; i += blockDim.x) xp[i] = x[i];  // X_prev .= X
        // rhs_row: the right-hand side goes to row k of the slot's matrix buffer (it rides through the blocked
        // factorisation); otherwise to znew (solved in place by chol_solve_kernel)
        if (mine) {
            double* zdst = rhs_row ? A + (size_t)k * k : z;
            for (int i = threadIdx.x; i < na; i += blockDim.x) zdst[i] = rs[sidx[i]];
        }
    }
}

// sharded big steps: failure flags of my slots into the exchange buffer / solutions and flags of ALL slots out of it
__global__ void big_fail_pack_kernel(const int* fail_loc, const int* map_loc, int nloc, double* xbuf, int k) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nloc) xbuf[(size_t)map_loc[b] * (k + 1) + k] = fail_loc[b] ? 1.0 : 0.0;
}
__global__ void __launch_bounds__(256) big_unpack_kernel(const double* xbuf, const int* big_list, const int* na_arr, int k,
                                                         double* znew, int* fail) {
    const int slot = blockIdx.x;
    const double* src = xbuf + (size_t)slot * (k + 1);
    double* z = znew + (size_t)big_list[slot] * k;
    for (int i = threadIdx.x; i < na_arr[slot]; i += blockDim.x) z[i] = src[i];
    if (threadIdx.x == 0) fail[slot] = src[k] != 0.0;
}

__global__ void __launch_bounds__(256) stlsq_big_finish_kernel(SweepDev S, const int* big_list, const int* idx_all,
                                                               const int* na_arr, const int* fail) {
    const int slot = blockIdx.x;
    const int e = big_list[slot];
    const int k = S.k;
    __shared__ double red[32];
    __shared__ int redi[32];
    __shared__ int s_tmp[4];
    const int* idx = idx_all + (size_t)slot * k;
    const int na = na_arr[slot];
    double* x = S.x + (size_t)e * k;
    const double* z = S.znew + (size_t)e * k;
    for (int i = threadIdx.x; i < na; i += blockDim.x) x[idx[i]] = z[i] * S.dinv[idx[i]];
    if (threadIdx.x == 0) {
        if (fail[slot]) S.eq[e].flags |= 1;
        S.eq[e].status = kEqRunning;
        S.eq[e].pad[1] = 0;
    }
    __syncthreads();
    if (S.eq[e].pad[0])
        finish_init(S, e, redi, x);  // initial solve of an implicit target
    else
        finish_step(S, e, red, redi, s_tmp, x, S.xprev + (size_t)e * k);
}

// ---------------------------------------------------------------------------------------------
// Minimum-norm big steps: rank-deficient masked systems of more than kPinvMax active terms
// ---------------------------------------------------------------------------------------------
// Same rule as pinv_solve_warp (x_p = G_pp^+ r_p on the UN-scaled block, eigenvalues below na * 16 eps * max count
// as zero), for systems that live in global memory: the block is gathered un-scaled, diagonalised by the symmetric
// eigen-solver behind eig_sym_device, and the pseudo-inverse applied by two small kernels.  Only reached when the
// batched Cholesky factorisation of a parked target broke down (fewer samples than active terms, collinear terms),
// i.e. where the reference's `A[p, :]' \ b` (STLSQ.jl:93-99, 135-140) returns its minimum-norm solution.
__global__ void __launch_bounds__(256) pinv_big_gather_kernel(SweepDev S, int e, const int* __restrict__ idx, int na,
                                                              double* __restrict__ G, double* __restrict__ b) {
    const int k = S.k;
    const int i = blockIdx.x;
    const double di = S.dinv[idx[i]];
    const double* src = S.As + (size_t)idx[i] * k;
    for (int j = threadIdx.x; j < na; j += blockDim.x) {
        const double dj = S.dinv[idx[j]];
        G[(size_t)i * na + j] = (di > 0.0 && dj > 0.0) ? src[idx[j]] / (di * dj) : 0.0;
    }
    if (threadIdx.x == 0) b[i] = di > 0.0 ? S.rs[(size_t)e * k + idx[i]] / di : 0.0;
}

// c_i = [lambda_i > tol] (v_i' b) / lambda_i ; one warp per eigenvector (column i of V: V[j + na * i])
__global__ void __launch_bounds__(256) pinv_big_project_kernel(const double* __restrict__ V, const double* __restrict__ lam,
                                                               int na, const double* __restrict__ b, double* __restrict__ c) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= na) return;
    const double tol = (double)na * 16.0 * 2.220446049250313e-16 * fmax(lam[na - 1], 0.0);  // ascending eigenvalues
    const double li = lam[i];
    double d = 0.0;
    if (li > tol) {
        const double* v = V + (size_t)i * na;
        for (int j = lane; j < na; j += 32) d = fma(v[j], b[j], d);
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        d /= li;
    }
    if (lane == 0) c[i] = d;
}

// x = V c, handed to the finish step in the scaled form it expects (x_j = z_j dinv_j); the slot's failure flag is
// cleared and bit 1 of the target's flags records that a minimum-norm solution was used
__global__ void __launch_bounds__(256) pinv_big_expand_kernel(SweepDev S, int e, int slot, const int* __restrict__ idx, int na,
                                                              const double* __restrict__ V, const double* __restrict__ c,
                                                              int* fail) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < na) {
        double x = 0.0;
        for (int i = 0; i < na; ++i) x = fma(V[(size_t)i * na + j], c[i], x);
        const double dj = S.dinv[idx[j]];
        S.znew[(size_t)e * S.k + j] = dj > 0.0 ? x / dj : 0.0;
    }
    if (j == 0) {
        fail[slot] = 0;
        S.eq[e].flags |= 2;
    }
}

// ---------------------------------------------------------------------------------------------
// ADMM / SR3: one iteration = shared-factor solve for every running target + fused prox step
// ---------------------------------------------------------------------------------------------
// rhs (scaled) for the next solve:  ADMM  b = r + rho (alpha - w);  SR3  b = r + nu x
__global__ void __launch_bounds__(256) relax_rhs_kernel(SweepDev S, int* active_matrix) {
    const int e = blockIdx.x;
    const int k = S.k;
    EqState& st = S.eq[e];
    if (threadIdx.x == 0) active_matrix[e] = (st.status == kEqRunning);
    if (st.status != kEqRunning) return;
    double* x = S.x + (size_t)e * k;
    double* xp = S.xprev + (size_t)e * k;
    double* z = S.znew + (size_t)e * k;
    const double* r = S.rraw + (size_t)e * k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const double xi = x[i];
        xp[i] = xi;  // X_prev .= X
        double b;
        if (S.algorithm == DDB_ALG_ADMM)
            b = r[i] + S.rho * (S.aux1[(size_t)e * k + i] - S.aux2[(size_t)e * k + i]);
        else
            b = r[i] + xi * S.rho;
        z[i] = (S.implicit && i == e) ? 0.0 : b * S.dinv[i];
    }
}

// ADMM / SR3 with a large library: every iteration solves the SAME matrix for all running targets.  One CTA
// per right-hand side streaming the whole factor (chol_solve_kernel) is bound by the L2 rate of a single SM
// (37 MB per solve at k = 2145); with the inverse of the equilibrated matrix formed once per sweep the
// iteration is a product  Zsol (n_t x k) = Z * Ninv  that reads Ninv once for all targets.
// CTA tile: 64 targets x 16 columns, K in chunks of 32 through shared memory; grid (ceil(k/16), ceil(n_t/64)).
__global__ void __launch_bounds__(256) ninv_apply_kernel(const double* __restrict__ Ninv, int k, const double* __restrict__ Z,
                                                         int n_t, const int* __restrict__ active, double* __restrict__ Zsol) {
    __shared__ double zs[64][33];
    __shared__ double ns[32][17];
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * 16, e0 = blockIdx.y * 64;
    const int te = tid >> 2, tc = (tid & 3) * 4;  // this thread: target e0 + te, columns c0 + tc .. + 3
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k0 = 0; k0 < k; k0 += 32) {
        for (int q = tid; q < 64 * 32; q += 256) {
            const int e = q >> 5, kk = q & 31;
            zs[e][kk] = (e0 + e < n_t && k0 + kk < k) ? Z[(size_t)(e0 + e) * k + k0 + kk] : 0.0;
        }
        for (int q = tid; q < 32 * 16; q += 256) {
            const int kk = q >> 4, c = q & 15;
            ns[kk][c] = (k0 + kk < k && c0 + c < k) ? Ninv[(size_t)(k0 + kk) * k + c0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            const double z = zs[te][kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fma(z, ns[kk][tc + j], acc[j]);
        }
        __syncthreads();
    }
    const int e = e0 + te;
    if (e < n_t && active[e]) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c0 + tc + j < k) Zsol[(size_t)e * k + c0 + tc + j] = acc[j];
    }
}

__device__ __forceinline__ double soft(double v, double lam) {
    const double a = fabs(v) - lam;
    const double sgn = (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : 0.0);
    return sgn * (a > 0.0 ? a : 0.0);
}

__global__ void __launch_bounds__(256) relax_finish_kernel(SweepDev S) {
    const int e = blockIdx.x;
    const int k = S.k;
    __shared__ double red[32];
    __shared__ int redi[32];
    __shared__ int s_tmp[4];
    EqState& st = S.eq[e];
    if (st.status != kEqRunning) return;
    double* x = S.x + (size_t)e * k;
    const double* z = S.zsol + (size_t)e * k;
    const double lam = S.lambdas[st.j];
    // implicit mode: row/column e deleted from the system through the rank-one correction (see SweepDev)
    const double corr = S.implicit ? z[e] / S.Ninv[(size_t)e * k + e] : 0.0;
    auto solved = [&](int i) -> double {
        if (!S.implicit) return z[i] * S.dinv[i];
        return (i == e) ? 0.0 : (z[i] - S.Ninv[(size_t)e * k + i] * corr) * S.dinv[i];
    };
    if (S.algorithm == DDB_ALG_ADMM) {
        double* alpha = S.aux1 + (size_t)e * k;
        double* w = S.aux2 + (size_t)e * k;
        const double t = lam / S.rho;
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            const double xi = solved(i);              // X = (B + rho (alpha - w)) / A
            const double a = soft(xi + w[i], t);      // alpha = prox(X + w, lambda/rho)
            alpha[i] = a;
            w[i] += xi - a;                           // w += X - alpha
            x[i] = xi;                                // thresholded by finish_step
        }
    } else {
        double* Wv = S.aux1 + (size_t)e * k;
        const double cadr = isnan(S.cad_rho) ? 5.0 * lam : S.cad_rho;
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            const double wi = solved(i);              // W = (B + nu X) / A
            Wv[i] = wi;
            double xi;
            if (S.proximal == DDB_PROX_HARD) xi = (fabs(wi) > sqrt(2.0 * lam)) ? wi : 0.0;
            else if (S.proximal == DDB_PROX_SOFT) xi = soft(wi, lam);
            else xi = (fabs(wi) > cadr) ? wi : soft(wi, lam);
            x[i] = xi;                                // X = prox(W, lambda)
        }
    }
    __syncthreads();
    finish_step(S, e, red, redi, s_tmp, x, S.xprev + (size_t)e * k);
}

__global__ void count_status_kernel(const EqState* eq, int n_t, int* counters) {
    int running = 0, big = 0;
    for (int e = threadIdx.x; e < n_t; e += blockDim.x) {
        running += (eq[e].status == kEqRunning);
        big += (eq[e].status == kEqNeedBig);
    }
    for (int o = 16; o > 0; o >>= 1) {
        running += __shfl_xor_sync(0xffffffffu, running, o);
        big += __shfl_xor_sync(0xffffffffu, big, o);
    }
    if (threadIdx.x == 0) {
        counters[4] = running;
        counters[3] = big;
    }
}

__global__ void list_big_kernel(const EqState* eq, int n_t, int* big_list, int* nbig) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int c = 0;
        for (int e = 0; e < n_t; ++e)
            if (eq[e].status == kEqNeedBig) big_list[c++] = e;
        *nbig = c;
    }
}

// ---------------------------------------------------------------------------------------------
// canonical candidate numbering
// ---------------------------------------------------------------------------------------------
// record_candidate hands out ids with a global atomicAdd from CTAs that run concurrently (one per target), so
// how the ids of different targets interleave depends on timing -- and differs between the GPUs of a sharded
// run, whose candidate-rss tables are summed entry by entry.  After the sweep the candidates are therefore
// renumbered into an order that only depends on the sweep's (deterministic) content: the zero models 0..n_t-1,
// then the candidates of target 0 in the order it recorded them (one per run), target 1, ...  The sparse
// records stay where they are in the pool (offsets are private to a rank); only the tables indexed by id move.
__global__ void __launch_bounds__(1024) canon_base_kernel(const EqState* __restrict__ eq, int n_t, int* __restrict__ base) {
    __shared__ int s_part[1024];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = n_t;
    __syncthreads();
    for (int e0 = 0; e0 < n_t; e0 += blockDim.x) {
        const int e = e0 + threadIdx.x;
        const int c = e < n_t ? eq[e].nruns : 0;
        s_part[threadIdx.x] = c;
        __syncthreads();
        for (int o = 1; o < (int)blockDim.x; o <<= 1) {  // inclusive Hillis-Steele scan
            const int t = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        if (e < n_t) base[e] = s_carry + s_part[threadIdx.x] - c;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry += s_part[threadIdx.x];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) canon_permute_kernel(SweepDev S, const int* __restrict__ base, int64_t* __restrict__ t_off,
                                                            int32_t* __restrict__ t_nnz, int32_t* __restrict__ t_dof,
                                                            int32_t* __restrict__ t_eq) {
    const int e = blockIdx.x;
    EqState& st = S.eq[e];
    const int nruns = st.nruns, b = base[e];
    for (int r = threadIdx.x; r < nruns; r += blockDim.x) {
        const int old = S.run_cand[(size_t)e * S.run_cap + r], nw = b + r;
        t_off[nw] = S.cand_off[old];
        t_nnz[nw] = S.cand_nnz[old];
        t_dof[nw] = S.cand_dof[old];
        t_eq[nw] = e;
        S.run_cand[(size_t)e * S.run_cap + r] = nw;
    }
    __syncthreads();
    if (threadIdx.x == 0 && nruns > 0) st.last_cand = b + nruns - 1;
}

// ---------------------------------------------------------------------------------------------
// selection along the recorded step sequence (solver.jl:94-115)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double score_of(int selector, double rss, int dof, double nobs) {
    if (selector == DDB_SEL_RSS) return rss;
    const double m2ll = nobs * log(rss / nobs);  // -2 loglikelihood  (DataDrivenSparse.jl:188-192)
    if (selector == DDB_SEL_BIC) return m2ll + dof * log(nobs);
    const double aic = m2ll + 2.0 * dof;
    if (selector == DDB_SEL_AIC) return aic;
    return aic + (2.0 * dof * (dof + 1.0)) / (nobs - dof - 1.0);
}

// rss[c] = streamed value of candidate c's group representative (summed over ranks on the sharded path);
// corr[c] = the exact quadratic correction of rss.cu (zero for candidates that were streamed themselves)
__global__ void __launch_bounds__(256) select_kernel(SweepDev S, const double* __restrict__ rss_rep,
                                                     const double* __restrict__ corr, int selector, double nobs, double* coef /* n_t x k col-major */,
                                                     double* lambda_opt, int32_t* iters, double* rss_out,
                                                     int32_t* dof_out, int32_t* flags) {
    const int e = blockIdx.x;
    const int k = S.k;
    __shared__ int s_best;
    __shared__ int s_bestj;
    const EqState& st = S.eq[e];
    auto rss = [&](int c) -> double {
        const double v = rss_rep[c] + corr[c];
        return v > 0.0 ? v : 0.0;
    };
    if (threadIdx.x == 0) {
        int best = e;  // zero model
        int bestj = 0;
        double best_score = score_of(selector, rss(e), S.cand_dof[e], nobs);
        for (int r = 0; r < st.nruns; ++r) {
            const int c = S.run_cand[(size_t)e * S.run_cap + r];
            const int jfirst = S.run_jfirst[(size_t)e * S.run_cap + r];
            const int jlast = S.run_j[(size_t)e * S.run_cap + r];
            const double sc = score_of(selector, rss(c), S.cand_dof[c], nobs);
            // A run is a maximal sequence of steps whose cache content did not change.  Its first step is
            // taken unconditionally when it belongs to the first threshold (`j == 1` in the reference),
            // otherwise iff score <= best; once taken, the remaining steps of the run tie with themselves
            // and `<=` moves the optimal threshold to the run's last one.
            if (jfirst == 0 || sc <= best_score) {
                best = c;
                bestj = jlast;
                best_score = sc;
            }
        }
        s_best = best;
        s_bestj = bestj;
        if (lambda_opt) lambda_opt[e] = S.lambdas[bestj < S.L ? bestj : S.L - 1];
        if (iters) iters[e] = st.counter;
        if (rss_out) rss_out[e] = rss(best);
        if (dof_out) dof_out[e] = S.cand_dof[best];
        if (flags) flags[e] = st.flags;
    }
    __syncthreads();
    if (coef) {
        for (int i = threadIdx.x; i < k; i += blockDim.x) coef[(size_t)e + (size_t)S.n_t * i] = 0.0;
        __syncthreads();
        const int c = s_best;
        const int64_t off = S.cand_off[c];
        for (int t = threadIdx.x; t < S.cand_nnz[c]; t += blockDim.x)
            coef[(size_t)e + (size_t)S.n_t * S.cand_idx[off + t]] = S.cand_val[off + t];
    }
}

}  // namespace ddb

// =============================================================================================
// Host side
// =============================================================================================
namespace ddb {

void SweepState::release_all() {
    DevBuf* all[] = {&As, &Lfac, &dinv, &rs, &x, &xprev, &aux1, &aux2, &znew, &active, &eq, &lambdas_d, &cand_idx,
                     &cand_val, &cand_off, &cand_nnz, &cand_dof, &cand_eq, &counters, &pool_used_d, &run_cand,
                     &run_jfirst, &run_j, &path_support, &path_coef, &path_iters, &big_ws, &big_idx, &big_na,
                     &big_list, &big_fail, &big_loc, &big_xbuf, &dense_ids, &dense_work, &active_matrix, &shared_fail, &rss, &rss_partials, &out_tmp, &cand_terms,
                     &Ninv, &tgt_terms, &rss_corr, &grp_buf, &rep_buf, &zsol, &canon_tmp, &pinv_ws};
    for (DevBuf* b : all) b->release();
}

void sweep_free(ddb_ctx* ctx) {
    if (ctx->sweep) {
        ctx->sweep->release_all();
        delete ctx->sweep;
        ctx->sweep = nullptr;
    }
}

static SweepDev make_dev(ddb_ctx* ctx, SweepState* S) {
    SweepDev D;
    memset(&D, 0, sizeof(D));
    D.k = S->k;
    D.n_t = S->n_t;
    D.L = S->L;
    D.W = S->W;
    D.maxiters = S->prm.maxiters;
    D.algorithm = S->prm.algorithm;
    D.proximal = S->prm.proximal;
    D.abstol = S->prm.abstol;
    D.reltol = S->prm.reltol;
    D.rho = S->prm.rho;
    D.cad_rho = S->prm.cad_rho;
    D.lambdas = S->lambdas_d.as<double>();
    D.As = S->As.as<double>();
    D.dinv = S->dinv.as<double>();
    D.rs = S->rs.as<double>();
    D.rraw = (ctx->imp_k > 0 ? ctx->imp_packed.as<double>() : ctx->d_packed.as<double>()) + (size_t)S->k * S->k;
    D.implicit = ctx->imp_k > 0;
    D.Ninv = D.implicit ? S->Ninv.as<double>() : nullptr;
    D.x = S->x.as<double>();
    D.xprev = S->xprev.as<double>();
    D.aux1 = S->prm.algorithm == DDB_ALG_STLSQ ? nullptr : S->aux1.as<double>();
    D.aux2 = S->prm.algorithm == DDB_ALG_ADMM ? S->aux2.as<double>() : nullptr;
    D.znew = S->znew.as<double>();
    D.zsol = D.znew;
    D.active = S->active.as<uint32_t>();
    D.eq = S->eq.as<EqState>();
    D.cand_idx = S->cand_idx.as<int32_t>();
    D.cand_val = S->cand_val.as<double>();
    D.cand_off = S->cand_off.as<int64_t>();
    D.cand_nnz = S->cand_nnz.as<int32_t>();
    D.cand_dof = S->cand_dof.as<int32_t>();
    D.cand_eq = S->cand_eq.as<int32_t>();
    D.pool_cap = S->pool_cap;
    D.cand_cap = S->cand_cap;
    D.counters = S->counters.as<int>();
    D.pool_used = S->pool_used_d.as<unsigned long long>();
    D.run_cand = S->run_cand.as<int32_t>();
    D.run_jfirst = S->run_jfirst.as<int32_t>();
    D.run_j = S->run_j.as<int32_t>();
    D.run_cap = S->run_cap;
    D.path_support = S->path_support.as<uint32_t>();
    D.path_coef = S->path_coef.as<double>();
    D.path_iters = S->path_iters.as<int32_t>();
    return D;
}

constexpr int kSmallMax = 128;

int sweep_run(ddb_ctx* ctx, const ddb_sweep_params* prm, const double* lambdas, int L, int64_t m_total) {
    if (!prm || !lambdas || L <= 0 || m_total <= 0) {
        set_error("ddb_sweep: params, lambdas must be non-null and L, m_total positive");
        return DDB_INVALID_ARG;
    }
    if (!ctx->d_packed.p || ctx->k <= 0 || (ctx->n_t <= 0 && ctx->imp_k <= 0)) {
        set_error("ddb_sweep: no normal-equation blocks in the context (run ddb_gram or ddb_gram_upload first)");
        return DDB_BAD_STATE;
    }
    if (prm->algorithm < DDB_ALG_STLSQ || prm->algorithm > DDB_ALG_SR3 || prm->maxiters <= 0 || !(prm->rho >= 0.0)) {
        set_error("ddb_sweep: invalid algorithm / maxiters / rho");
        return DDB_INVALID_ARG;
    }
    if (prm->algorithm != DDB_ALG_STLSQ && !(prm->rho > 0.0)) {
        set_error("ddb_sweep: ADMM rho / SR3 nu must be positive");
        return DDB_INVALID_ARG;
    }
    for (int j = 0; j < L; ++j)
        if (!(lambdas[j] > 0.0)) {
            set_error("ddb_sweep: thresholds must be positive (STLSQ.jl:56)");
            return DDB_INVALID_ARG;
        }
    if (!ctx->sweep) ctx->sweep = new SweepState();
    SweepState* S = ctx->sweep;
    S->valid = false;
    S->have_rss = false;
    const bool implicit = ctx->imp_k > 0;
    const int k = implicit ? ctx->imp_k : ctx->k, n_t = implicit ? ctx->imp_k : ctx->n_t;
    const int W = (k + 31) / 32;
    S->k = k;
    S->n_t = n_t;
    S->L = L;
    S->W = W;
    S->m_total = m_total;
    S->prm = *prm;
    S->lambdas.assign(lambdas, lambdas + L);
    std::sort(S->lambdas.begin(), S->lambdas.end());  // solver.jl:77-79
    const int64_t steps_bound = (int64_t)L * prm->maxiters;
    S->run_cap = (int)std::min<int64_t>(steps_bound, 1 << 16) + 1;
    S->cand_cap = n_t + n_t * (int)std::min<int64_t>(steps_bound, 8192);
    S->pool_cap = std::max<int64_t>((int64_t)8 << 20, (int64_t)4 * n_t * k);

    const size_t kk = (size_t)k * k, nk = (size_t)n_t * k;
    DDB_TRY(S->As.reserve(kk * 8));
    DDB_TRY(S->Lfac.reserve((kk + nk) * 8));  // + n_t rows: the right-hand sides ride through the factorisation
    DDB_TRY(S->dinv.reserve((size_t)k * 8));
    DDB_TRY(S->rs.reserve(nk * 8));
    DDB_TRY(S->x.reserve(nk * 8));
    DDB_TRY(S->xprev.reserve(nk * 8));
    DDB_TRY(S->znew.reserve(nk * 8));
    if (prm->algorithm != DDB_ALG_STLSQ) DDB_TRY(S->aux1.reserve(nk * 8));
    if (prm->algorithm == DDB_ALG_ADMM) DDB_TRY(S->aux2.reserve(nk * 8));
    DDB_TRY(S->active.reserve((size_t)n_t * W * 4));
    DDB_TRY(S->eq.reserve((size_t)n_t * sizeof(EqState)));
    DDB_TRY(S->lambdas_d.reserve((size_t)L * 8));
    DDB_TRY(S->cand_idx.reserve((size_t)S->pool_cap * 4));
    DDB_TRY(S->cand_val.reserve((size_t)S->pool_cap * 8));
    DDB_TRY(S->cand_off.reserve((size_t)S->cand_cap * 8));
    DDB_TRY(S->cand_nnz.reserve((size_t)S->cand_cap * 4));
    DDB_TRY(S->cand_dof.reserve((size_t)S->cand_cap * 4));
    DDB_TRY(S->cand_eq.reserve((size_t)S->cand_cap * 4));
    DDB_TRY(S->counters.reserve(64));
    DDB_TRY(S->pool_used_d.reserve(16));
    DDB_TRY(S->run_cand.reserve((size_t)n_t * S->run_cap * 4));
    DDB_TRY(S->run_jfirst.reserve((size_t)n_t * S->run_cap * 4));
    DDB_TRY(S->run_j.reserve((size_t)n_t * S->run_cap * 4));
    DDB_TRY(S->path_support.reserve((size_t)n_t * L * W * 4));
    DDB_TRY(S->path_coef.reserve((size_t)n_t * L * k * 8));
    DDB_TRY(S->path_iters.reserve((size_t)n_t * L * 4));
    DDB_TRY(S->big_list.reserve((size_t)n_t * 4 + 16));
    DDB_TRY(S->big_na.reserve((size_t)n_t * 4));
    DDB_TRY(S->big_fail.reserve((size_t)n_t * 4));
    DDB_TRY(S->active_matrix.reserve((size_t)n_t * 4));
    DDB_TRY(S->shared_fail.reserve(16));

    cudaStream_t st = ctx->stream;
    int launches = 0, big_steps = 0, pinv_steps = 0;
    DDB_CUDA_TRY(cudaMemcpyAsync(S->lambdas_d.p, S->lambdas.data(), (size_t)L * 8, cudaMemcpyHostToDevice, st));
    int h_counters[8] = {n_t, 0, 0, 0, 0, 0, 0, 0};  // candidates 0..n_t-1 are the zero models
    DDB_CUDA_TRY(cudaMemcpyAsync(S->counters.p, h_counters, sizeof(h_counters), cudaMemcpyHostToDevice, st));
    DDB_CUDA_TRY(cudaMemsetAsync(S->pool_used_d.p, 0, 16, st));
    DDB_CUDA_TRY(cudaMemsetAsync(S->shared_fail.p, 0, 16, st));
    DDB_CUDA_TRY(cudaMemsetAsync(S->path_iters.p, 0, (size_t)n_t * L * 4, st));
    DDB_CUDA_TRY(cudaMemsetAsync(S->path_support.p, 0, (size_t)n_t * L * W * 4, st));

    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[4], st));
    // ---- preparation + shared factorisation --------------------------------------------------
    const double* G = implicit ? ctx->imp_packed.as<double>() : ctx->d_packed.as<double>();
    const double* R = G + kk;
    const double rho = prm->rho;  // 0 for plain STLSQ
    {
        const size_t count = kk + nk + (size_t)n_t;
        nonfinite_check_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(G, count, S->shared_fail.as<int>() + 1);
        ++launches;
    }
    prep_diag_kernel<<<(k + 255) / 256, 256, 0, st>>>(G, k, rho, S->dinv.as<double>());
    prep_scale_kernel<<<(unsigned)((kk + 255) / 256), 256, 0, st>>>(G, k, rho, S->dinv.as<double>(), S->As.as<double>());
    prep_rhs_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, st>>>(R, k, n_t, S->dinv.as<double>(), S->rs.as<double>());
    launches += 3;
    const size_t solve_smem = ((size_t)((k + 31) / 32) * 32 + 32 * 33) * sizeof(double);
    if (solve_smem > 200 * 1024) {
        set_error("ddb_sweep: library of %d terms exceeds the shared-memory triangular solver (max ~24000)", k);
        return DDB_UNSUPPORTED;
    }
    DDB_CUDA_TRY(cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_smem));
    const bool shared_factor = !(implicit && prm->algorithm == DDB_ALG_STLSQ);  // see sweep_init_kernel
    // ADMM / SR3 iterations through the explicit inverse when the library is large (see ninv_apply_kernel)
    const char* ienv = getenv("DDB_SWEEP_INVERSE");
    const bool use_inverse = prm->algorithm != DDB_ALG_STLSQ && (ienv ? atoi(ienv) != 0 : k >= 256);
    if (use_inverse) DDB_TRY(S->zsol.reserve(nk * 8));
    const char* fenv = getenv("DDB_CHOL_FMA");  // 1: the previous 32-wide FMA-pipe factorisation (A/B measurements)
    const bool chol_fma = fenv && atoi(fenv) != 0;
    if (shared_factor) {
        DDB_CUDA_TRY(cudaMemcpyAsync(S->Lfac.p, S->As.p, kk * 8, cudaMemcpyDeviceToDevice, st));
        if (chol_fma) {
            DDB_CUDA_TRY(cudaMemcpyAsync(S->znew.p, S->rs.p, nk * 8, cudaMemcpyDeviceToDevice, st));
            if (implicit) {
                implicit_prep_kernel<<<(unsigned)((kk + 255) / 256), 256, 0, st>>>(nullptr, S->znew.as<double>(), k);
                ++launches;
            }
            DDB_TRY(chol_batched(ctx, S->Lfac.as<double>(), 0, k, nullptr, k, 1, S->shared_fail.as<int>(), &launches));
            chol_solve_kernel<<<dim3(n_t, 1), 256, solve_smem, st>>>(S->Lfac.as<double>(), 0, k, nullptr, k,
                                                                     S->znew.as<double>(), (size_t)k, n_t, nullptr, nullptr);
            ++launches;
        } else {
            // the n_t scaled right-hand sides are appended as rows k .. k + n_t - 1 of the factor's buffer: the
            // factorisation forward-substitutes them on the way, one backward kernel finishes the initial full solve
            double* zrows = S->Lfac.as<double>() + kk;
            DDB_CUDA_TRY(cudaMemcpyAsync(zrows, S->rs.p, nk * 8, cudaMemcpyDeviceToDevice, st));
            if (implicit) {  // zero the own entry of every target's right-hand side
                implicit_prep_kernel<<<(unsigned)((kk + 255) / 256), 256, 0, st>>>(nullptr, zrows, k);
                ++launches;
            }
            DDB_TRY(chol_factor_shared(ctx, S->Lfac.as<double>(), k, S->shared_fail.as<int>(), n_t, &launches));
            DDB_TRY(chol_back_batched(ctx, S->Lfac.as<double>(), 0, k, nullptr, k, 1, k, n_t, S->znew.as<double>(), 0, (size_t)k,
                                      nullptr, &launches));
        }
        if (implicit || use_inverse) {
            // inverse of the equilibrated matrix (identity right-hand sides)
            DDB_TRY(S->Ninv.reserve(kk * 8));
            implicit_prep_kernel<<<(unsigned)((kk + 255) / 256), 256, 0, st>>>(S->Ninv.as<double>(), nullptr, k);
            ++launches;
            DDB_TRY(chol_solve_many(ctx, S->Lfac.as<double>(), k, S->Ninv.as<double>(), k, &launches));
        }
    }
    SweepDev D = make_dev(ctx, S);
    if (use_inverse) D.zsol = S->zsol.as<double>();
    if (shared_factor && prm->algorithm == DDB_ALG_STLSQ) {
        // a factorisation that broke down (fewer samples than terms, collinear terms) is not an error -- the initial
        // least squares then runs as a masked solve with the minimum-norm fallback (pinv_solve_warp up to kPinvMax
        // terms, pinv_big_* through a big step beyond)
        D.init_masked = S->shared_fail.as<int>();
    }
    sweep_init_kernel<<<n_t, 256, 0, st>>>(D);
    ++launches;
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[5], st));

    // ---- the sweep -----------------------------------------------------------------------------
    int h_shared_fail = 0;
    if (prm->algorithm == DDB_ALG_STLSQ) {
        const size_t base_smem = ((size_t)kSmallMax * (kSmallMax + 1) + kSmallMax) * sizeof(double) + (kSmallMax + 2) * sizeof(int);
        const int x_in_smem = base_smem + 2 * (size_t)k * sizeof(double) <= 200 * 1024;
        const size_t small_smem = base_smem + (x_in_smem ? 2 * (size_t)k * sizeof(double) : 0);
        DDB_CUDA_TRY(cudaFuncSetAttribute(stlsq_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem));
        std::vector<int> h_big;
        for (int round = 0;; ++round) {
            stlsq_small_kernel<<<n_t, 256, small_smem, st>>>(D, kSmallMax, x_in_smem);
            count_status_kernel<<<1, 32, 0, st>>>(D.eq, n_t, D.counters);
            launches += 2;
            DDB_CUDA_TRY(cudaMemcpyAsync(h_counters, S->counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
            DDB_CUDA_TRY(cudaStreamSynchronize(st));
            if (h_counters[1]) {
                set_error("ddb_sweep: candidate storage overflow (%d candidates); too many distinct iterates", h_counters[0]);
                return DDB_OUT_OF_MEMORY;
            }
            const int nbig = h_counters[3];
            if (nbig == 0) break;
            // ---- big step: targets whose active set does not fit the shared-memory solver ----------
            ++big_steps;
            const size_t ws_stride = chol_fma ? kk : kk + (size_t)k;  // + one row: the right-hand side
            // Sharded run: the parked targets are dealt to the ranks round robin (SURVEY 8e "equations per GPU"), every
            // rank factorises its share and ONE sum all-reduce of [solution | failure flag] per parked target (zeros for
            // the slots of the other ranks, so the sum is exact and every rank ends up with bit-identical iterates)
            // replaces nranks redundant copies of the same factorisations.
            const bool shard_big = ctx->comm && ctx->nranks > 1 && !chol_fma && nbig >= 2;
            const int nr = shard_big ? ctx->nranks : 1, rk = shard_big ? ctx->rank : 0;
            const int nloc = nbig > rk ? (nbig - rk + nr - 1) / nr : 0;
            DDB_TRY(S->big_ws.reserve((size_t)std::max(1, nloc) * ws_stride * 8));
            DDB_TRY(S->big_idx.reserve((size_t)nbig * k * 4));
            int *na_loc = nullptr, *fail_loc = nullptr, *map_loc = nullptr;
            if (shard_big) {
                DDB_TRY(S->big_loc.reserve((size_t)3 * n_t * 4));
                DDB_TRY(S->big_xbuf.reserve((size_t)nbig * (k + 1) * 8));
                na_loc = S->big_loc.as<int>(), fail_loc = na_loc + n_t, map_loc = fail_loc + n_t;
                DDB_CUDA_TRY(cudaMemsetAsync(fail_loc, 0, (size_t)n_t * 4, st));
                DDB_CUDA_TRY(cudaMemsetAsync(S->big_xbuf.p, 0, (size_t)nbig * (k + 1) * 8, st));
            }
            list_big_kernel<<<1, 32, 0, st>>>(D.eq, n_t, S->big_list.as<int>(), S->big_list.as<int>() + n_t);
            DDB_CUDA_TRY(cudaMemsetAsync(S->big_fail.p, 0, (size_t)n_t * 4, st));
            stlsq_big_gather_kernel<<<dim3(64, nbig), 256, (size_t)k * sizeof(int), st>>>(
                D, S->big_list.as<int>(), S->big_ws.as<double>(), ws_stride, S->big_idx.as<int>(), S->big_na.as<int>(),
                chol_fma ? 0 : 1, nr, rk, na_loc, map_loc);
            launches += 2;
            if (shard_big) {
                if (nloc > 0) {
                    DDB_TRY(chol128_batched(ctx, S->big_ws.as<double>(), ws_stride, k, na_loc, k, nloc, fail_loc, 1, k, &launches));
                    DDB_TRY(chol_back_batched(ctx, S->big_ws.as<double>(), ws_stride, k, na_loc, k, nloc, k, 1,
                                              S->big_xbuf.as<double>(), (size_t)k + 1, 0, map_loc, nullptr));
                    big_fail_pack_kernel<<<(nloc + 63) / 64, 64, 0, st>>>(fail_loc, map_loc, nloc, S->big_xbuf.as<double>(), k);
                    launches += 2;
                }
                DDB_TRY(comm_allreduce(ctx, S->big_xbuf.as<double>(), (int64_t)nbig * (k + 1)));
                big_unpack_kernel<<<nbig, 256, 0, st>>>(S->big_xbuf.as<double>(), S->big_list.as<int>(), S->big_na.as<int>(), k,
                                                        S->znew.as<double>(), S->big_fail.as<int>());
                ++launches;
            } else if (chol_fma) {
                DDB_TRY(chol_batched(ctx, S->big_ws.as<double>(), kk, k, S->big_na.as<int>(), k, nbig,
                                     S->big_fail.as<int>(), &launches));
                chol_solve_kernel<<<dim3(1, nbig), 256, solve_smem, st>>>(S->big_ws.as<double>(), kk, k,
                                                                          S->big_na.as<int>(), k, S->znew.as<double>(),
                                                                          (size_t)k, 1, nullptr, S->big_list.as<int>());
            } else {
                DDB_TRY(chol128_batched(ctx, S->big_ws.as<double>(), ws_stride, k, S->big_na.as<int>(), k, nbig,
                                        S->big_fail.as<int>(), 1, k, &launches));
                DDB_TRY(chol_back_batched(ctx, S->big_ws.as<double>(), ws_stride, k, S->big_na.as<int>(), k, nbig, k, 1,
                                          S->znew.as<double>(), (size_t)k, 0, S->big_list.as<int>(), nullptr));
            }
            {   // factorisations that broke down: minimum-norm steps (pinv_big_*), as `A[p, :]' \ b` takes in the reference.
                // Sharded runs: the flags are the all-reduced ones, every rank repeats the same few fallback solves.
                h_big.resize((size_t)3 * nbig);
                DDB_CUDA_TRY(cudaMemcpyAsync(h_big.data(), S->big_fail.p, (size_t)nbig * 4, cudaMemcpyDeviceToHost, st));
                DDB_CUDA_TRY(cudaStreamSynchronize(st));
                bool any_fail = false;
                for (int b = 0; b < nbig; ++b) any_fail |= h_big[b] != 0;
                if (any_fail) {
                    int h_nf = 0;
                    DDB_CUDA_TRY(cudaMemcpyAsync(&h_nf, S->shared_fail.as<int>() + 1, 4, cudaMemcpyDeviceToHost, st));
                    DDB_CUDA_TRY(cudaMemcpyAsync(h_big.data() + nbig, S->big_na.p, (size_t)nbig * 4, cudaMemcpyDeviceToHost, st));
                    DDB_CUDA_TRY(cudaMemcpyAsync(h_big.data() + 2 * nbig, S->big_list.p, (size_t)nbig * 4, cudaMemcpyDeviceToHost, st));
                    DDB_CUDA_TRY(cudaStreamSynchronize(st));
                    for (int b = 0; b < nbig && !h_nf; ++b) {  // NaN / Inf in the blocks: reported below, nothing to solve
                        if (!h_big[b]) continue;
                        const int na = h_big[nbig + b], e = h_big[2 * nbig + b];
                        DDB_TRY(S->pinv_ws.reserve(((size_t)na * na + 3 * (size_t)na) * 8));
                        double* Gp = S->pinv_ws.as<double>();
                        double *lam = Gp + (size_t)na * na, *bp = lam + na, *cp = bp + na;
                        const int* idx = S->big_idx.as<int>() + (size_t)b * k;
                        pinv_big_gather_kernel<<<na, 256, 0, st>>>(D, e, idx, na, Gp, bp);
                        DDB_CUDA_TRY(cudaGetLastError());
                        DDB_TRY(eig_sym_device(ctx, Gp, na, lam));
                        pinv_big_project_kernel<<<(na + 7) / 8, 256, 0, st>>>(Gp, lam, na, bp, cp);
                        pinv_big_expand_kernel<<<(na + 255) / 256, 256, 0, st>>>(D, e, b, idx, na, Gp, cp, S->big_fail.as<int>());
                        launches += 3;
                        ++pinv_steps;
                    }
                }
            }
            stlsq_big_finish_kernel<<<nbig, 256, 0, st>>>(D, S->big_list.as<int>(), S->big_idx.as<int>(),
                                                          S->big_na.as<int>(), S->big_fail.as<int>());
            launches += 2;
            DDB_CUDA_TRY(cudaGetLastError());
        }
    } else {
        const int64_t max_steps = steps_bound;
        for (int64_t it = 0; it < max_steps; ++it) {
            relax_rhs_kernel<<<n_t, 256, 0, st>>>(D, S->active_matrix.as<int>());
            if (use_inverse)
                ninv_apply_kernel<<<dim3((k + 15) / 16, (n_t + 63) / 64), 256, 0, st>>>(
                    S->Ninv.as<double>(), k, S->znew.as<double>(), n_t, S->active_matrix.as<int>(), S->zsol.as<double>());
            else
                chol_solve_kernel<<<dim3(1, n_t), 256, solve_smem, st>>>(S->Lfac.as<double>(), 0, k, nullptr, k,
                                                                         S->znew.as<double>(), (size_t)k, 1,
                                                                         S->active_matrix.as<int>(), nullptr);
            relax_finish_kernel<<<n_t, 256, 0, st>>>(D);
            launches += 3;
            if ((it & 7) == 7 || it + 1 == max_steps) {
                count_status_kernel<<<1, 32, 0, st>>>(D.eq, n_t, D.counters);
                ++launches;
                DDB_CUDA_TRY(cudaMemcpyAsync(h_counters, S->counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
                DDB_CUDA_TRY(cudaStreamSynchronize(st));
                if (h_counters[1]) {
                    set_error("ddb_sweep: candidate storage overflow (%d candidates)", h_counters[0]);
                    return DDB_OUT_OF_MEMORY;
                }
                if (h_counters[4] == 0) break;
            }
        }
        DDB_CUDA_TRY(cudaGetLastError());
    }
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[6], st));
    unsigned long long h_pool = 0;
    DDB_CUDA_TRY(cudaMemcpyAsync(h_counters, S->counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaMemcpyAsync(&h_pool, S->pool_used_d.p, 8, cudaMemcpyDeviceToHost, st));
    int h_nonfinite = 0;
    DDB_CUDA_TRY(cudaMemcpyAsync(&h_shared_fail, S->shared_fail.p, 4, cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaMemcpyAsync(&h_nonfinite, S->shared_fail.as<int>() + 1, 4, cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    if (h_counters[1]) {
        set_error("ddb_sweep: candidate storage overflow");
        return DDB_OUT_OF_MEMORY;
    }
    {   // canonical (rank-independent) candidate numbering, see canon_base_kernel
        const int ncand = h_counters[0], nrec = ncand - n_t;
        if (nrec > 0) {
            const size_t o_nnz = (size_t)ncand * 8, o_dof = o_nnz + (size_t)ncand * 4, o_eq = o_dof + (size_t)ncand * 4;
            DDB_TRY(S->canon_tmp.reserve(o_eq + (size_t)ncand * 4 + (size_t)n_t * 4 + 64));
            unsigned char* tb = static_cast<unsigned char*>(S->canon_tmp.p);
            int64_t* t_off = reinterpret_cast<int64_t*>(tb);
            int32_t* t_nnz = reinterpret_cast<int32_t*>(tb + o_nnz);
            int32_t* t_dof = reinterpret_cast<int32_t*>(tb + o_dof);
            int32_t* t_eq = reinterpret_cast<int32_t*>(tb + o_eq);
            int* t_base = reinterpret_cast<int*>(tb + o_eq + (size_t)ncand * 4);
            canon_base_kernel<<<1, 1024, 0, st>>>(D.eq, n_t, t_base);
            canon_permute_kernel<<<n_t, 256, 0, st>>>(D, t_base, t_off, t_nnz, t_dof, t_eq);
            DDB_CUDA_TRY(cudaMemcpyAsync(D.cand_off + n_t, t_off + n_t, (size_t)nrec * 8, cudaMemcpyDeviceToDevice, st));
            DDB_CUDA_TRY(cudaMemcpyAsync(D.cand_nnz + n_t, t_nnz + n_t, (size_t)nrec * 4, cudaMemcpyDeviceToDevice, st));
            DDB_CUDA_TRY(cudaMemcpyAsync(D.cand_dof + n_t, t_dof + n_t, (size_t)nrec * 4, cudaMemcpyDeviceToDevice, st));
            DDB_CUDA_TRY(cudaMemcpyAsync(D.cand_eq + n_t, t_eq + n_t, (size_t)nrec * 4, cudaMemcpyDeviceToDevice, st));
            launches += 2;
            DDB_CUDA_TRY(cudaGetLastError());
            DDB_CUDA_TRY(cudaEventRecord(ctx->ev[6], st));  // the renumbering is part of the sweep's wall time
            DDB_CUDA_TRY(cudaStreamSynchronize(st));
        }
    }
    comm_collect_time(ctx);
    float ms_f = 0.f, ms_s = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms_f, ctx->ev[4], ctx->ev[5]));
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms_s, ctx->ev[5], ctx->ev[6]));
    ctx->timings.factor_ms = ms_f;
    ctx->timings.sweep_ms = ms_s;
    ctx->timings.sweep_launches = launches;
    ctx->timings.big_steps = big_steps;
    ctx->timings.candidates = h_counters[0];
    S->ncand = h_counters[0];
    S->pool_used = (int64_t)h_pool;
    S->valid = true;
    if (h_nonfinite) {
        S->valid = false;
        set_error("ddb_sweep: the normal-equation blocks contain NaN or Inf (one or more measurements contain `NaN` or `Inf`)");
        return DDB_NONFINITE_INPUT;
    }
    if (h_shared_fail && !(shared_factor && prm->algorithm == DDB_ALG_STLSQ)) {  // STLSQ: solved by masked minimum-norm steps
        set_error("ddb_sweep: the equilibrated normal matrix is not positive definite (rank-deficient library "
                  "or fewer samples than terms); this backend solves through the Gram matrix and has no QR fallback");
        return DDB_NOT_POSITIVE_DEFINITE;
    }
    return DDB_OK;
}

int select_run(ddb_ctx* ctx, double* coef, double* lambda_opt, int32_t* iters, double* rss, int32_t* dof,
               uint32_t* support_path, double* coef_path, int32_t* iters_path, int32_t* flags) {
    SweepState* S = ctx->sweep;
    if (!S || !S->valid || !S->have_rss) {
        set_error("ddb_select: run ddb_sweep and ddb_rss first");
        return DDB_BAD_STATE;
    }
    const int k = S->k, n_t = S->n_t, L = S->L, W = S->W;
    cudaStream_t st = ctx->stream;
    // device staging: coef (n_t*k) | lambda (n_t) | rss (n_t) | iters (n_t) | dof (n_t) | flags (n_t)
    const size_t nk = (size_t)n_t * k;
    DDB_TRY(S->out_tmp.reserve(nk * 8 + (size_t)n_t * (8 + 8 + 4 + 4 + 4) + 64));
    double* d_coef = S->out_tmp.as<double>();
    double* d_lam = d_coef + nk;
    double* d_rss = d_lam + n_t;
    int32_t* d_it = reinterpret_cast<int32_t*>(d_rss + n_t);
    int32_t* d_dof = d_it + n_t;
    int32_t* d_fl = d_dof + n_t;
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[4], st));
    SweepDev D = make_dev(ctx, S);
    select_kernel<<<n_t, 256, 0, st>>>(D, S->rss_ptr, S->rss_corr.as<double>(), S->prm.selector, (double)S->m_total, d_coef, d_lam, d_it, d_rss,
                                       d_dof, d_fl);
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[5], st));
    if (coef) DDB_CUDA_TRY(cudaMemcpyAsync(coef, d_coef, nk * 8, cudaMemcpyDeviceToHost, st));
    if (lambda_opt) DDB_CUDA_TRY(cudaMemcpyAsync(lambda_opt, d_lam, (size_t)n_t * 8, cudaMemcpyDeviceToHost, st));
    if (rss) DDB_CUDA_TRY(cudaMemcpyAsync(rss, d_rss, (size_t)n_t * 8, cudaMemcpyDeviceToHost, st));
    if (iters) DDB_CUDA_TRY(cudaMemcpyAsync(iters, d_it, (size_t)n_t * 4, cudaMemcpyDeviceToHost, st));
    if (dof) DDB_CUDA_TRY(cudaMemcpyAsync(dof, d_dof, (size_t)n_t * 4, cudaMemcpyDeviceToHost, st));
    if (flags) DDB_CUDA_TRY(cudaMemcpyAsync(flags, d_fl, (size_t)n_t * 4, cudaMemcpyDeviceToHost, st));
    if (support_path)
        DDB_CUDA_TRY(cudaMemcpyAsync(support_path, S->path_support.p, (size_t)n_t * L * W * 4, cudaMemcpyDeviceToHost, st));
    if (coef_path)
        DDB_CUDA_TRY(cudaMemcpyAsync(coef_path, S->path_coef.p, (size_t)n_t * L * k * 8, cudaMemcpyDeviceToHost, st));
    if (iters_path)
        DDB_CUDA_TRY(cudaMemcpyAsync(iters_path, S->path_iters.p, (size_t)n_t * L * 4, cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[6], st));
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    comm_collect_time(ctx);
    float a = 0.f, b = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&a, ctx->ev[4], ctx->ev[5]));
    DDB_CUDA_TRY(cudaEventElapsedTime(&b, ctx->ev[5], ctx->ev[6]));
    ctx->timings.select_ms = a;
    ctx->timings.d2h_ms = b;
    ctx->timings.sweep_launches += 1;
    return DDB_OK;
}

}  // namespace ddb

using namespace ddb;

extern "C" {

int ddb_implicit_select(ddb_ctx* ctx, const int32_t* idx, int kk) {
    if (!ctx) {
        set_error("ddb_implicit_select: null context");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    return implicit_select(ctx, idx, kk);
}

int ddb_sweep(ddb_ctx* ctx, const ddb_sweep_params* params, const double* lambdas, int L, int64_t m_total) {
    if (!ctx) {
        set_error("ddb_sweep: null context");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    return sweep_run(ctx, params, lambdas, L, m_total);
}

int64_t ddb_rss_count(const ddb_ctx* ctx) {
    if (!ctx || !ctx->sweep || !ctx->sweep->valid) return 0;
    return ctx->sweep->ncand;
}

int ddb_rss(ddb_ctx* ctx, double* dRss) {
    if (!ctx) {
        set_error("ddb_rss: null context");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (!ctx->sweep || !ctx->sweep->valid) {
        set_error("ddb_rss: run ddb_sweep first");
        return DDB_BAD_STATE;
    }
    return rss_run(ctx, ctx->sweep, dRss);
}

int ddb_rss_buffer(ddb_ctx* ctx, double** dRss) {
    if (!ctx || !dRss) {
        set_error("ddb_rss_buffer: null argument");
        return DDB_INVALID_ARG;
    }
    if (!ctx->sweep || !ctx->sweep->valid) {
        set_error("ddb_rss_buffer: run ddb_sweep first");
        return DDB_BAD_STATE;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    DDB_TRY(ctx->sweep->rss.reserve((size_t)std::max(1, ctx->sweep->ncand) * 8));
    *dRss = ctx->sweep->rss.as<double>();
    return DDB_OK;
}

int ddb_select(ddb_ctx* ctx, double* coef, double* lambda_opt, int32_t* iters, double* rss, int32_t* dof,
               uint32_t* support_path, double* coef_path, int32_t* iters_path, int32_t* flags) {
    if (!ctx) {
        set_error("ddb_select: null context");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    return select_run(ctx, coef, lambda_opt, iters, rss, dof, support_path, coef_path, iters_path, flags);
}

int ddb_solve_sparse(ddb_ctx* ctx, const double* X, const double* Y, int n_t, int64_t m,
                     const ddb_sweep_params* params, const double* lambdas, int L, double* coef, double* lambda_opt,
                     int32_t* iters, double* rss, int32_t* dof, uint32_t* support_path, int32_t* flags) {
    DDB_TRY(ddb_upload_gram(ctx, X, Y, n_t, m, nullptr));  // copies overlapped by the Gram kernels
    DDB_TRY(ddb_sweep(ctx, params, lambdas, L, m));
    DDB_TRY(ddb_rss(ctx, nullptr));
    return ddb_select(ctx, coef, lambda_opt, iters, rss, dof, support_path, nullptr, nullptr, flags);
}

}  // extern "C"
