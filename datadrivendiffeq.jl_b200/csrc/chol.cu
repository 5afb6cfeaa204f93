// Blocked Cholesky of the equilibrated normal matrices with 128-wide block columns, the bulk of the work on the FP64
// tensor pipe (DMMA.8x8x4), batched over matrices, with the forward substitution of the right-hand sides folded in.
//
// Used for the factorisation every target shares (initial full least squares of STLSQ -- STLSQ.jl:93-99 through the
// normal equations; the fixed factor of ADMM / SR3 -- ADMM.jl:78-81, SR3.jl:108-111), for the masked systems of
// STLSQ steps whose active set exceeds the shared-memory solver (STLSQ.jl:128-140; one matrix per parked target) and
// for P of the DMD operator (lib/DataDrivenDMD/src/algorithms.jl:80-83).
//
// Right-looking, row-major lower triangle in place.  Per block column j (128 wide):
//   1. chol_diag128_kernel    one CTA per matrix: the 128 x 128 diagonal block is factorised in shared memory in four
//                             32-column panels -- the 32 x 32 pivot block in the registers of one warp (lane = row,
//                             shuffles), the rows below it by substitution (one thread per row), the rest of the block
//                             as a small DMMA product;
//   2. chol_panel128_kernel   rows below, 128 per CTA:  X L_jj' = A[rows, j]  by blocked substitution -- for each of
//                             the four 32-column sub-blocks a DMMA product with the sub-blocks already solved, then a
//                             32-step substitution per row.  No explicit inverse of the diagonal block is used
//                             (an inverse-based panel is not backward stable; the README Lorenz system, condition
//                             1e11 after equilibration, shows it).  The result is also written TRANSPOSED into the
//                             unused upper triangle, so the buffer holds L below and L' above the diagonal: the
//                             backward substitution then reads contiguous rows;
//   3. chol_gemm_kernel<SYRK> A[rest, rest] -= L[rest, j] L[rest, j]'  (DMMA, lower 128 x 128 tiles only).
// Right-hand sides appended as EXTRA ROWS below the matrix (row index `xrow` onwards, same leading dimension) pass
// through steps 2 and 3 like any other row: when the factorisation ends they hold y = L^-1 b (the last rows of the
// Cholesky factor of the bordered matrix [[A, b], [b', .]] are b' L^-T).  chol_back_kernel finishes with x = L^-T y.
// The previous version walked 32-wide block columns on the FMA pipe (67 passes over the trailing matrix).
#include "common.cuh"
#include "mma_common.cuh"
#include <algorithm>

namespace ddb {

constexpr int kCB = 128;           // block column width
constexpr int kCP = kCB + 4;       // shared-memory pitch: 132 = 4 (mod 16), conflict-free DMMA fragment loads
constexpr int kCK = 32;            // contraction chunk of the GEMMs
constexpr int kCKP = kCK + 4;      // 36 = 4 (mod 16)
constexpr int kCStage = 2 * kCB * kCKP;
constexpr double kPivotTol = 64.0 * 2.220446049250313e-16;  // pivots of the equilibrated (unit-diagonal) matrices below this: breakdown

struct CholBatch {
    double* base;        // matrix b at base + b * stride, row-major, leading dimension ld
    size_t stride;
    int ld;
    const int* na_arr;   // order of matrix b, or null: all of order na_all
    int na_all;
    int* fail;           // per matrix: set to 1 when a pivot is not positive
    int extra;           // right-hand sides appended as rows xrow .. xrow + extra - 1 (0: none)
    int xrow;
    long long* trace;    // debugging (tools/chol_bench.cu): clock64 at the phase boundaries of CTA 0; null in production
};

__device__ __forceinline__ void cp8c(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// One 8 x 8 output sub-tile of C (+)= A B' on shared-memory operands (A, B point at the first of their 8 rows, the
// contraction index is contiguous): c0 = C[g][2t], c1 = C[g][2t + 1] with g = lane / 4, t = lane % 4.
__device__ __forceinline__ void frag_nt(double& c0, double& c1, const double* A, int lda, const double* B, int ldb, int K,
                                        int lane) {
    const int g = lane >> 2, t = lane & 3;
    const double* pa = A + g * lda + t;
    const double* pb = B + g * ldb + t;
    for (int kk = 0; kk < K; kk += 4) dmma884(c0, c1, pa[kk], pb[kk]);
}

// 1 / sqrt(d) for the pivots: single-precision estimate + two Newton steps in FP64 (~100 cycles on the serial chain of
// the pivot block instead of ~250 for sqrt + reciprocal; within 2 ulp).  Outside the single-precision range: slow path.
__device__ __forceinline__ double fast_rsqrt(double d) {
    if (!(d > 1e-30 && d < 1e30)) return 1.0 / sqrt(d);
    double y = (double)rsqrtf((float)d);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double r = fma(-d * y, y, 1.0);   // 1 - d y^2
        y = fma(0.5 * y, r, y);
    }
    return y;
}

// One row of X Ld' = A against a 32 x 32 lower-triangular block in shared memory (Ld[j][c] at ldp[j * pitch + c],
// reciprocal diagonal in rinv): right-looking in registers, 496 FMAs.
__device__ __forceinline__ void row_subst32(double* row, const double* ldp, int pitch, const double* rinv) {
    double x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = row[c];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        x[c] *= rinv[c];
#pragma unroll
        for (int j = c + 1; j < 32; ++j) x[j] = fma(-x[c], ldp[j * pitch + c], x[j]);
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) row[c] = x[c];
}

__global__ void __launch_bounds__(256, 1) chol_diag128_kernel(CholBatch cb, int j0) {
    const int b = blockIdx.x;
    const int na = cb.na_arr ? cb.na_arr[b] : cb.na_all;
    if (j0 >= na) return;
    const int nb = min(kCB, na - j0);
    extern __shared__ double dsm[];
    double* S = dsm;  // 128 x 132
    __shared__ double rinv[32];
    __shared__ int s_bad;
    double* A = cb.base + (size_t)b * cb.stride;
    const int ld = cb.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    if (tid == 0) s_bad = 0;
    long long tr0 = 0, tr_piv = 0, tr_sub = 0, tr_upd = 0, tr_mark = 0;
    const bool tracing = cb.trace && b == 0 && tid == 0;
    if (tracing) tr0 = clock64();
    for (int e = tid; e < kCB * kCB; e += 256) {  // asynchronous copies: every load of the block is in flight at once
        const int i = e >> 7, j = e & 127;
        if (i < nb && j <= i) cp8c(S + i * kCP + j, A + (size_t)(j0 + i) * ld + j0 + j);
        else S[i * kCP + j] = (i == j ? 1.0 : 0.0);
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (tracing) tr_mark = clock64(), cb.trace[0] += tr_mark - tr0;

#pragma unroll 1
    for (int c0 = 0; c0 < nb; c0 += 32) {
        // ---- 1. the 32 x 32 pivot block: lane = row, the row lives in registers, columns one after the other
        if (warp == 0) {
            double a[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] = S[(c0 + lane) * kCP + c0 + j];
            int bad = 0;
            double my_rinv = 1.0;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                double d = __shfl_sync(0xffffffffu, a[c], c);
                if (!(d > kPivotTol)) {  // unit-diagonal (equilibrated) matrices: numerically rank deficient
                    bad = 1;
                    d = 1.0;  // keep going with a harmless pivot; the flag invalidates the result
                }
                const double ri = fast_rsqrt(d);
                const double piv = d * ri;
                const double l = (lane == c) ? piv : a[c] * ri;
                a[c] = l;
                if (lane == c) my_rinv = ri;
#pragma unroll
                for (int j = c + 1; j < 32; ++j) {
                    // unconditional: entries above the diagonal (j > lane) hold garbage that is never read by a
                    // valid lane (d comes from lane c, lj from lane j) and never stored.  (Broadcasting the column
                    // through shared memory instead of these shuffles was measured: not faster -- the block's time is
                    // the chain of ~11 dependent FP64 operations per column, not the shuffle unit.)
                    const double lj = __shfl_sync(0xffffffffu, l, j);
                    a[j] = fma(-l, lj, a[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j <= lane) S[(c0 + lane) * kCP + c0 + j] = a[j];
            rinv[lane] = my_rinv;
            if (bad && lane == 0) s_bad = 1;
        }
        __syncthreads();
        if (tracing) { const long long t = clock64(); tr_piv += t - tr_mark; tr_mark = t; }
        const int t0 = c0 + 32, Rr = kCB - t0;  // rows / columns of the block that are still open
        if (Rr > 0) {
            // ---- 2. rows below the pivot block: X Ld' = A, one thread per row
            if (tid < Rr) row_subst32(S + (t0 + tid) * kCP + c0, S + c0 * kCP + c0, kCP, rinv);
            __syncthreads();
            if (tracing) { const long long t = clock64(); tr_sub += t - tr_mark; tr_mark = t; }
            // ---- 3. rest of the block:  S[i][j] -= sum_t X[i][t] X[j][t]   (lower 8 x 8 sub-tiles, DMMA)
            const int T8 = Rr / 8;
            for (int q = warp; q < T8 * T8; q += 8) {
                const int i8 = q / T8, j8 = q % T8;
                if (j8 > i8) continue;
                double u0 = 0.0, u1 = 0.0;
                frag_nt(u0, u1, S + (t0 + 8 * i8) * kCP + c0, kCP, S + (t0 + 8 * j8) * kCP + c0, kCP, 32, lane);
                const int r = 8 * i8 + g, c = 8 * j8 + 2 * t4;
                double* o = S + (t0 + r) * kCP + t0 + c;
                if (c <= r) o[0] -= u0;
                if (c + 1 <= r) o[1] -= u1;
            }
            __syncthreads();
            if (tracing) { const long long t = clock64(); tr_upd += t - tr_mark; tr_mark = t; }
        }
    }
    if (tracing) cb.trace[1] += tr_piv, cb.trace[2] += tr_sub, cb.trace[3] += tr_upd;
    if (tid == 0 && s_bad) atomicExch(&cb.fail[b], 1);
    // the factor goes back in place: lower triangle of the block, and mirrored above the diagonal (L' for the
    // backward substitution, like the panel step does for the rows below)
    for (int e = tid; e < kCB * kCB; e += 256) {
        const int i = e >> 7, j = e & 127;  // row i of the buffer: L[i][j] for j <= i, L'[i][j] = L[j][i] beyond (coalesced)
        if (i < nb && j < nb) A[(size_t)(j0 + i) * ld + j0 + j] = (j <= i) ? S[i * kCP + j] : S[j * kCP + i];
    }
    if (tracing) cb.trace[4] += clock64() - tr_mark, cb.trace[5] += 1;
}

// Rows below the diagonal block (and the appended right-hand sides), 128 per CTA:  X L_jj' = A[rows, j0 .. j0 + nb).
// Logical row r: physical row t0 + r for r < rows = na - t0, xrow + (r - rows) for the right-hand sides.
// grid (row tiles, batch)
__global__ void __launch_bounds__(256, 1) chol_panel128_kernel(CholBatch cb, int j0) {
    const int b = blockIdx.y;
    const int na = cb.na_arr ? cb.na_arr[b] : cb.na_all;
    if (j0 >= na) return;
    const int nb = min(kCB, na - j0);
    const int t0 = j0 + nb, rows = na - t0, rows_all = rows + cb.extra;
    const int r0 = blockIdx.x * kCB;
    if (r0 >= rows_all) return;
    extern __shared__ double psm[];
    double* At = psm;                 // 128 x 132: this CTA's rows of the panel
    double* Lq = psm + kCB * kCP;     // 32 x 132: block row q of L_jj (columns 0 .. 32 q + 31)
    __shared__ double rinv[32];
    double* A = cb.base + (size_t)b * cb.stride;
    const int ld = cb.ld;
    const int xoff = cb.xrow - t0 - rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    for (int e = tid; e < kCB * kCB; e += 256) {
        const int r = e >> 7, c = e & 127, rr = r0 + r;
        if (rr < rows_all && c < nb) cp8c(At + r * kCP + c, A + (size_t)(t0 + rr + (rr >= rows ? xoff : 0)) * ld + j0 + c);
        else At[r * kCP + c] = 0.0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll 1
    for (int cq = 0; cq < nb; cq += 32) {
        const int nbq = min(32, nb - cq);
        __syncthreads();  // previous sub-block finished with Lq
        for (int e = tid; e < 32 * (cq + 32); e += 256) {
            const int i = e / (cq + 32), j = e % (cq + 32);
            if (i < nbq && j <= cq + i) cp8c(Lq + i * kCP + j, A + (size_t)(j0 + cq + i) * ld + j0 + j);
            else Lq[i * kCP + j] = (j == cq + i ? 1.0 : 0.0);
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");  // also covers the tile itself
        __syncthreads();
        if (tid < 32) rinv[tid] = 1.0 / Lq[tid * kCP + cq + tid];
        if (cq > 0) {
            // At[:, cq .. cq + 31] -= At[:, 0 .. cq) Lq[:, 0 .. cq)'   (16 x 4 sub-tiles of 8 x 8, eight per warp)
#pragma unroll 1
            for (int s = 0; s < 8; ++s) {
                const int q = warp + 8 * s, i8 = q >> 2, j8 = q & 3;
                double u0 = 0.0, u1 = 0.0;
                frag_nt(u0, u1, At + 8 * i8 * kCP, kCP, Lq + 8 * j8 * kCP, kCP, cq, lane);
                double* o = At + (8 * i8 + g) * kCP + cq + 8 * j8 + 2 * t4;
                o[0] -= u0;
                o[1] -= u1;
            }
        }
        __syncthreads();
        if (tid < kCB) row_subst32(At + tid * kCP + cq, Lq + cq, kCP, rinv);
    }
    __syncthreads();
    for (int e = tid; e < kCB * kCB; e += 256) {
        const int r = e >> 7, c = e & 127, rr = r0 + r;
        if (rr < rows_all && c < nb) A[(size_t)(t0 + rr + (rr >= rows ? xoff : 0)) * ld + j0 + c] = At[r * kCP + c];
    }
    // L' above the diagonal: A[j0 + c][t0 + rr] for the matrix rows (coalesced along rr)
    for (int e = tid; e < kCB * kCB; e += 256) {
        const int c = e >> 7, r = e & 127, rr = r0 + r;
        if (rr < rows && c < nb) A[(size_t)(j0 + c) * ld + t0 + rr] = At[r * kCP + c];
    }
}

// SYRK:   A[row(r)][t0 + c] -= sum_t A[row(r)][j0 + t] A[t0 + c][j0 + t]   (c < rows, c <= r), lower tiles only
// t0 = j0 + nb, rows = na - t0; r runs over the matrix rows below the block column and then over the appended
// right-hand sides: row(r) = t0 + r for r < rows, xrow + (r - rows) otherwise.  grid (column tiles, row tiles, batch)
__global__ void __launch_bounds__(256, 1) chol_syrk128_kernel(CholBatch cb, int j0) {
    const int b = blockIdx.z;
    const int na = cb.na_arr ? cb.na_arr[b] : cb.na_all;
    if (j0 >= na) return;
    const int nb = min(kCB, na - j0);
    const int t0 = j0 + nb, rows = na - t0, rows_all = rows + cb.extra;
    const int r0 = blockIdx.y * kCB, c0 = blockIdx.x * kCB;
    if (r0 >= rows_all || c0 > r0 || c0 >= rows) return;
    extern __shared__ __align__(16) double gsm[];
    double* A = cb.base + (size_t)b * cb.stride;
    const int ld = cb.ld;
    const double* Bop = A + (size_t)(t0 + c0) * ld + j0;
    const size_t ldb = (size_t)ld;
    const int bcols = rows - c0;  // live rows of the B operand
    const int xoff = cb.xrow - t0 - rows;     // physical row of r >= rows: t0 + r + xoff
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = (warp >> 2) * 64, col0 = (warp & 3) * 32;
    const int nchunks = (nb + kCK - 1) / kCK;

    auto stage_load = [&](int chunk, int buf) {
        double* sa = gsm + (size_t)buf * kCStage;
        double* sb = sa + kCB * kCKP;
        const int k0 = chunk * kCK;
#pragma unroll 4
        for (int e = tid; e < kCB * kCK; e += 256) {
            const int r = e >> 5, kk = e & 31;
            double* da = sa + r * kCKP + kk;
            const int rr = r0 + r;
            if (rr < rows_all && k0 + kk < nb) cp8c(da, A + (size_t)(t0 + rr + (rr >= rows ? xoff : 0)) * ld + j0 + k0 + kk);
            else *da = 0.0;
            double* db = sb + r * kCKP + kk;
            if (r < bcols && k0 + kk < nb) cp8c(db, Bop + (size_t)r * ldb + k0 + kk);
            else *db = 0.0;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    double acc[8][4][2];
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
        const double* sa = gsm + (size_t)buf * kCStage + (row0 + g) * kCKP + t;
        const double* sb = gsm + (size_t)buf * kCStage + kCB * kCKP + (col0 + g) * kCKP + t;
#pragma unroll
        for (int kk = 0; kk < kCK / 4; ++kk) {
            double a[8], bq[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = sa[8 * i * kCKP + 4 * kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) bq[j] = sb[8 * j * kCKP + 4 * kk];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bq[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = r0 + row0 + 8 * i + g;
        if (r >= rows_all) continue;
        double* arow = A + (size_t)(t0 + r + (r >= rows ? xoff : 0)) * ld;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = c0 + col0 + 8 * j + 2 * t + h;
                // one thread per element and launch: the reduction (RED.ADD.F64, fire and forget) gives exactly
                // A - acc without the load -> subtract -> store round trip on the tile's critical path
                if (c < rows && c <= r) atomicAdd(arow + t0 + c, -acc[i][j][h]);
            }
        }
    }
}

// The same trailing update with 64 x 64 tiles, four warps (32 x 32 patches) and three CTAs per SM: a single matrix of
// order 2145 gives at most 136 tiles of 128 x 128 -- one wave whose length is the latency of ONE tile (17 us of DMMA
// on one SM) whatever the tile count; quarter-size tiles cut that latency by four and still fill the machine
// (544 tiles on 444 slots).  Batches that fill the machine with 128 x 128 tiles keep those (fewer fragment loads per
// DMMA).  grid (column tiles, row tiles, batch)
constexpr int kSB = 64;
constexpr int kSStage = 2 * kSB * kCKP;
__global__ void __launch_bounds__(128, 3) chol_syrk64_kernel(CholBatch cb, int j0) {
    const int b = blockIdx.z;
    const int na = cb.na_arr ? cb.na_arr[b] : cb.na_all;
    if (j0 >= na) return;
    const int nb = min(kCB, na - j0);
    const int t0 = j0 + nb, rows = na - t0, rows_all = rows + cb.extra;
    const int r0 = blockIdx.y * kSB, c0 = blockIdx.x * kSB;
    if (r0 >= rows_all || c0 > r0 || c0 >= rows) return;
    extern __shared__ __align__(16) double gsm[];
    double* A = cb.base + (size_t)b * cb.stride;
    const int ld = cb.ld;
    const double* Bop = A + (size_t)(t0 + c0) * ld + j0;
    const int bcols = rows - c0;
    const int xoff = cb.xrow - t0 - rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = (warp >> 1) * 32, col0 = (warp & 1) * 32;
    const int nchunks = (nb + kCK - 1) / kCK;

    auto stage_load = [&](int chunk, int buf) {
        double* sa = gsm + (size_t)buf * kSStage;
        double* sb = sa + kSB * kCKP;
        const int k0 = chunk * kCK;
#pragma unroll 4
        for (int e = tid; e < kSB * kCK; e += 128) {
            const int r = e >> 5, kk = e & 31;
            double* da = sa + r * kCKP + kk;
            const int rr = r0 + r;
            if (rr < rows_all && k0 + kk < nb) cp8c(da, A + (size_t)(t0 + rr + (rr >= rows ? xoff : 0)) * ld + j0 + k0 + kk);
            else *da = 0.0;
            double* db = sb + r * kCKP + kk;
            if (r < bcols && k0 + kk < nb) cp8c(db, Bop + (size_t)r * ld + k0 + kk);
            else *db = 0.0;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
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
        const double* sa = gsm + (size_t)buf * kSStage + (row0 + g) * kCKP + t;
        const double* sb = gsm + (size_t)buf * kSStage + kSB * kCKP + (col0 + g) * kCKP + t;
#pragma unroll
        for (int kk = 0; kk < kCK / 4; ++kk) {
            double a[4], bq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sa[8 * i * kCKP + 4 * kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) bq[j] = sb[8 * j * kCKP + 4 * kk];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bq[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + row0 + 8 * i + g;
        if (r >= rows_all) continue;
        double* arow = A + (size_t)(t0 + r + (r >= rows ? xoff : 0)) * ld;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = c0 + col0 + 8 * j + 2 * t + h;
                if (c < rows && c <= r) atomicAdd(arow + t0 + c, -acc[i][j][h]);  // RED: see chol_syrk128_kernel
            }
        }
    }
}

// Backward substitution x = L^-T y, one CTA per (right-hand side, matrix), left-looking over 32-row blocks from the
// bottom: the contributions of the entries already solved are dot products along the ROWS of L' (stored above the
// diagonal by the PANEL step: coalesced), the 32 x 32 diagonal block is substituted by one warp.
// y: row `yrow0 + rhs` of the matrix buffer (where the factorisation left L^-1 b); x -> xout + (xmap ? xmap[b] : b) *
// xstride_b + rhs * xstride_r.
__global__ void __launch_bounds__(1024) chol_back_kernel(CholBatch cb, int yrow0, double* xout, size_t xstride_b, size_t xstride_r,
                                                        const int* xmap) {
    const int b = blockIdx.y, rhs = blockIdx.x;
    const int na = cb.na_arr ? cb.na_arr[b] : cb.na_all;
    if (na <= 0) return;
    const double* A = cb.base + (size_t)b * cb.stride;
    const int ld = cb.ld;
    extern __shared__ double bsm[];
    double* x = bsm;                          // na (padded to 32)
    double* blk = bsm + ((na + 31) / 32) * 32;  // 32 x 33: the diagonal block, blk[i][j] = L[j0 + j][j0 + i] (j >= i)
    __shared__ double part[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* y = A + (size_t)(yrow0 + rhs) * ld;
    for (int i = tid; i < na; i += blockDim.x) x[i] = y[i];
    __syncthreads();
    for (int j0 = ((na - 1) / 32) * 32; j0 >= 0; j0 -= 32) {
        const int nbk = min(32, na - j0), end = j0 + nbk;
        for (int e = tid; e < 32 * 32; e += blockDim.x) {
            const int i = e >> 5, j = e & 31;  // blk[i][j] = L[j0 + j][j0 + i] for j >= i (lower triangle, read transposed)
            blk[i * 33 + j] = (i < nbk && j < nbk && j >= i) ? A[(size_t)(j0 + j) * ld + j0 + i] : (i == j ? 1.0 : 0.0);
        }
        // s_i = sum_{j >= end} L'[j0 + i][j] x[j]: one warp per row (32 warps: the memory-level parallelism comes from
        // the thread count; the compiler interleaves load and use inside a thread whatever the source says)
        if (warp < nbk) {
            const double* row = A + (size_t)(j0 + warp) * ld;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int j = end + lane;
            for (; j + 224 < na; j += 256) {
                const double v0 = __ldg(row + j), v1 = __ldg(row + j + 32), v2 = __ldg(row + j + 64), v3 = __ldg(row + j + 96);
                const double v4 = __ldg(row + j + 128), v5 = __ldg(row + j + 160), v6 = __ldg(row + j + 192), v7 = __ldg(row + j + 224);
                s0 = fma(v0, x[j], s0);
                s1 = fma(v1, x[j + 32], s1);
                s2 = fma(v2, x[j + 64], s2);
                s3 = fma(v3, x[j + 96], s3);
                s0 = fma(v4, x[j + 128], s0);
                s1 = fma(v5, x[j + 160], s1);
                s2 = fma(v6, x[j + 192], s2);
                s3 = fma(v7, x[j + 224], s3);
            }
            for (; j < na; j += 32) s0 = fma(__ldg(row + j), x[j], s0);
            double sv = (s0 + s1) + (s2 + s3);
            for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
            if (lane == 0) part[warp] = sv;
        }
        __syncthreads();
        if (warp == 0) {
            double xi = (lane < nbk) ? x[j0 + lane] - part[lane] : 0.0;
            const double rd = 1.0 / blk[lane * 33 + lane];  // reciprocal pivots: no division on the serial chain
            for (int c = nbk - 1; c >= 0; --c) {
                const double xc = __shfl_sync(0xffffffffu, xi * rd, c);
                if (lane == c) xi = xc;
                if (lane < c) xi = fma(-blk[lane * 33 + c], xc, xi);
            }
            if (lane < nbk) x[j0 + lane] = xi;
        }
        __syncthreads();
    }
    double* out = xout + (size_t)(xmap ? xmap[b] : b) * xstride_b + (size_t)rhs * xstride_r;
    for (int i = tid; i < na; i += blockDim.x) out[i] = x[i];
}

// Factorise `batch` matrices in place (see the header); `extra` right-hand-side rows at row xrow of every matrix.
// (A look-ahead variant -- the trailing update of block column j cut in two, the tiles of the next block column on
// the main stream and the rest on a side stream underneath the next diagonal block and panel -- was measured and
// gains nothing: 1.895 vs 1.912 ms at k = 2145.  Every step is ONE wave of latency-bound tiles either way, so the
// part that stays on the main stream takes as long as the whole update.)
int chol128_batched(ddb_ctx* ctx, double* base, size_t stride, int ld, const int* d_na, int na_max, int batch, int* d_fail,
                    int extra, int xrow, int* launches) {
    if (na_max <= 0 || batch <= 0) return DDB_OK;
    CholBatch cb{base, stride, ld, d_na, na_max, d_fail, extra, xrow, nullptr};
    const size_t dsmem = (size_t)kCB * kCP * sizeof(double);
    const size_t psmem = (size_t)(kCB + 32) * kCP * sizeof(double);
    const size_t gsmem = 2 * (size_t)kCStage * sizeof(double);
    DDB_CUDA_TRY(cudaFuncSetAttribute(chol_diag128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
    DDB_CUDA_TRY(cudaFuncSetAttribute(chol_panel128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    DDB_CUDA_TRY(cudaFuncSetAttribute(chol_syrk128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
    const size_t g64smem = 2 * (size_t)kSStage * sizeof(double);
    DDB_CUDA_TRY(cudaFuncSetAttribute(chol_syrk64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g64smem));
    cudaStream_t st = ctx->stream;
    int n = 0;
    for (int j0 = 0; j0 < na_max; j0 += kCB) {
        chol_diag128_kernel<<<batch, 256, dsmem, st>>>(cb, j0);
        ++n;
        const int rem = std::max(na_max - j0 - kCB, 0);  // matrix rows below this block column (largest matrix)
        if (rem + extra <= 0) break;
        const int Tr = (rem + extra + kCB - 1) / kCB, Tc = (rem + kCB - 1) / kCB;
        chol_panel128_kernel<<<dim3(Tr, batch), 256, psmem, st>>>(cb, j0);
        ++n;
        if (Tc > 0) {
            // lower tiles of 128 x 128 in this launch; fewer than two per SM: quarter-size tiles (see chol_syrk64_kernel)
            const long long tiles128 = (long long)batch * (Tc * (Tc + 1) / 2 + (Tr - Tc) * Tc);
            if (tiles128 < 2LL * ctx->sm_count) {
                const int Tr64 = (rem + extra + kSB - 1) / kSB, Tc64 = (rem + kSB - 1) / kSB;
                chol_syrk64_kernel<<<dim3(Tc64, Tr64, batch), 128, g64smem, st>>>(cb, j0);
            } else {
                chol_syrk128_kernel<<<dim3(Tc, Tr, batch), 256, gsmem, st>>>(cb, j0);
            }
            ++n;
        }
    }
    DDB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += n;
    return DDB_OK;
}

// x = L^-T y for `nrhs` right-hand sides per matrix (see chol_back_kernel).
int chol_back_batched(ddb_ctx* ctx, double* base, size_t stride, int ld, const int* d_na, int na_max, int batch, int yrow0,
                      int nrhs, double* xout, size_t xstride_b, size_t xstride_r, const int* xmap, int* launches) {
    if (na_max <= 0 || batch <= 0 || nrhs <= 0) return DDB_OK;
    CholBatch cb{base, stride, ld, d_na, na_max, nullptr, 0, 0, nullptr};
    const size_t smem = ((size_t)((na_max + 31) / 32) * 32 + 32 * 33) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("triangular solver: order %d exceeds the shared-memory right-hand-side buffer", na_max);
        return DDB_UNSUPPORTED;
    }
    DDB_CUDA_TRY(cudaFuncSetAttribute(chol_back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chol_back_kernel<<<dim3(nrhs, batch), 1024, smem, ctx->stream>>>(cb, yrow0, xout, xstride_b, xstride_r, xmap);
    DDB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return DDB_OK;
}

}  // namespace ddb
