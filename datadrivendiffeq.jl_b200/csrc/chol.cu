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
//                             32-column panels -- a left-looking DMMA update of the panel, the 32 x 32 pivot block in
//                             the registers of one warp (lane = row; the column is broadcast through shared memory
//                             and the chain per column is reciprocal -> scale -> FMA, square roots after the loop),
//                             the rows below it by substitution (one thread per row);
//   2. chol_panel128_kernel / chol_panel32_kernel   rows below, 128 or 32 per CTA:  X L_jj' = A[rows, j]  by blocked
//                             substitution -- for each of the four 32-column sub-blocks a DMMA product with the
//                             sub-blocks already solved, then a 32-step substitution per row.  No explicit inverse of
//                             the diagonal block is used (an inverse-based panel is not backward stable; the README
//                             Lorenz system, condition 1e11 after equilibration, shows it).  The result is also written
//                             TRANSPOSED into the unused upper triangle, so the buffer holds L below and L' above the
//                             diagonal (chol_solve_many of trsm.cu reads contiguous rows of L');
//   3. chol_gemm_kernel<SYRK> A[rest, rest] -= L[rest, j] L[rest, j]'  (DMMA, lower 128 x 128 tiles only).
// Right-hand sides appended as EXTRA ROWS below the matrix (row index `xrow` onwards, same leading dimension) pass
// through steps 2 and 3 like any other row: when the factorisation ends they hold y = L^-1 b (the last rows of the
// Cholesky factor of the bordered matrix [[A, b], [b', .]] are b' L^-T).  chol_back_kernel finishes with x = L^-T y.
// These are latency-bound kernels on one or a few SMs; tools/lat_probe.cu holds the single-warp numbers they were
// tuned with (dependent DFMA 8.8 cycles, 64-bit shuffle 27 / 8.4 issue, DMMA 26 / 16.5 issue, reciprocal chain 39) and
// tools/chol_bench.cu prints the cycles of every phase.  k = 2145 with 64 right-hand sides: factor 1.85 -> 1.15 ms,
// back substitution 0.48 -> 0.37 ms.
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
// reciprocal diagonal in rinv), right-looking in registers, 496 FMAs; row[c * xstride] is entry c of the row.  The
// entry of L a step needs is requested one column ahead, right after the register it replaces was used: left to itself
// ptxas keeps ~8 of the 496 broadcast loads in flight and the 30-cycle shared-memory latency shows ~60 times (3.0
// thousand cycles per block for 1.1 thousand cycles of FP64 issue; 2.4 with the explicit order, tools/chol_bench.cu).
// (Running the recurrence on values pre-scaled by the pivots -- one dependent FMA per column instead of a multiply
// and an FMA -- was measured and is slower: the block is bound by issue and load latency, not by the chain.)
__device__ __forceinline__ void row_subst32_pf(double* row, int xstride, const double* ldp, int pitch, const double* rinv) {
    double x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = row[c * xstride];
    double lc[32];  // column c of L below the diagonal; every entry is replaced by the one of column c + 1 right after use
#pragma unroll
    for (int j = 1; j < 32; ++j) lc[j] = ldp[j * pitch];
    double rc = rinv[0];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        x[c] *= rc;
        if (c + 1 < 32) rc = rinv[c + 1];
        if (c + 1 < 32) x[c + 1] = fma(-x[c], lc[c + 1], x[c + 1]);
#pragma unroll
        for (int j = c + 2; j < 32; ++j) {
            x[j] = fma(-x[c], lc[j], x[j]);
            lc[j] = ldp[j * pitch + c + 1];
        }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) row[c * xstride] = x[c];
}

__global__ void __launch_bounds__(256, 1) chol_diag128_kernel(CholBatch cb, int j0) {
    const int b = blockIdx.x;
    const int na = cb.na_arr ? cb.na_arr[b] : cb.na_all;
    if (j0 >= na) return;
    const int nb = min(kCB, na - j0);
    extern __shared__ double dsm[];
    double* S = dsm;  // 128 x 132
    __shared__ double rinv[32];
    __shared__ __align__(16) double colb[2 * 32];  // the current column of the pivot block, broadcast to every lane
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
        // ---- 0. LEFT-looking update of this 32-column panel (rows c0 .. 127) with everything solved so far:
        //   S[i][c0 + j] -= sum_{t < c0} S[i][t] S[c0 + j][t].  A warp owns column sub-tile warp % 4 of the row
        //   sub-tiles warp / 4 + 2 v: one B fragment per k-step, up to six independent accumulator chains (a DMMA is
        //   26 cycles of latency and 16.5 of issue, tools/lat_probe.cu; the right-looking form of this update ran one
        //   chain of eight per sub-tile and cost 11 of the kernel's 45 thousand cycles).
        if (c0 > 0) {
            const int j8 = warp & 3, ib = warp >> 2, nv = (kCB - c0) / 16;  // 6, 4, 2 row sub-tiles per warp
            double u[6][2];
#pragma unroll
            for (int v = 0; v < 6; ++v) u[v][0] = u[v][1] = 0.0;
            const double* pa = S + (c0 + 8 * ib + g) * kCP + t4;
            const double* pb = S + (c0 + 8 * j8 + g) * kCP + t4;
#pragma unroll 1
            for (int kk = 0; kk < c0; kk += 8) {  // two k-steps of fragments requested at once
                double af[2][6], bf[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    bf[q] = pb[kk + 4 * q];
#pragma unroll
                    for (int v = 0; v < 6; ++v)
                        if (v < nv) af[q][v] = pa[16 * v * kCP + kk + 4 * q];
                }
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int v = 0; v < 6; ++v)
                        if (v < nv) dmma884(u[v][0], u[v][1], af[q][v], bf[q]);
            }
#pragma unroll
            for (int v = 0; v < 6; ++v) {
                if (v >= nv) continue;
                const int r = 8 * (ib + 2 * v) + g, c = 8 * j8 + 2 * t4;  // relative to (c0, c0); upper part of the pivot block dropped
                double* o = S + (c0 + r) * kCP + c0 + c;
                if (c <= r) o[0] -= u[v][0];
                if (c + 1 <= r) o[1] -= u[v][1];
            }
            __syncthreads();
            if (tracing) { const long long t = clock64(); tr_upd += t - tr_mark; tr_mark = t; }
        }
        // ---- 1. the 32 x 32 pivot block: lane = row, the row lives in registers, columns one after the other.
        // The serial chain per column is  d -> 1/d -> w = a/d -> a_j -= w a_jc  (hardware reciprocal estimate + five
        // dependent FP64 operations): the update uses the UNNORMALISED column (L_jc L_ic = a_jc a_ic / d), whose
        // broadcasts do not wait for the reciprocal; the square root that turns the column into L (and 1 / L_cc) is
        // taken after the loop.  Was: d -> rsqrt -> l = a ri -> broadcast l by 64-bit shuffles -> update, 443 cycles per
        // column; now 170.
        if (warp == 0) {
            double a[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] = S[(c0 + lane) * kCP + c0 + j];
            int bad = 0;
            double my_rinv = 1.0;
            // The column travels through shared memory (one store per lane, 128-bit broadcast loads): a 64-bit shuffle
            // is two instructions and 8.4 cycles of the shuffle unit per value (tools/lat_probe.cu).  The loop body is
            // branch-free -- a pivot outside (kPivotTol, 1e200) is replaced by 1 and flagged, so the reciprocal needs no
            // slow path -- and the square roots are taken AFTER the loop, one column per lane: the warp issues in
            // order, so "independent" work between two columns still delays the chain by its own latency.
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                double* cbuf = colb + (c & 1) * 32;
                cbuf[lane] = a[c];
                __syncwarp();
                double sh[32];
#pragma unroll
                for (int p2 = (c >> 1); p2 < 16; ++p2) {
                    const double2 v = *reinterpret_cast<const double2*>(cbuf + 2 * p2);
                    sh[2 * p2] = v.x, sh[2 * p2 + 1] = v.y;
                }
                double d = sh[c];
                if (!(d > kPivotTol && d < 1e200)) {  // unit-diagonal (equilibrated) matrices: numerically rank deficient
                    bad = 1;
                    d = 1.0;  // keep going with a harmless pivot; the flag invalidates the result
                }
                double y;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));  // ~20 bits; one third-order step: e^3 ~ 2^-60
                const double e = fma(-d, y, 1.0);
                const double e2 = fma(e, e, e);
                const double w = a[c] * fma(y, e2, y);
#pragma unroll
                for (int j = c + 1; j < 32; ++j) a[j] = fma(-w, sh[j], a[j]);
                if (lane == c) my_rinv = d;  // the pivot of column `lane`
            }
            // L = (unnormalised columns) / sqrt(pivot): lane c takes the root of pivot c, everybody scales
            my_rinv = fast_rsqrt(my_rinv);
            colb[lane] = my_rinv;
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 32; ++c) a[c] *= colb[c];  // (lane == c: d / sqrt(d))
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
            if (tid < Rr) row_subst32_pf(S + (t0 + tid) * kCP + c0, 1, S + c0 * kCP + c0, kCP, rinv);
            __syncthreads();
            if (tracing) { const long long t = clock64(); tr_sub += t - tr_mark; tr_mark = t; }
        }
    }
    if (tracing) cb.trace[1] += tr_piv, cb.trace[2] += tr_sub, cb.trace[3] += tr_upd;
    if (tid == 0 && s_bad) atomicExch(&cb.fail[b], 1);
    // the factor goes back in place: lower triangle of the block, and mirrored above the diagonal (L' for the
    // backward substitution, like the panel step does for the rows below)
    for (int e = tid; e < kCB * kCB; e += 256) {
        const int i = e >> 7, j = e & 127;
        if (i < nb && j <= i) A[(size_t)(j0 + i) * ld + j0 + j] = S[i * kCP + j];
    }
    // L'[i][j] = L[j][i] for j > i in patches of 4 rows (i) x 8 columns (j) per warp instruction: the transposed reads
    // S[j][i] hit 16 distinct 8-byte banks per half-warp (a row of 32 consecutive j is an 8-way conflict at this pitch)
    for (int pch = warp; pch < 32 * 16; pch += 8) {
        const int ib = 4 * (pch & 31), jb = 8 * (pch >> 5);
        if (jb + 7 <= ib) continue;  // entirely on or below the diagonal
        const int i = ib + (lane & 3), j = jb + (lane >> 2);
        if (j > i && j < nb) A[(size_t)(j0 + i) * ld + j0 + j] = S[j * kCP + i];
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
    double* Lqb = psm + kCB * kCP;    // 2 x (32 x 132): block row q of L_jj (columns 0 .. 32 q + 31), double-buffered
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
    auto fetch_lq = [&](int cq, double* Lq) {  // block row cq / 32 of L_jj: 32 x (cq + 32), identity beyond nb
        const int nbq = min(32, nb - cq), w = cq + 32;
        for (int i = warp; i < 32; i += 8)
            for (int j = lane; j < w; j += 32) {
                if (i < nbq && j <= cq + i) cp8c(Lq + i * kCP + j, A + (size_t)(j0 + cq + i) * ld + j0 + j);
                else Lq[i * kCP + j] = (j == cq + i ? 1.0 : 0.0);
            }
    };
    fetch_lq(0, Lqb);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const bool tracing = cb.trace && b == 0 && blockIdx.x == 0 && tid == 0;
    long long tr_mark = 0, tr_wait = 0, tr_prod = 0, tr_sub = 0;
    if (tracing) tr_mark = clock64();
#define DDB_TR(acc) if (tracing) { const long long t_ = clock64(); acc += t_ - tr_mark; tr_mark = t_; }
#pragma unroll 1
    for (int cq = 0; cq < nb; cq += 32) {
        double* Lq = Lqb + ((cq >> 5) & 1) * 32 * kCP;
        // the next block row travels underneath this one (its buffer was last read two sub-blocks ago, and every thread
        // has passed the barrier that closed that sub-block)
        if (cq + 32 < nb) fetch_lq(cq + 32, Lqb + (((cq >> 5) + 1) & 1) * 32 * kCP);
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;" ::: "memory");  // everything but the prefetch
        __syncthreads();
        DDB_TR(tr_wait)
        if (tid < 32) rinv[tid] = 1.0 / Lq[tid * kCP + cq + tid];
        if (cq > 0) {
            // At[:, cq .. cq + 31] -= At[:, 0 .. cq) Lq[:, 0 .. cq)'   (16 x 4 sub-tiles of 8 x 8; a warp owns column
            // sub-tile warp % 4 of the eight row sub-tiles warp / 4 + 2 s: one B fragment per k-step, eight independent
            // accumulator chains)
            const int j8 = warp & 3, ib = warp >> 2;
            double u[8][2];
#pragma unroll
            for (int v = 0; v < 8; ++v) u[v][0] = u[v][1] = 0.0;
            const double* pa = At + (8 * ib + g) * kCP + t4;
            const double* pb = Lq + (8 * j8 + g) * kCP + t4;
#pragma unroll 2
            for (int kk = 0; kk < cq; kk += 4) {
                const double bf = pb[kk];
#pragma unroll
                for (int v = 0; v < 8; ++v) dmma884(u[v][0], u[v][1], pa[16 * v * kCP + kk], bf);
            }
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                double* o = At + (8 * (ib + 2 * v) + g) * kCP + cq + 8 * j8 + 2 * t4;
                o[0] -= u[v][0];
                o[1] -= u[v][1];
            }
        }
        __syncthreads();
        DDB_TR(tr_prod)
        if (tid < kCB) row_subst32_pf(At + tid * kCP + cq, 1, Lq + cq, kCP, rinv);
        __syncthreads();  // closes the sub-block: Lq and rinv may be rewritten
        DDB_TR(tr_sub)
    }
    if (tracing) cb.trace[8] += tr_wait, cb.trace[9] += tr_prod, cb.trace[10] += tr_sub;
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
    if (tracing) cb.trace[11] += clock64() - tr_mark, cb.trace[12] += 1;
#undef DDB_TR
}

// The same panel step with 32 rows per CTA and the tile held COLUMN-major in shared memory, for launches that would
// not fill the machine with 128-row CTAs (a single matrix of order 2145: 17 CTAs whose DMMA products -- 1.6 Mflop on
// one SM -- and four 32-column substitutions took 22 us of the 31 us step; tools/chol_bench.cu prints the phases).
// T[c][r] with pitch 36 = 4 (mod 16): the DMMA A fragments (row g, contraction index t4) sit at (kk + t4) * 36 + g --
// 16 distinct 8-byte banks per half-warp -- AND the substitution (lane = row, one column after the other) and the
// mirrored store read consecutive addresses; the row-major layout of chol_panel128_kernel costs an 8-way bank conflict
// on every one of those.  Global <-> shared copies move 4-row x 8-column patches per warp instruction (conflict-free
// on the shared side, 64 contiguous bytes per row on the global side).  grid (ceil(rows_all / 32), batch), 128 threads.
constexpr int kPR = 32;   // rows per CTA
constexpr int kPT = 36;   // pitch of the column-major tile
__global__ void __launch_bounds__(128) chol_panel32_kernel(CholBatch cb, int j0) {
    const int b = blockIdx.y;
    const int na = cb.na_arr ? cb.na_arr[b] : cb.na_all;
    if (j0 >= na) return;
    const int nb = min(kCB, na - j0);
    const int t0 = j0 + nb, rows = na - t0, rows_all = rows + cb.extra;
    const int r0 = blockIdx.x * kPR;
    if (r0 >= rows_all) return;
    extern __shared__ double psm[];
    double* T = psm;                  // 128 x 36, T[c][r]
    double* Lqb = psm + kCB * kPT;    // 2 x (32 x 132): block row q of L_jj (columns 0 .. 32 q + 31), double-buffered
    __shared__ double rinv[32];
    double* A = cb.base + (size_t)b * cb.stride;
    const int ld = cb.ld;
    const int xoff = cb.xrow - t0 - rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    // patch p = (row group p % 8, column group p / 8): rows 4 (p % 8) + lane % 4, columns 8 (p / 8) + lane / 4
    for (int pch = warp; pch < 128; pch += 4) {
        const int r = 4 * (pch & 7) + (lane & 3), c = 8 * (pch >> 3) + (lane >> 2), rr = r0 + r;
        if (rr < rows_all && c < nb) cp8c(T + c * kPT + r, A + (size_t)(t0 + rr + (rr >= rows ? xoff : 0)) * ld + j0 + c);
        else T[c * kPT + r] = 0.0;
    }
    auto fetch_lq = [&](int cq, double* Lq) {  // block row cq / 32 of L_jj: 32 x (cq + 32), identity beyond nb
        const int nbq = min(32, nb - cq), w = cq + 32;
        for (int i = warp; i < 32; i += 4)
            for (int j = lane; j < w; j += 32) {
                if (i < nbq && j <= cq + i) cp8c(Lq + i * kCP + j, A + (size_t)(j0 + cq + i) * ld + j0 + j);
                else Lq[i * kCP + j] = (j == cq + i ? 1.0 : 0.0);
            }
    };
    fetch_lq(0, Lqb);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const bool tracing = cb.trace && b == 0 && blockIdx.x == 0 && tid == 0;
    long long tr_mark = 0, tr_wait = 0, tr_prod = 0, tr_sub = 0;
    if (tracing) tr_mark = clock64();
#define DDB_TR(acc) if (tracing) { const long long t_ = clock64(); acc += t_ - tr_mark; tr_mark = t_; }
#pragma unroll 1
    for (int cq = 0; cq < nb; cq += 32) {
        double* Lq = Lqb + ((cq >> 5) & 1) * 32 * kCP;
        // the next block row travels underneath this one (its buffer was last read two sub-blocks ago, and every thread
        // has passed the barrier that closed that sub-block)
        if (cq + 32 < nb) fetch_lq(cq + 32, Lqb + (((cq >> 5) + 1) & 1) * 32 * kCP);
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;" ::: "memory");  // everything but the prefetch
        __syncthreads();
        DDB_TR(tr_wait)
        if (tid < 32) rinv[tid] = 1.0 / Lq[tid * kCP + cq + tid];
        if (cq > 0) {
            // T[cq .. cq + 31][:] -= Lq[:, 0 .. cq) T[0 .. cq)[:]   (4 x 4 sub-tiles of 8 x 8; warp w owns column
            // sub-tile w and all four row sub-tiles: one B fragment per k-step, four independent accumulator chains)
            double u[4][2];
#pragma unroll
            for (int v = 0; v < 4; ++v) u[v][0] = u[v][1] = 0.0;
            const double* pa = T + t4 * kPT + g;
            const double* pb = Lq + (8 * warp + g) * kCP + t4;
#pragma unroll 1
            for (int kk = 0; kk < cq; kk += 16) {  // cq is a multiple of 32: four k-steps of fragments requested at once
                double af[4][4], bf[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    bf[q] = pb[kk + 4 * q];
#pragma unroll
                    for (int v = 0; v < 4; ++v) af[q][v] = pa[(kk + 4 * q) * kPT + 8 * v];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int v = 0; v < 4; ++v) dmma884(u[v][0], u[v][1], af[q][v], bf[q]);
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                double* o = T + (cq + 8 * warp + 2 * t4) * kPT + 8 * v + g;
                o[0] -= u[v][0];
                o[kPT] -= u[v][1];
            }
        }
        __syncthreads();
        DDB_TR(tr_prod)
        // one row per lane: X L_qq' = T[cq .. cq + 31][row], right-looking in registers
        if (tid < kPR) row_subst32_pf(T + cq * kPT + tid, kPT, Lq + cq, kCP, rinv);
        __syncthreads();  // closes the sub-block: Lq and rinv may be rewritten
        DDB_TR(tr_sub)
    }
    if (tracing) cb.trace[8] += tr_wait, cb.trace[9] += tr_prod, cb.trace[10] += tr_sub;
    for (int pch = warp; pch < 128; pch += 4) {
        const int r = 4 * (pch & 7) + (lane & 3), c = 8 * (pch >> 3) + (lane >> 2), rr = r0 + r;
        if (rr < rows_all && c < nb) A[(size_t)(t0 + rr + (rr >= rows ? xoff : 0)) * ld + j0 + c] = T[c * kPT + r];
    }
    // L' above the diagonal: A[j0 + c][t0 + rr] for the matrix rows (coalesced along rr, consecutive in T)
    for (int c = warp; c < nb; c += 4) {
        const int rr = r0 + lane;
        if (rr < rows) A[(size_t)(j0 + c) * ld + t0 + rr] = T[c * kPT + lane];
    }
    if (tracing) cb.trace[11] += clock64() - tr_mark, cb.trace[12] += 1;
#undef DDB_TR
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

// Backward substitution x = L^-T y, one CTA per (right-hand side, matrix), RIGHT-looking over 32-row blocks from the
// bottom, software-pipelined so that only the 32-step chain of a diagonal block is serial:
//   step s (block rows j0 .. j0 + 31):
//     warp 0        r_s -= N_s x_(s-1)   (the 32 x 32 block of L' that couples block s to the block solved last,
//                   prefetched into shared memory during the previous step), then the substitution against the diagonal
//                   block.  The block is prefetched PRE-SCALED, M[i][c] = L[j0 + c][j0 + i] / L[j0 + c][j0 + c], so the
//                   chain per column is one shuffle and one FMA on the unnormalised value (x_c = r_c / L_cc off chain);
//     warps 1 .. 31 r[i] -= sum_c L[jp + c][i] x_(s-1)[c] for every row i above block s (jp = first row of block
//                   s - 1): one thread per row, reads along the ROWS of L (lower triangle, coalesced over i), and the
//                   prefetch of the blocks step s + 1 needs.
// The previous version was left-looking (dot products along the rows of L' for the 32 rows of the block, then the
// chain): 68 x (L2 round trips + chain) at k = 2145, 0.47 ms for 64 right-hand sides; this one 0.47 -> see DESIGN 4.2.
// y: row `yrow0 + rhs` of the matrix buffer (where the factorisation left L^-1 b); x -> xout + (xmap ? xmap[b] : b) *
// xstride_b + rhs * xstride_r.
constexpr int kBackThreads = 1024;
constexpr int kBackWorkers = kBackThreads - 32;
constexpr int kBP = 33;  // pitch of the 32 x 32 blocks in shared memory
__global__ void __launch_bounds__(kBackThreads) chol_back_kernel(CholBatch cb, int yrow0, double* xout, size_t xstride_b,
                                                                size_t xstride_r, const int* xmap) {
    const int b = blockIdx.y, rhs = blockIdx.x;
    const int na = cb.na_arr ? cb.na_arr[b] : cb.na_all;
    if (na <= 0) return;
    const double* A = cb.base + (size_t)b * cb.stride;
    const int ld = cb.ld;
    extern __shared__ double bsm[];
    const int na32 = ((na + 31) / 32) * 32;
    double* x = bsm;                     // na32: right-hand side, overwritten block by block with the solution
    double* Dm = bsm + na32;             // [2][32 x 33] pre-scaled diagonal blocks, Dm[i][c] = L[j0+c][j0+i] / L[j0+c][j0+c] (i < c)
    double* Nn = Dm + 2 * 32 * kBP;      // [2][32 x 33] coupling blocks, Nn[i][c] = L[jp + c][j0 + i]
    double* rdv = Nn + 2 * 32 * kBP;     // [2][32] reciprocal pivots
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* y = A + (size_t)(yrow0 + rhs) * ld;
    for (int i = tid; i < na32; i += kBackThreads) x[i] = (i < na) ? y[i] : 0.0;

    // blocks of step `s`: diagonal block at j0, coupling to the block below it (rows jp .. jp + nbp - 1 of L)
    auto fetch_blocks = [&](int e, int j0, int nbk, int jp, int nbp, int buf) {
        const int c = e >> 5, i = e & 31;
        double v = 0.0, d = 1.0, w = 0.0;
        if (c < nbk) {
            d = A[(size_t)(j0 + c) * ld + j0 + c];
            if (i < c) v = A[(size_t)(j0 + c) * ld + j0 + i];
        }
        if (c < nbp && i < nbk) w = A[(size_t)(jp + c) * ld + j0 + i];
        const double rd = 1.0 / d;
        Dm[buf * 32 * kBP + i * kBP + c] = v * rd;
        Nn[buf * 32 * kBP + i * kBP + c] = w;
        if (i == c) rdv[buf * 32 + c] = rd;
    };

    const int jlast = ((na - 1) / 32) * 32;
    fetch_blocks(tid, jlast, na - jlast, 0, 0, 0);
    __syncthreads();
    int buf = 0;
    for (int j0 = jlast; j0 >= 0; j0 -= 32, buf ^= 1) {
        const int nbk = min(32, na - j0);
        const int jp = j0 + 32, nbp = (j0 == jlast) ? 0 : min(32, na - jp);  // the block solved in the previous step
        if (warp == 0) {
            const double* mrow = Dm + buf * 32 * kBP + lane * kBP;
            double xi = x[j0 + lane];
            if (nbp > 0) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                const double* nn = Nn + buf * 32 * kBP + lane * kBP;
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    s0 = fma(nn[c], x[jp + c], s0);
                    s1 = fma(nn[c + 1], x[jp + c + 1], s1);
                    s2 = fma(nn[c + 2], x[jp + c + 2], s2);
                    s3 = fma(nn[c + 3], x[jp + c + 3], s3);
                }
                xi -= (s0 + s1) + (s2 + s3);
            }
            // (64 registers per thread at 1024 threads: the scaled row is read in chunks of eight ahead of the chain)
#pragma unroll
            for (int c8 = 24; c8 >= 0; c8 -= 8) {
                double m8[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) m8[q] = mrow[c8 + q];
#pragma unroll
                for (int q = 7; q >= 0; --q) {
                    if (c8 + q == 0) continue;
                    const double xc = __shfl_sync(0xffffffffu, xi, c8 + q);
                    xi = fma(-m8[q], xc, xi);  // m = 0 for lane >= c and beyond a ragged block
                }
            }
            // x of this block is published after the barrier below (the workers still read x of the previous block and
            // write r above this block, never this block)
            x[j0 + lane] = xi * rdv[buf * 32 + lane];
        } else {
            const int w = tid - 32;
            if (j0 >= 32 && w < 1024) {  // blocks of the next step (j0 - 32; always a full block)
                fetch_blocks(w, j0 - 32, 32, j0, nbk, buf ^ 1);
                if (w + kBackWorkers < 1024) fetch_blocks(w + kBackWorkers, j0 - 32, 32, j0, nbk, buf ^ 1);
            }
            if (nbp > 0) {
                const double* Lp = A + (size_t)jp * ld;
                for (int i = w; i < j0; i += kBackWorkers) {
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                    if (nbp == 32) {
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            s0 = fma(Lp[(size_t)c * ld + i], x[jp + c], s0);
                            s1 = fma(Lp[(size_t)(c + 1) * ld + i], x[jp + c + 1], s1);
                            s2 = fma(Lp[(size_t)(c + 2) * ld + i], x[jp + c + 2], s2);
                            s3 = fma(Lp[(size_t)(c + 3) * ld + i], x[jp + c + 3], s3);
                        }
                    } else {
                        for (int c = 0; c < nbp; ++c) s0 = fma(Lp[(size_t)c * ld + i], x[jp + c], s0);
                    }
                    x[i] -= (s0 + s1) + (s2 + s3);
                }
            }
        }
        __syncthreads();
    }
    double* out = xout + (size_t)(xmap ? xmap[b] : b) * xstride_b + (size_t)rhs * xstride_r;
    for (int i = tid; i < na; i += kBackThreads) out[i] = x[i];
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
    const size_t psmem = (size_t)(kCB + 64) * kCP * sizeof(double);
    const size_t gsmem = 2 * (size_t)kCStage * sizeof(double);
    DDB_CUDA_TRY(cudaFuncSetAttribute(chol_diag128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
    DDB_CUDA_TRY(cudaFuncSetAttribute(chol_panel128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    const size_t p32smem = ((size_t)kCB * kPT + 2 * 32 * kCP) * sizeof(double);
    DDB_CUDA_TRY(cudaFuncSetAttribute(chol_panel32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p32smem));
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
        if ((long long)Tr * batch < ctx->sm_count)  // too few 128-row CTAs to fill the machine: 32 rows per CTA
            chol_panel32_kernel<<<dim3((rem + extra + kPR - 1) / kPR, batch), 128, p32smem, st>>>(cb, j0);
        else
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
    const size_t smem = ((size_t)((na_max + 31) / 32) * 32 + 4 * 32 * kBP + 64) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("triangular solver: order %d exceeds the shared-memory right-hand-side buffer", na_max);
        return DDB_UNSUPPORTED;
    }
    DDB_CUDA_TRY(cudaFuncSetAttribute(chol_back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chol_back_kernel<<<dim3(nrhs, batch), kBackThreads, smem, ctx->stream>>>(cb, yrow0, xout, xstride_b, xstride_r, xmap);
    DDB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return DDB_OK;
}

}  // namespace ddb
