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
#include "mma_common.cuh"
#include "sweep.cuh"
#include <cmath>
#include <cstdlib>
#include <string>
#include <unordered_map>

namespace ddb {

// One sparse term of a candidate, decoded once per pass: the coefficient and the shared-memory ROWS of its
// factors (states 0..n-1; row n + n_t holds ones, so a missing factor multiplies by 1.0 and the product
// needs no predicate).  o32 = first two factor rows pre-multiplied by the tile pitch (degree <= 2 path reads
// only the first 16 bytes), o16 = all kMaxFactors rows.
struct __align__(16) RssTerm {
    double val;
    int32_t o32[2];
    uint16_t o16[kMaxFactors];
};
static_assert(sizeof(RssTerm) == 32, "RssTerm layout");

constexpr int kRssTile = 32;   // samples per shared-memory tile (one per lane)
constexpr int kRssAcc = 32;    // candidates per warp per pass (register accumulators per lane)
constexpr int kRssWarps = 8;
// A staged tile holds `sub` groups of kRssTile samples (sub = 1 for wide problems, up to 8 when a sample is only
// a few doubles, so that a tile is ~16 KB either way and the per-tile synchronisation is amortised); it is
// stored TRANSPOSED [state][sample] with the odd pitch 32 sub + 1 -> "lane = sample" reads are conflict-free.
constexpr int kRssMaxSub = 8;

// `map` (implicit mode) translates indices of the selected sub-library into rows of the installed library;
// cand_val == nullptr writes coefficient 1 (the records of the implicit targets).
__global__ void rss_expand_kernel(const int32_t* __restrict__ cand_idx, const double* __restrict__ cand_val,
                                  int64_t nterms, const uint16_t* __restrict__ factors, int ones_row, int pitch,
                                  const int32_t* __restrict__ map, RssTerm* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nterms) return;
    RssTerm r;
    r.val = cand_val ? cand_val[t] : 1.0;
    const int term = cand_idx ? cand_idx[t] : (int)t;
    const uint16_t* f = factors + (size_t)(map ? map[term] : term) * kMaxFactors;
#pragma unroll
    for (int d = 0; d < kMaxFactors; ++d) r.o16[d] = (f[d] == kNoFactor) ? (uint16_t)ones_row : f[d];
    r.o32[0] = (int)r.o16[0] * pitch;
    r.o32[1] = (int)r.o16[1] * pitch;
    out[t] = r;
}

struct RssArgs {
    const double* X;
    const double* Y;
    int n, n_t;
    int64_t m;
    int ncand;
    const int64_t* cand_off;
    const int32_t* cand_nnz;
    const int32_t* cand_eq;
    const RssTerm* terms;
    const RssTerm* targets;  // implicit mode: target e is a library term (one record each); null = row e of Y
    const double* cand_shift;  // per candidate: constant added to its fit (affine term normalisation) or null
    int term_cap;            // term records of one CTA's candidates that fit in shared memory
    int sub, pitch;          // groups of 32 samples per staged tile; row pitch of the tile = 32 sub + 1
    double* partials;  // gridDim.x x ncand
};

__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}

// Stage one tile of kRssTile samples TRANSPOSED into shared memory ([row][sample], rows = n states then
// n_t targets) with 8-byte asynchronous copies: global reads stay coalesced (a tile of X is one contiguous
// run of 32 n doubles), the transposition makes "lane = sample" accesses conflict-free.  Element e of the
// run goes to row e % w, sample e / w; the (row, sample) pair is advanced incrementally (no division).
struct RssWalk {
    int i0, s0, di, ds;  // element threadIdx.x -> (row i0, sample s0); + blockDim.x elements -> (+di, +ds)
    __device__ __forceinline__ void init(int w) {
        i0 = threadIdx.x % w;
        s0 = threadIdx.x / w;
        di = blockDim.x % w;
        ds = blockDim.x / w;
    }
};
__device__ __forceinline__ void rss_stage_rows(double* buf, const double* src, int w, int valid, int pitch,
                                               const RssWalk& k) {
    int i = k.i0, s = k.s0;
    const int total = valid * w;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        cp_async8(buf + i * pitch + s, src + e);
        i += k.di;
        s += k.ds;
        if (i >= w) {
            i -= w;
            ++s;
        }
    }
}

// The tile loop.  SMEM_TERMS: the candidates' term records live in shared memory (s_terms); otherwise they are
// read from the global table.  Every shared-memory address is formed from `sm` directly so that the compiler
// keeps the accesses in the shared address space (LDS with 32-bit addressing, no generic-pointer arithmetic).
template <int NF, bool SMEM_TERMS>
__device__ __forceinline__ void rss_tiles(const RssArgs& a, double* sm, const int4* s_meta, const RssTerm* s_terms,
                                          const double* s_shift, int nslots, int cbase, int64_t t_lo, int64_t t_hi, int tp,
                                          double (&acc)[kRssAcc]) {
    const int rows = a.n + a.n_t + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = a.pitch, tile_samples = kRssTile * a.sub;
    const int tile_doubles = rows * pitch;
    RssWalk wx, wy;
    wx.init(a.n);
    wy.init(a.n_t);
    auto stage = [&](int which, int64_t tile) {
        const int64_t s0 = tile * tile_samples;
        const int valid = (int)min((int64_t)tile_samples, a.m - s0);
        double* dst = sm + which * tile_doubles;
        rss_stage_rows(dst, a.X + s0 * a.n, a.n, valid, pitch, wx);
        rss_stage_rows(dst + a.n * pitch, a.Y + s0 * a.n_t, a.n_t, valid, pitch, wy);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (t_lo < t_hi) stage(0, t_lo);
    for (int64_t tile = t_lo; tile < t_hi; ++tile) {
        const int cur = (int)((tile - t_lo) & 1);
        if (tile + 1 < t_hi) {
            stage(cur ^ 1, tile + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int valid = (int)min((int64_t)tile_samples, a.m - tile * tile_samples);
#pragma unroll 1
        for (int ss = 0; ss < a.sub; ++ss) {
        const double* xs = sm + cur * tile_doubles + ss * kRssTile + lane;
        const bool live = ss * kRssTile + lane < valid;
        if (ss * kRssTile >= valid) break;
#pragma unroll
        for (int q = 0; q < kRssAcc; ++q) {
            if (q >= nslots) break;  // slots are filled in order: one exit instead of 32 guarded bodies
            {
                const int4 mt = s_meta[warp + kRssWarps * q];
                const int nnz = SMEM_TERMS ? mt.w : mt.y, eq = mt.z;  // shared-memory records are padded to 4s
                const RssTerm* r = SMEM_TERMS ? s_terms + mt.x : a.terms + __ldg(a.cand_off + cbase + warp + kRssWarps * q);
                const int4* r16 = reinterpret_cast<const int4*>(s_terms) + mt.x;  // compact records (degree <= 2)
                double fit = s_shift[warp + kRssWarps * q];
                if (NF <= 2 && SMEM_TERMS) {
                    for (int t = 0; t < nnz; t += 4) {  // the padded count is a multiple of 4
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            // {val, (o32[0], o32[1])} as two doubles: the coefficient lands in a register pair and
                            // the offsets are the halves of the second one -- no moves
                            const double2 w = reinterpret_cast<const double2*>(r16)[t + u];
                            fit = fma(w.x, xs[__double2loint(w.y)] * xs[__double2hiint(w.y)], fit);
                        }
                    }
                } else if (NF <= 2) {
                    for (int t = 0; t < nnz; ++t) {
                        const int4 w = __ldg(reinterpret_cast<const int4*>(r + t));
                        fit = fma(__hiloint2double(w.y, w.x), xs[w.z] * xs[w.w], fit);
                    }
                } else {
                    for (int t = 0; t < nnz; ++t) {
                        const int4 w = SMEM_TERMS ? *reinterpret_cast<const int4*>(r + t) : __ldg(reinterpret_cast<const int4*>(r + t));
                        const int4 o = SMEM_TERMS ? *(reinterpret_cast<const int4*>(r + t) + 1)
                                                  : __ldg(reinterpret_cast<const int4*>(r + t) + 1);  // 8 x uint16 rows
                        double p = xs[w.z] * xs[w.w];
                        const uint32_t ow[4] = {(uint32_t)o.x, (uint32_t)o.y, (uint32_t)o.z, (uint32_t)o.w};
#pragma unroll
                        for (int d = 2; d < NF; ++d) p *= xs[((ow[d >> 1] >> (16 * (d & 1))) & 0xFFFFu) * pitch];
                        fit = fma(__hiloint2double(w.y, w.x), p, fit);
                    }
                }
                double yv;
                if (a.targets) {
                    const int4 w = __ldg(reinterpret_cast<const int4*>(a.targets + eq));
                    yv = xs[w.z] * xs[w.w];
                    if (NF > 2) {
                        const int4 o = __ldg(reinterpret_cast<const int4*>(a.targets + eq) + 1);
                        const uint32_t ow[4] = {(uint32_t)o.x, (uint32_t)o.y, (uint32_t)o.z, (uint32_t)o.w};
#pragma unroll
                        for (int d = 2; d < NF; ++d) yv *= xs[((ow[d >> 1] >> (16 * (d & 1))) & 0xFFFFu) * pitch];
                    }
                } else {
                    yv = xs[(a.n + eq) * pitch];
                }
                const double res = live ? fit - yv : 0.0;
                acc[q] = fma(res, res, acc[q]);
            }
        }
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled two iterations later
    }
}

template <int NF>
__global__ void __launch_bounds__(kRssWarps * 32) rss_kernel(const RssArgs a) {
    extern __shared__ double sm[];
    const int rows = a.n + a.n_t + 1;  // + the row of ones
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cbase = blockIdx.y * (kRssWarps * kRssAcc);
    // this CTA's sample range (whole tiles)
    const int pitch = a.pitch;
    const int64_t ntiles = (a.m + kRssTile * a.sub - 1) / (kRssTile * a.sub);
    const int64_t t_lo = ntiles * blockIdx.x / gridDim.x, t_hi = ntiles * (blockIdx.x + 1) / gridDim.x;
    for (int e = threadIdx.x; e < 2 * pitch; e += blockDim.x)
        sm[(e / pitch) * rows * pitch + (rows - 1) * pitch + e % pitch] = 1.0;

    // candidates of this warp: cbase + warp + 8 * slot
    int nslots = 0;
    for (int q = 0; q < kRssAcc; ++q)
        if (cbase + warp + kRssWarps * q < a.ncand) nslots = q + 1;
    // The candidates' metadata and term records are the same for every tile: they are copied into shared
    // memory once (when they fit), so the inner loop's dependent loads are shared-memory latency, not L2.
    int4* s_meta = reinterpret_cast<int4*>(sm + (((2 * (size_t)rows * pitch) + 1) & ~size_t(1)));  // [256] {term offset, nnz, eq, padded nnz}
    RssTerm* s_terms = reinterpret_cast<RssTerm*>(s_meta + kRssWarps * kRssAcc);
    __shared__ int s_scan[kRssWarps];
    __shared__ int s_total;
    __shared__ double s_shift[kRssWarps * kRssAcc];
    {
        const int lc = threadIdx.x;  // local candidate id: warp w, slot q  <->  lc = w + 8 q
        const int c = cbase + lc;
        s_shift[lc] = (a.cand_shift && c < a.ncand) ? a.cand_shift[c] : 0.0;
        const int nnz = c < a.ncand ? a.cand_nnz[c] : 0;
        // shared-memory mode: every candidate's record list is padded to a multiple of 4 (padding records have
        // coefficient 0 and read the row of ones), so the term loop runs in branch-free groups of four
        const int n4 = (nnz + 3) & ~3;
        int v = n4;  // inclusive block scan of the padded counts -> offsets
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) s_scan[warp] = v;
        __syncthreads();
        int base = 0, total = 0;
        for (int w = 0; w < kRssWarps; ++w) {
            if (w < warp) base += s_scan[w];
            total += s_scan[w];
        }
        if (threadIdx.x == 0) s_total = total;
        const int loff = base + v - n4;
        // degree <= 2 keeps only the first 16 bytes of a record ({val, o32[0], o32[1]}): twice the capacity
        constexpr int kRec = NF <= 2 ? 1 : 2;  // int4 per record
        const bool fits = (int64_t)total * kRec <= (int64_t)a.term_cap * 2;
        s_meta[lc] = make_int4(loff, nnz, c < a.ncand ? a.cand_eq[c] : 0, n4);
        if (fits) {
            RssTerm pad;
            pad.val = 0.0;
            pad.o32[0] = pad.o32[1] = (rows - 1) * pitch;
#pragma unroll
            for (int d = 0; d < kMaxFactors; ++d) pad.o16[d] = (uint16_t)(rows - 1);
            const RssTerm* src = c < a.ncand ? a.terms + a.cand_off[c] : nullptr;
            int4* dst = reinterpret_cast<int4*>(s_terms) + (size_t)loff * kRec;
            for (int t = 0; t < n4; ++t) {
                const RssTerm rec = (t < nnz) ? src[t] : pad;
                const int4* p = reinterpret_cast<const int4*>(&rec);
                dst[t * kRec] = p[0];
                if (kRec == 2) dst[t * kRec + 1] = p[1];
            }
        }
        __syncthreads();
    }
    const int tp = s_total;
    double acc[kRssAcc];
#pragma unroll
    for (int q = 0; q < kRssAcc; ++q) acc[q] = 0.0;
    if ((int64_t)tp * (NF <= 2 ? 1 : 2) <= (int64_t)a.term_cap * 2)
        rss_tiles<NF, true>(a, sm, s_meta, s_terms, s_shift, nslots, cbase, t_lo, t_hi, tp, acc);
    else
        rss_tiles<NF, false>(a, sm, s_meta, s_terms, s_shift, nslots, cbase, t_lo, t_hi, tp, acc);
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

// ---------------------------------------------------------------------------------------------
// Few sparse candidates on wide samples (the STLSQ result at C3: 64 models of 4 terms, 1 KB per sample): lane = CANDIDATE.
//
// The kernel above maps lanes to samples, which needs the tile transposed on the way into shared memory (4096 8-byte
// cp.async with index arithmetic per 32 samples) and walks every candidate's term list with loads and branches: it is
// instruction bound (35 % of the HBM rate).  Here every lane keeps ITS candidate -- up to 8 (coefficient, factor,
// factor) triples -- in registers for the whole pass, the sample tile stays in its natural layout ([sample][state]:
// exactly the global layout), so a tile is fetched by TWO bulk copies through the TMA unit (X part, Y part) into a
// 3-deep ring, and a sample costs a warp 4 instructions per term with no loop overhead.  Warps 0-3 serve candidates
// 0-31, warps 4-7 candidates 32-63 of a pass; the four warps of a group take the samples of a tile round robin.
// Per-candidate sums: lane register -> fixed-order sum over the group's warps -> per-CTA partial (deterministic).
constexpr int kLaneTile = 32;    // samples per tile
constexpr int kLaneStages = 3;
constexpr int kLaneTerms = 8;    // terms per candidate held in registers

template <int NT>
__global__ void __launch_bounds__(256, 2) rss_lane_kernel(const RssArgs a) {
    extern __shared__ __align__(128) unsigned char lsm_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(lsm_raw);           // [kLaneStages]
    double* ring = reinterpret_cast<double*>(lsm_raw + 128);
    const int n = a.n, n_t = a.n_t, w = n + n_t;
    const int tile_doubles = kLaneTile * w;
    __shared__ double part[8][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp >> 2, sub = warp & 3;
    const int c = blockIdx.y * 64 + grp * 32 + lane;
    const bool live = c < a.ncand;
    // this lane's candidate: coefficients and factor columns (a missing factor is the constant 1)
    double val[NT];
    int fa[NT], fb[NT];
    int eq = 0;
    double shift = 0.0;
    {
        const int nnz = live ? a.cand_nnz[c] : 0;
        const RssTerm* t = live ? a.terms + a.cand_off[c] : nullptr;
        const int ones = n + n_t;
#pragma unroll
        for (int q = 0; q < NT; ++q) {
            val[q] = 0.0;
            fa[q] = fb[q] = -1;
            if (q < nnz) {
                val[q] = t[q].val;
                fa[q] = t[q].o16[0] == ones ? -1 : (int)t[q].o16[0];
                fb[q] = t[q].o16[1] == ones ? -1 : (int)t[q].o16[1];
            }
        }
        if (live) eq = a.cand_eq[c];
        if (live && a.cand_shift) shift = a.cand_shift[c];
    }
    if (tid == 0) {
        for (int i = 0; i < kLaneStages; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t ntiles = (a.m + kLaneTile - 1) / kLaneTile;
    const int64_t t_lo = ntiles * blockIdx.x / gridDim.x, t_hi = ntiles * (blockIdx.x + 1) / gridDim.x;
    const int64_t full_tiles = a.m / kLaneTile;  // tiles below this index hold kLaneTile samples (bulk copies)
    auto request = [&](int64_t tile) {  // thread 0: fetch `tile` into its ring slot
        const int slot = (int)((tile - t_lo) % kLaneStages);
        double* dst = ring + (size_t)slot * tile_doubles;
        const int64_t s0 = tile * kLaneTile;
        mbar_expect_tx(&full[slot], (uint32_t)(tile_doubles * sizeof(double)));
        bulk_g2s(dst, a.X + s0 * n, (uint32_t)(kLaneTile * n * sizeof(double)), &full[slot]);
        bulk_g2s(dst + kLaneTile * n, a.Y + s0 * n_t, (uint32_t)(kLaneTile * n_t * sizeof(double)), &full[slot]);
    };
    if (tid == 0)
        for (int i = 0; i < kLaneStages - 1 && t_lo + i < t_hi && t_lo + i < full_tiles; ++i) request(t_lo + i);
    double acc = 0.0;
    for (int64_t tile = t_lo; tile < t_hi; ++tile) {
        const int it = (int)(tile - t_lo);
        const int slot = it % kLaneStages;
        double* xs = ring + (size_t)slot * tile_doubles;
        int valid = kLaneTile;
        if (tile < full_tiles) {
            // refill the slot the previous iteration finished with (every warp passed the barrier at its end)
            if (tid == 0 && tile + kLaneStages - 1 < t_hi && tile + kLaneStages - 1 < full_tiles) request(tile + kLaneStages - 1);
            mbar_wait(&full[slot], (uint32_t)((it / kLaneStages) & 1));
        } else {  // ragged last tile of the data set: plain loads (a bulk copy needs 16-byte multiples)
            valid = (int)(a.m - tile * kLaneTile);
            const int64_t s0 = tile * kLaneTile;
            for (int e = tid; e < valid * n; e += 256) xs[e] = a.X[s0 * n + e];
            for (int e = tid; e < valid * n_t; e += 256) xs[kLaneTile * n + e] = a.Y[s0 * n_t + e];
            __syncthreads();
        }
        const double* ys = xs + kLaneTile * n;
        for (int sidx = sub; sidx < valid; sidx += 4) {
            const double* x = xs + sidx * n;
            double fit = shift;
#pragma unroll
            for (int q = 0; q < NT; ++q) {
                const double xa = fa[q] >= 0 ? x[fa[q]] : 1.0;
                const double xb = fb[q] >= 0 ? x[fb[q]] : 1.0;
                fit = fma(val[q], xa * xb, fit);
            }
            const double res = fit - ys[sidx * n_t + eq];
            acc = fma(res, res, acc);
        }
        __syncthreads();  // the slot may be refilled
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (sub == 0 && live) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) s += part[4 * grp + q][lane];
        a.partials[(size_t)blockIdx.x * a.ncand + c] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Few candidates on NARROW samples (C2: Lorenz, 3 states + 3 targets = 48 bytes per sample, three models of 2-3 terms):
// thread = SAMPLE.  The tiled kernels above pay a transposing stage-in and per-tile barriers that only amortise over
// many candidates; here a thread reads its sample straight from global memory (a warp's samples are contiguous), keeps
// it in a private shared-memory row (odd pitch: dynamic factor indices without bank conflicts), and walks the
// candidates' term records, which sit in shared memory once per CTA.  Per-candidate sums: thread -> warp shuffle tree
// -> CTA (fixed order) -> per-CTA partial -> rss_reduce_kernel.  C2: 0.70 -> see DESIGN 4.3 (HBM time 0.07 ms).
constexpr int kNarrowStates = 16;   // n + n_t
constexpr int kNarrowCands = 32;
constexpr int kNarrowTerms = 256;

template <int NF>
__global__ void __launch_bounds__(256) rss_narrow_kernel(const RssArgs a) {
    __shared__ RssTerm s_terms[kNarrowTerms];
    __shared__ int s_off[kNarrowCands + 1], s_eq[kNarrowCands];
    __shared__ double s_shift[kNarrowCands];
    __shared__ double s_red[8][kNarrowCands];
    extern __shared__ double rows[];  // 256 x pitch: [x_0 .. x_(n-1) | y_0 .. y_(n_t-1) | 1.0], then ncand x 256 sums
    const int n = a.n, n_t = a.n_t, w = n + n_t;
    const int pitch = (w + 1) | 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nc = a.ncand;
    if (tid == 0) {  // compact copy of the candidates' records (their pool offsets are arbitrary)
        int o = 0;
        for (int c = 0; c < nc; ++c) s_off[c] = o, o += a.cand_nnz[c];
        s_off[nc] = min(o, kNarrowTerms);
    }
    if (tid < nc) {
        s_eq[tid] = a.cand_eq[tid];
        s_shift[tid] = a.cand_shift ? a.cand_shift[tid] : 0.0;
    }
    __syncthreads();
    for (int c = 0; c < nc; ++c)
        for (int t = tid; t < s_off[c + 1] - s_off[c]; t += 256) s_terms[s_off[c] + t] = a.terms[a.cand_off[c] + t];
    double* row = rows + tid * pitch;
    row[w] = 1.0;
    __syncthreads();
    double* acc = rows + 256 * pitch + tid;  // [candidate][thread]: this thread's running sums (conflict-free)
    for (int c = 0; c < nc; ++c) acc[c * 256] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t s = (int64_t)blockIdx.x * 256 + tid; s < a.m; s += stride) {
        for (int i = 0; i < n; ++i) row[i] = a.X[s * n + i];
        for (int i = 0; i < n_t; ++i) row[n + i] = a.Y[s * n_t + i];
        for (int c = 0; c < nc; ++c) {
            double fit = s_shift[c];
            for (int t = s_off[c]; t < s_off[c + 1]; ++t) {
                const RssTerm& r = s_terms[t];
                double p = row[r.o16[0]];
#pragma unroll
                for (int d = 1; d < NF; ++d) p *= row[r.o16[d]];  // factors beyond the library's degree are ones
                fit = fma(r.val, p, fit);
            }
            const double res = fit - row[n + s_eq[c]];
            acc[c * 256] = fma(res, res, acc[c * 256]);
        }
    }
    for (int c = 0; c < nc; ++c) {
        double v = acc[c * 256];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][c] = v;
    }
    __syncthreads();
    if (tid < nc) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += s_red[q][tid];
        a.partials[(size_t)blockIdx.x * nc + tid] = v;
    }
}

__global__ void rss_reduce_kernel(const double* __restrict__ partials, int nparts, int ncand, double* __restrict__ rss) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partials[(size_t)p * ncand + c];  // fixed order
    rss[c] = s;
}

// The streaming pass proper: candidates given as device arrays (CSR of (term, value) + target per candidate).
// `map` translates term indices into rows of the installed library (implicit view) or is null; `scale` (per
// term of the CANDIDATE index space, or null) divides the values (term normalisation: the candidates live in
// the scaled space, the data are unscaled).  Writes ncand sums to dRss (device).
struct RssCand {
    int ncand;
    int64_t nterms;
    const int64_t* off;
    const int32_t* nnz;
    const int32_t* eq;
    const int32_t* idx;
    const double* val;
};

__global__ void rss_scale_kernel(RssTerm* __restrict__ terms, const int32_t* __restrict__ cand_idx, int64_t nterms,
                                 const double* __restrict__ scale) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nterms) terms[t].val /= scale[cand_idx[t]];
}
// affine term normalisation (theta~_j = (theta_j - mu_j) / scale_j): the candidate c~ evaluates on the raw terms as
// sum_j (c~_j / scale_j) theta_j + shift with shift = -sum_j (c~_j / scale_j) mu_j; terms[].val is already divided
__global__ void rss_shift_kernel(const RssTerm* __restrict__ terms, const int32_t* __restrict__ cand_idx,
                                 const int64_t* __restrict__ off, const int32_t* __restrict__ nnz, int ncand,
                                 const double* __restrict__ mu, double* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    double s = 0.0;
    for (int t = 0; t < nnz[c]; ++t) s -= terms[off[c] + t].val * mu[cand_idx[off[c] + t]];
    shift[c] = s;
}

// groups of 32 samples per staged tile (~16 KB per tile) and the resulting row pitch
static void rss_tile_shape(const ddb_ctx* ctx, int* sub_out, int* pitch_out) {
    const int rows_all = ctx->n + ctx->n_t + 1;
    int sub = 1;
    while (sub < kRssMaxSub && (size_t)rows_all * kRssTile * (2 * sub) * 8 <= 16 * 1024) sub *= 2;
    *sub_out = sub;
    *pitch_out = kRssTile * sub + 1;
}

static int rss_core(ddb_ctx* ctx, const RssCand& c, const int32_t* map, const RssTerm* targets, const double* scale,
                    DevBuf& terms, DevBuf& partials, double* dRss, int* launches, const double* mu = nullptr,
                    const double* extra_shift = nullptr) {
    if (ctx->view_empty) {  // no sample in this rank's view: every streamed sum is zero
        DDB_CUDA_TRY(cudaMemsetAsync(dRss, 0, (size_t)std::max(1, c.ncand) * 8, ctx->stream));
        return DDB_OK;
    }
    if (!ctx->dX || !ctx->dY || ctx->m <= 0) {
        set_error("ddb_rss: no data bound");
        return DDB_BAD_STATE;
    }
    if (ctx->factors_nt != ctx->n_t || !ctx->d_factors.p) {
        set_error("ddb_rss: the library factor table is missing (run ddb_gram on this context first)");
        return DDB_BAD_STATE;
    }
    cudaStream_t st = ctx->stream;
    const int ncand = c.ncand;
    int sub, pitch;
    rss_tile_shape(ctx, &sub, &pitch);
    DDB_TRY(terms.reserve((size_t)std::max<int64_t>(1, c.nterms) * sizeof(RssTerm)));
    if (c.nterms > 0) {
        rss_expand_kernel<<<(unsigned)((c.nterms + 255) / 256), 256, 0, st>>>(c.idx, c.val, c.nterms, ctx->d_factors.as<uint16_t>(),
                                                                              ctx->n + ctx->n_t, pitch, map, terms.as<RssTerm>());
        ++*launches;
        if (scale) {
            rss_scale_kernel<<<(unsigned)((c.nterms + 255) / 256), 256, 0, st>>>(terms.as<RssTerm>(), c.idx, c.nterms, scale);
            ++*launches;
        }
    }
    const double* cand_shift = extra_shift;
    if (mu) {  // affine normalisation: per-candidate constant (see rss_shift_kernel)
        DDB_TRY(ctx->d_cand_shift.reserve((size_t)std::max(1, ncand) * 8));
        rss_shift_kernel<<<(ncand + 255) / 256, 256, 0, st>>>(terms.as<RssTerm>(), c.idx, c.off, c.nnz, ncand, mu,
                                                              ctx->d_cand_shift.as<double>());
        ++*launches;
        cand_shift = ctx->d_cand_shift.as<double>();
    }
    // few candidates on narrow samples: thread = sample (see rss_narrow_kernel)
    {
        const char* nenv = getenv("DDB_RSS_NARROW");
        const int w = ctx->n + ctx->n_t;
        if (!(nenv && atoi(nenv) == 0) && !targets && w <= kNarrowStates && ncand > 0 && ncand <= kNarrowCands &&
            c.nterms <= kNarrowTerms && ctx->m >= 4096) {
            const int pitch_n = (w + 1) | 1;
            const int nparts = (int)std::min<int64_t>((ctx->m + 255) / 256, (int64_t)ctx->sm_count * 8);
            DDB_TRY(partials.reserve((size_t)nparts * ncand * 8));
            auto nkern = ctx->max_degree <= 2 ? rss_narrow_kernel<2> : ctx->max_degree <= 3 ? rss_narrow_kernel<3>
                         : ctx->max_degree <= 5 ? rss_narrow_kernel<5> : rss_narrow_kernel<kMaxFactors>;
            DDB_CUDA_TRY(cudaFuncSetAttribute(nkern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)((size_t)256 * (pitch_n + ncand) * sizeof(double))));
            RssArgs a;
            a.X = ctx->dX;
            a.Y = ctx->dY;
            a.n = ctx->n;
            a.n_t = ctx->n_t;
            a.m = ctx->m;
            a.ncand = ncand;
            a.cand_off = c.off;
            a.cand_nnz = c.nnz;
            a.cand_eq = c.eq;
            a.terms = terms.as<RssTerm>();
            a.targets = nullptr;
            a.cand_shift = cand_shift;
            a.term_cap = 0;
            a.sub = 1;
            a.pitch = pitch;
            a.partials = partials.as<double>();
            nkern<<<nparts, 256, (size_t)256 * (pitch_n + ncand) * sizeof(double), st>>>(a);
            rss_reduce_kernel<<<(ncand + 255) / 256, 256, 0, st>>>(partials.as<double>(), nparts, ncand, dRss);
            *launches += 2;
            DDB_CUDA_TRY(cudaGetLastError());
            return DDB_OK;
        }
    }
    // few sparse candidates on wide samples: the lane-per-candidate kernel (see rss_lane_kernel)
    const char* lenv = getenv("DDB_RSS_LANE");
    const bool lane_ok = !(lenv && atoi(lenv) == 0) && !targets && ctx->max_degree <= 2 && ncand > 0 && ncand <= 128 &&
                         (size_t)(ctx->n + ctx->n_t) * 8 >= 512 && ctx->m >= 4 * kLaneTile &&
                         ((size_t)kLaneTile * ctx->n * 8) % 16 == 0 && ((size_t)kLaneTile * ctx->n_t * 8) % 16 == 0 &&
                         (((uintptr_t)ctx->dX) & 15) == 0 && (((uintptr_t)ctx->dY) & 15) == 0;
    if (lane_ok) {
        std::vector<int32_t> h_nnz(ncand);
        DDB_CUDA_TRY(cudaMemcpyAsync(h_nnz.data(), c.nnz, (size_t)ncand * 4, cudaMemcpyDeviceToHost, st));
        DDB_CUDA_TRY(cudaStreamSynchronize(st));
        int nnz_max = 0;
        for (int v : h_nnz) nnz_max = std::max(nnz_max, v);
        const size_t lsmem = 128 + (size_t)kLaneStages * kLaneTile * (ctx->n + ctx->n_t) * sizeof(double);
        if (nnz_max <= kLaneTerms && lsmem <= 110 * 1024) {
            const int64_t lt = (ctx->m + kLaneTile - 1) / kLaneTile;
            const int nparts = (int)std::min<int64_t>(lt, (int64_t)ctx->sm_count * 2);
            DDB_TRY(partials.reserve((size_t)nparts * ncand * 8));
            RssArgs a;
            a.X = ctx->dX;
            a.Y = ctx->dY;
            a.n = ctx->n;
            a.n_t = ctx->n_t;
            a.m = ctx->m;
            a.ncand = ncand;
            a.cand_off = c.off;
            a.cand_nnz = c.nnz;
            a.cand_eq = c.eq;
            a.terms = terms.as<RssTerm>();
            a.targets = nullptr;
            a.cand_shift = cand_shift;
            a.term_cap = 0;
            a.sub = 1;
            a.pitch = pitch;
            a.partials = partials.as<double>();
            auto kern = nnz_max <= 4 ? rss_lane_kernel<4> : rss_lane_kernel<kLaneTerms>;
            DDB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
            kern<<<dim3(nparts, (ncand + 63) / 64), 256, lsmem, st>>>(a);
            rss_reduce_kernel<<<(ncand + 255) / 256, 256, 0, st>>>(partials.as<double>(), nparts, ncand, dRss);
            *launches += 2;
            DDB_CUDA_TRY(cudaGetLastError());
            return DDB_OK;
        }
    }
    const int passes = (ncand + kRssWarps * kRssAcc - 1) / (kRssWarps * kRssAcc);
    const int rows_all = ctx->n + ctx->n_t + 1;
    const int64_t ntiles = (ctx->m + kRssTile * sub - 1) / (kRssTile * sub);
    const int nparts = (int)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * 4);
    DDB_TRY(partials.reserve((size_t)nparts * ncand * 8));
    RssArgs a;
    a.X = ctx->dX;
    a.Y = ctx->dY;
    a.n = ctx->n;
    a.n_t = ctx->n_t;
    a.m = ctx->m;
    a.ncand = ncand;
    a.cand_off = c.off;
    a.cand_nnz = c.nnz;
    a.cand_eq = c.eq;
    a.terms = terms.as<RssTerm>();
    a.targets = targets;
    a.cand_shift = cand_shift;
    a.partials = partials.as<double>();
    const size_t smem_tiles = ((2 * (size_t)pitch * rows_all + 1) & ~size_t(1)) * sizeof(double);
    a.sub = sub;
    a.pitch = pitch;
    if (smem_tiles > 180 * 1024) {
        set_error("ddb_rss: n + n_t too large for the shared-memory sample tile");
        return DDB_UNSUPPORTED;
    }
    // candidate metadata + as many term records as keep two CTAs per SM: 228 KB per SM, 1 KB reserved per CTA and
    // 2.2 KB of static shared memory (s_shift, s_scan) leave 110 KB of dynamic shared memory each.  (112 KB -- the
    // figure before s_shift existed -- silently dropped the kernel to ONE CTA per SM: C4 / ADMM 56 -> 98 ms.)
    const size_t smem_meta = (size_t)kRssWarps * kRssAcc * sizeof(int4);
    constexpr size_t kRssDynBudget = 110 * 1024;
    const size_t budget = smem_tiles + smem_meta < kRssDynBudget ? kRssDynBudget - smem_tiles - smem_meta : 0;
    a.term_cap = (int)std::min<size_t>(budget / sizeof(RssTerm), 1u << 15);
    const size_t smem = smem_tiles + smem_meta + (size_t)a.term_cap * sizeof(RssTerm);
    auto kern = ctx->max_degree <= 2 ? rss_kernel<2> : rss_kernel<kMaxFactors>;
    DDB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(nparts, passes), kRssWarps * 32, smem, st>>>(a);
    rss_reduce_kernel<<<(ncand + 255) / 256, 256, 0, st>>>(partials.as<double>(), nparts, ncand, dRss);
    *launches += 2;
    DDB_CUDA_TRY(cudaGetLastError());
    return DDB_OK;
}

// Candidates of one target that share a support S (ADMM and SR3 record one per iteration: thousands on one
// support) are not streamed one by one.  For a reference point c0 on S,
//     rss(c) = rss(c0) - 2 d' g0 + d' G_SS d,      d = c - c0,   g0 = r_S - G_SS c0,
// and with c0 = the least-squares solution on S the cross term vanishes up to rounding: rss(c) is a SUM of
// two non-negative terms, so there is no cancellation (unlike y'y - 2 c'r + c'Gc).  Only c0 is streamed
// (exactly, sample by sample); the small terms come from the Gram blocks in extended precision on the host.
// This is also more accurate than streaming c itself when the fit is nearly exact (the per-sample residual
// y - c'theta then loses its leading digits).  Supports of more than kGroupMax terms and singletons are
// streamed directly.
constexpr int kGroupMax = 16;
// Candidates with many terms are scored by the DMMA pass of rss_dense.cu.  Its cost per candidate and sample is that
// of ALL library terms at the tensor rate (2 k / 512 DMMAs, ~53 SM cycles at k = 2145 and the measured 66 % of the DMMA
// peak); the term-by-term kernels above cost ~0.34 SM cycles per term: break-even near k / 14 terms.
constexpr int kDenseMin = 96;
static int dense_threshold(int k) { return std::max(kDenseMin, k / 14); }
bool rss_dense_ok(const ddb_ctx* ctx);
int rss_dense_run(ddb_ctx* ctx, const int32_t* ids, int nd, const int64_t* off, const int32_t* nnz, const int32_t* eq,
                  const int32_t* idx, const double* val, const double* scale, const int32_t* map, const double* shift,
                  DevBuf& work, double* out, int* launches);

__global__ void gather_gss_kernel(const double* __restrict__ packed, int k, int n_t, const int32_t* __restrict__ gidx,
                                  const int32_t* __restrict__ gsize, const int32_t* __restrict__ geq,
                                  double* __restrict__ out) {
    // group g: out[g] = [G_SS (kGroupMax x kGroupMax, row-major) | r_S (kGroupMax)]
    const int g = blockIdx.x;
    const int s = gsize[g], e = geq[g];
    const int32_t* idx = gidx + (size_t)g * kGroupMax;
    double* o = out + (size_t)g * (kGroupMax * kGroupMax + kGroupMax);
    const double* R = packed + (size_t)k * k;
    for (int t = threadIdx.x; t < s * s; t += blockDim.x) {
        const int i = t / s, j = t % s;
        o[i * kGroupMax + j] = packed[(size_t)idx[i] + (size_t)k * idx[j]];
    }
    for (int i = threadIdx.x; i < s; i += blockDim.x) o[kGroupMax * kGroupMax + i] = R[(size_t)idx[i] + (size_t)k * e];
}

__global__ void scatter_group_rss_kernel(const double* __restrict__ rep, const int32_t* __restrict__ group, int ncand,
                                         double* __restrict__ rss) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncand) rss[c] = group[c] >= 0 ? rep[group[c]] : 0.0;
}

// Candidates 0 .. n_t-1 are the zero models: their rss is sum_s y_e(s)^2 = yy[e], which the Gram stage already
// holds (a sum of squares: no cancellation), so they are never streamed.  The streamed part is 0 and yy goes
// into the correction that select adds AFTER the all-reduce (yy is the reduced, global value on every rank).
__global__ void zero_model_rss_kernel(const double* __restrict__ yy, int n_t, double* __restrict__ rss,
                                      double* __restrict__ corr) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_t) {
        rss[e] = 0.0;
        corr[e] = yy[e];
    }
}

int rss_run(ddb_ctx* ctx, SweepState* S, double* dRss) {
    const int ncand = S->ncand;
    const bool implicit = ctx->imp_k > 0;
    cudaStream_t st = ctx->stream;
    if (!dRss) {
        DDB_TRY(S->rss.reserve((size_t)std::max(1, ncand) * 8));
        dRss = S->rss.as<double>();
    }
    DDB_TRY(S->rss_corr.reserve((size_t)std::max(1, ncand) * 8));
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[4], st));
    int launches = 0;
    if (implicit) {  // one record per target: the library term on the left-hand side
        DDB_TRY(S->tgt_terms.reserve((size_t)ctx->imp_k * sizeof(RssTerm)));
        int tsub, tpitch;
        rss_tile_shape(ctx, &tsub, &tpitch);
        rss_expand_kernel<<<(ctx->imp_k + 255) / 256, 256, 0, st>>>(nullptr, nullptr, ctx->imp_k, ctx->d_factors.as<uint16_t>(),
                                                                   ctx->n + ctx->n_t, tpitch, ctx->imp_idx.as<int32_t>(),
                                                                   S->tgt_terms.as<RssTerm>());
        ++launches;
    }
    const int32_t* map = implicit ? ctx->imp_idx.as<int32_t>() : nullptr;
    const RssTerm* targets = implicit ? S->tgt_terms.as<RssTerm>() : nullptr;
    const double* scale = (!implicit && ctx->term_scale_on) ? ctx->d_term_scale.as<double>() : nullptr;
    const double* mu = (!implicit && ctx->term_scale_on && ctx->term_shift_on) ? ctx->d_term_shift.as<double>() : nullptr;

    // ---- group the candidates by (target, support) -----------------------------------------------
    const int64_t nterms = S->pool_used;
    const char* genv = getenv("DDB_RSS_GROUPING");
    const bool grouping = !(genv && atoi(genv) == 0) && ncand > 2 * S->n_t && nterms > 0;
    // yy of the space the sweep ran in (explicit: [G | R | yy]; implicit view: [G' | R' | diag G'])
    const int n_t = S->n_t;
    const double* yy_dev = (implicit ? ctx->imp_packed.as<double>() : ctx->d_packed.as<double>()) + (size_t)S->k * S->k +
                           (size_t)S->k * n_t;
    if (!grouping) {
        DDB_CUDA_TRY(cudaMemsetAsync(S->rss_corr.p, 0, (size_t)std::max(1, ncand) * 8, st));
        zero_model_rss_kernel<<<(n_t + 255) / 256, 256, 0, st>>>(yy_dev, n_t, dRss, S->rss_corr.as<double>());
        ++launches;
        if (ncand > n_t) {  // candidates n_t .. ncand-1 are the recorded models
            RssCand c{ncand - n_t, nterms, S->cand_off.as<int64_t>() + n_t, S->cand_nnz.as<int32_t>() + n_t,
                      S->cand_eq.as<int32_t>() + n_t, S->cand_idx.as<int32_t>(), S->cand_val.as<double>()};
            DDB_TRY(rss_core(ctx, c, map, targets, scale, S->cand_terms, S->rss_partials, dRss + n_t, &launches, mu));
        }
    } else {
        std::vector<int64_t> off(ncand);
        std::vector<int32_t> nnz(ncand), eq(ncand), idx((size_t)nterms);
        std::vector<double> val((size_t)nterms);
        DDB_CUDA_TRY(cudaMemcpyAsync(off.data(), S->cand_off.p, (size_t)ncand * 8, cudaMemcpyDeviceToHost, st));
        DDB_CUDA_TRY(cudaMemcpyAsync(nnz.data(), S->cand_nnz.p, (size_t)ncand * 4, cudaMemcpyDeviceToHost, st));
        DDB_CUDA_TRY(cudaMemcpyAsync(eq.data(), S->cand_eq.p, (size_t)ncand * 4, cudaMemcpyDeviceToHost, st));
        DDB_CUDA_TRY(cudaMemcpyAsync(idx.data(), S->cand_idx.p, (size_t)nterms * 4, cudaMemcpyDeviceToHost, st));
        DDB_CUDA_TRY(cudaMemcpyAsync(val.data(), S->cand_val.p, (size_t)nterms * 8, cudaMemcpyDeviceToHost, st));
        DDB_CUDA_TRY(cudaStreamSynchronize(st));
        // groups: key = (target, support); a group is "shared" when it has >= 2 members and 1..kGroupMax terms
        std::unordered_map<std::string, int> key2group;
        std::vector<int32_t> group(ncand);
        std::vector<int> first, members;  // first member and member count per group
        std::string key;
        for (int c = 0; c < ncand; ++c) {
            key.assign(reinterpret_cast<const char*>(&eq[c]), 4);
            key.append(reinterpret_cast<const char*>(idx.data() + off[c]), (size_t)nnz[c] * 4);
            auto it = key2group.find(key);
            if (it == key2group.end()) {
                it = key2group.emplace(key, (int)first.size()).first;
                first.push_back(c);
                members.push_back(0);
            }
            group[c] = it->second;
            ++members[it->second];
        }
        // split groups that cannot use the identity back into singletons
        const int ng0 = (int)first.size();
        std::vector<char> shared(ng0);
        int nshared = 0;
        for (int g = 0; g < ng0; ++g) {
            shared[g] = members[g] >= 2 && nnz[first[g]] >= 1 && nnz[first[g]] <= kGroupMax;
            nshared += shared[g];
        }
        // representative list: shared groups first (one each), then every other candidate on its own
        std::vector<int> gslot(ng0, -1);
        std::vector<int32_t> g_idx((size_t)std::max(1, nshared) * kGroupMax, 0), g_size(std::max(1, nshared)), g_eq(std::max(1, nshared));
        int ns = 0;
        for (int g = 0; g < ng0; ++g)
            if (shared[g]) {
                gslot[g] = ns;
                const int c = first[g];
                g_size[ns] = nnz[c];
                g_eq[ns] = eq[c];
                for (int t = 0; t < nnz[c]; ++t) g_idx[(size_t)ns * kGroupMax + t] = idx[off[c] + t];
                ++ns;
            }
        // G_SS and r_S of the shared groups (in the space the sweep ran in: scaled blocks when a term
        // normalisation is on, the gathered sub-blocks in implicit mode)
        const int kk = implicit ? ctx->imp_k : ctx->k, ntt = implicit ? ctx->imp_k : ctx->n_t;
        const double* packed = implicit ? ctx->imp_packed.as<double>() : ctx->d_packed.as<double>();
        const size_t gstride = kGroupMax * kGroupMax + kGroupMax;
        std::vector<double> gss((size_t)std::max(1, ns) * gstride);
        if (ns > 0) {
            const size_t b_idx = (size_t)ns * kGroupMax * 4, b_sz = (size_t)ns * 4;
            const size_t o_sz = (b_idx + 7) & ~size_t(7), o_eq = (o_sz + b_sz + 7) & ~size_t(7), o_out = (o_eq + b_sz + 7) & ~size_t(7);
            DDB_TRY(S->grp_buf.reserve(o_out + (size_t)ns * gstride * 8));
            unsigned char* gb = static_cast<unsigned char*>(S->grp_buf.p);
            DDB_CUDA_TRY(cudaMemcpyAsync(gb, g_idx.data(), b_idx, cudaMemcpyHostToDevice, st));
            DDB_CUDA_TRY(cudaMemcpyAsync(gb + o_sz, g_size.data(), b_sz, cudaMemcpyHostToDevice, st));
            DDB_CUDA_TRY(cudaMemcpyAsync(gb + o_eq, g_eq.data(), b_sz, cudaMemcpyHostToDevice, st));
            gather_gss_kernel<<<ns, 256, 0, st>>>(packed, kk, ntt, reinterpret_cast<const int32_t*>(gb),
                                                  reinterpret_cast<const int32_t*>(gb + o_sz),
                                                  reinterpret_cast<const int32_t*>(gb + o_eq),
                                                  reinterpret_cast<double*>(gb + o_out));
            ++launches;
            DDB_CUDA_TRY(cudaMemcpyAsync(gss.data(), gb + o_out, (size_t)ns * gstride * 8, cudaMemcpyDeviceToHost, st));
            DDB_CUDA_TRY(cudaStreamSynchronize(st));
        }
        // least-squares reference point per shared group (Cholesky in extended precision), corrections
        std::vector<double> corr(ncand, 0.0), c0((size_t)std::max(1, ns) * kGroupMax, 0.0);
        std::vector<long double> g0((size_t)std::max(1, ns) * kGroupMax, 0.0L);
        for (int g = 0; g < ng0; ++g) {
            if (!shared[g]) continue;
            const int q = gslot[g], s = g_size[q];
            const double* Gs = gss.data() + (size_t)q * gstride;
            const double* rS = Gs + kGroupMax * kGroupMax;
            long double L[kGroupMax][kGroupMax], z[kGroupMax];
            bool ok = true;
            for (int i = 0; i < s && ok; ++i)
                for (int j = 0; j <= i; ++j) {
                    long double a = Gs[i * kGroupMax + j];
                    for (int t = 0; t < j; ++t) a -= L[i][t] * L[j][t];
                    if (i == j) {
                        if (!(a > 0.0L)) {
                            ok = false;
                            break;
                        }
                        L[i][i] = sqrtl(a);
                    } else {
                        L[i][j] = a / L[j][j];
                    }
                }
            if (!ok) {  // singular sub-block: stream every member directly
                shared[g] = 0;
                continue;
            }
            for (int i = 0; i < s; ++i) {
                long double a = rS[i];
                for (int t = 0; t < i; ++t) a -= L[i][t] * z[t];
                z[i] = a / L[i][i];
            }
            for (int i = s - 1; i >= 0; --i) {
                long double a = z[i];
                for (int t = i + 1; t < s; ++t) a -= L[t][i] * z[t];
                z[i] = a / L[i][i];
            }
            for (int i = 0; i < s; ++i) c0[(size_t)q * kGroupMax + i] = (double)z[i];
            for (int i = 0; i < s; ++i) {  // g0 = r_S - G_SS c0 with the ROUNDED c0 that is streamed
                long double a = rS[i];
                for (int j = 0; j < s; ++j) a -= (long double)Gs[i * kGroupMax + j] * (long double)c0[(size_t)q * kGroupMax + j];
                g0[(size_t)q * kGroupMax + i] = a;
            }
        }
        // representatives: shared groups (their c0), then all remaining candidates
        std::vector<int64_t> r_off;
        std::vector<int32_t> r_nnz, r_eq, r_idx;
        std::vector<double> r_val;
        std::vector<int> rep_of_group(ng0, -1);
        for (int g = 0; g < ng0; ++g) {
            if (!shared[g]) continue;
            const int q = gslot[g], s = g_size[q];
            rep_of_group[g] = (int)r_off.size();
            r_off.push_back((int64_t)r_idx.size());
            r_nnz.push_back(s);
            r_eq.push_back(g_eq[q]);
            for (int t = 0; t < s; ++t) {
                r_idx.push_back(g_idx[(size_t)q * kGroupMax + t]);
                r_val.push_back(c0[(size_t)q * kGroupMax + t]);
            }
        }
        std::vector<double> yy_h(n_t);
        DDB_CUDA_TRY(cudaMemcpyAsync(yy_h.data(), yy_dev, (size_t)n_t * 8, cudaMemcpyDeviceToHost, st));
        DDB_CUDA_TRY(cudaStreamSynchronize(st));
        std::vector<int32_t> rep(ncand);
        // dense candidates (explicit libraries, no affine term shift): scored on the tensor pipe, see rss_dense.cu
        const bool dense_on = !implicit && !mu && rss_dense_ok(ctx);
        const int dense_min = dense_threshold(ctx->k);
        std::vector<int32_t> dense_ids;
        for (int c = 0; c < ncand; ++c) {
            const int g = group[c];
            if (nnz[c] == 0) {  // a zero model: rss = yy[target], never streamed
                rep[c] = -1;
                corr[c] = yy_h[eq[c]];
            } else if (shared[g]) {
                rep[c] = rep_of_group[g];
                const int q = gslot[g], s = g_size[q];
                const double* Gs = gss.data() + (size_t)q * gstride;
                long double d[kGroupMax], acc = 0.0L;
                for (int t = 0; t < s; ++t) d[t] = (long double)val[off[c] + t] - (long double)c0[(size_t)q * kGroupMax + t];
                for (int i = 0; i < s; ++i) {
                    long double gd = 0.0L;
                    for (int j = 0; j < s; ++j) gd += (long double)Gs[i * kGroupMax + j] * d[j];
                    acc += d[i] * (gd - 2.0L * g0[(size_t)q * kGroupMax + i]);
                }
                corr[c] = (double)acc;
            } else if (dense_on && nnz[c] >= dense_min) {
                rep[c] = -2 - (int)dense_ids.size();  // placed behind the streamed representatives below
                dense_ids.push_back(c);
            } else {
                rep[c] = (int)r_off.size();
                r_off.push_back((int64_t)r_idx.size());
                r_nnz.push_back(nnz[c]);
                r_eq.push_back(eq[c]);
                for (int t = 0; t < nnz[c]; ++t) {
                    r_idx.push_back(idx[off[c] + t]);
                    r_val.push_back(val[off[c] + t]);
                }
            }
        }
        const int nsparse = (int)r_off.size(), nd = (int)dense_ids.size();
        for (int c = 0; c < ncand; ++c)
            if (rep[c] <= -2) rep[c] = nsparse + (-2 - rep[c]);
        const int nrep = nsparse + nd;  // results: [streamed representatives | dense candidates]
        const int64_t rterms = (int64_t)r_idx.size();
        const size_t o_nnz = (size_t)nrep * 8, o_eq = o_nnz + (size_t)nrep * 4, o_idx = (o_eq + (size_t)nrep * 4 + 7) & ~size_t(7);
        const size_t o_val = (o_idx + (size_t)std::max<int64_t>(1, rterms) * 4 + 7) & ~size_t(7);
        const size_t o_rep = o_val + (size_t)std::max<int64_t>(1, rterms) * 8;
        const size_t o_map = o_rep + (size_t)nrep * 8;
        DDB_TRY(S->rep_buf.reserve(o_map + (size_t)ncand * 4 + 64));
        unsigned char* rb = static_cast<unsigned char*>(S->rep_buf.p);
        if (nsparse) {
            DDB_CUDA_TRY(cudaMemcpyAsync(rb, r_off.data(), (size_t)nsparse * 8, cudaMemcpyHostToDevice, st));
            DDB_CUDA_TRY(cudaMemcpyAsync(rb + o_nnz, r_nnz.data(), (size_t)nsparse * 4, cudaMemcpyHostToDevice, st));
            DDB_CUDA_TRY(cudaMemcpyAsync(rb + o_eq, r_eq.data(), (size_t)nsparse * 4, cudaMemcpyHostToDevice, st));
        }
        if (rterms) {
            DDB_CUDA_TRY(cudaMemcpyAsync(rb + o_idx, r_idx.data(), (size_t)rterms * 4, cudaMemcpyHostToDevice, st));
            DDB_CUDA_TRY(cudaMemcpyAsync(rb + o_val, r_val.data(), (size_t)rterms * 8, cudaMemcpyHostToDevice, st));
        }
        DDB_CUDA_TRY(cudaMemcpyAsync(rb + o_map, rep.data(), (size_t)ncand * 4, cudaMemcpyHostToDevice, st));
        DDB_CUDA_TRY(cudaMemcpyAsync(S->rss_corr.p, corr.data(), (size_t)ncand * 8, cudaMemcpyHostToDevice, st));
        RssCand c{nsparse, rterms, reinterpret_cast<const int64_t*>(rb), reinterpret_cast<const int32_t*>(rb + o_nnz),
                  reinterpret_cast<const int32_t*>(rb + o_eq), reinterpret_cast<const int32_t*>(rb + o_idx),
                  reinterpret_cast<const double*>(rb + o_val)};
        if (nsparse > 0)
            DDB_TRY(rss_core(ctx, c, map, targets, scale, S->cand_terms, S->rss_partials, reinterpret_cast<double*>(rb + o_rep),
                             &launches, mu));
        if (nd > 0) {
            DDB_TRY(S->dense_ids.reserve((size_t)nd * 4));
            DDB_CUDA_TRY(cudaMemcpyAsync(S->dense_ids.p, dense_ids.data(), (size_t)nd * 4, cudaMemcpyHostToDevice, st));
            DDB_TRY(rss_dense_run(ctx, S->dense_ids.as<int32_t>(), nd, S->cand_off.as<int64_t>(), S->cand_nnz.as<int32_t>(),
                                  S->cand_eq.as<int32_t>(), S->cand_idx.as<int32_t>(), S->cand_val.as<double>(), scale, nullptr,
                                  nullptr, S->dense_work, reinterpret_cast<double*>(rb + o_rep) + nsparse, &launches));
        }
        scatter_group_rss_kernel<<<(ncand + 255) / 256, 256, 0, st>>>(reinterpret_cast<const double*>(rb + o_rep),
                                                                      reinterpret_cast<const int32_t*>(rb + o_map), ncand, dRss);
        ++launches;
        DDB_CUDA_TRY(cudaGetLastError());
        DDB_CUDA_TRY(cudaStreamSynchronize(st));  // the host vectors above are stack-lifetime
        ctx->timings.candidates = ncand;
        S->nstreamed = nrep;
        if (getenv("DDB_RSS_DEBUG"))
            fprintf(stderr, "rss grouping: %d candidates (%lld terms) -> %d streamed (%lld terms) + %d dense, %d shared groups\n",
                    ncand, (long long)nterms, nsparse, (long long)rterms, nd, ns);
    }
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[5], st));
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    ctx->timings.rss_ms = ms;
    ctx->timings.sweep_launches += launches;
    S->rss_ptr = dRss;
    S->have_rss = true;
    return DDB_OK;
}

// ddb_residual: rss of an arbitrary coefficient matrix (n_t x k column-major, host) over the current samples
int residual_run(ddb_ctx* ctx, const double* coef, const double* offset, double* rss_out) {
    const int k = ctx->k, n_t = ctx->n_t;
    if (k <= 0 || n_t <= 0 || !coef || !rss_out) {
        set_error("ddb_residual: library and data must be set, coef and rss non-null");
        return DDB_INVALID_ARG;
    }
    if (ctx->factors_nt != n_t || !ctx->d_factors.p) {
        extern int gram_ensure_factors(ddb_ctx * ctx, int kaug_pad);
        DDB_TRY(gram_ensure_factors(ctx, ((k + n_t + 127) / 128) * 128));
    }
    std::vector<int64_t> off(n_t);
    std::vector<int32_t> nnz(n_t), eq(n_t), idx;
    std::vector<double> val;
    for (int e = 0; e < n_t; ++e) {
        off[e] = (int64_t)idx.size();
        eq[e] = e;
        for (int j = 0; j < k; ++j) {
            const double v = coef[(size_t)e + (size_t)n_t * j];
            if (v != 0.0) {
                idx.push_back(j);
                val.push_back(v);
            }
        }
        nnz[e] = (int32_t)((int64_t)idx.size() - off[e]);
    }
    const int64_t nterms = (int64_t)idx.size();
    cudaStream_t st = ctx->stream;
    DevBuf& B = ctx->d_resid;  // [off | nnz | eq | idx | val | rss | offset]
    const size_t o_nnz = (size_t)n_t * 8, o_eq = o_nnz + (size_t)n_t * 4, o_idx = (o_eq + (size_t)n_t * 4 + 7) & ~size_t(7);
    const size_t o_val = (o_idx + (size_t)std::max<int64_t>(1, nterms) * 4 + 7) & ~size_t(7);
    const size_t o_rss = o_val + (size_t)std::max<int64_t>(1, nterms) * 8;
    DDB_TRY(B.reserve(o_rss + 2 * (size_t)n_t * 8));
    unsigned char* base = static_cast<unsigned char*>(B.p);
    DDB_CUDA_TRY(cudaMemcpyAsync(base, off.data(), (size_t)n_t * 8, cudaMemcpyHostToDevice, st));
    DDB_CUDA_TRY(cudaMemcpyAsync(base + o_nnz, nnz.data(), (size_t)n_t * 4, cudaMemcpyHostToDevice, st));
    DDB_CUDA_TRY(cudaMemcpyAsync(base + o_eq, eq.data(), (size_t)n_t * 4, cudaMemcpyHostToDevice, st));
    if (nterms) {
        DDB_CUDA_TRY(cudaMemcpyAsync(base + o_idx, idx.data(), (size_t)nterms * 4, cudaMemcpyHostToDevice, st));
        DDB_CUDA_TRY(cudaMemcpyAsync(base + o_val, val.data(), (size_t)nterms * 8, cudaMemcpyHostToDevice, st));
    }
    RssCand c{n_t, nterms, reinterpret_cast<const int64_t*>(base), reinterpret_cast<const int32_t*>(base + o_nnz),
              reinterpret_cast<const int32_t*>(base + o_eq), reinterpret_cast<const int32_t*>(base + o_idx),
              reinterpret_cast<const double*>(base + o_val)};
    const double* d_offset = nullptr;
    if (offset) {
        DDB_CUDA_TRY(cudaMemcpyAsync(base + o_rss + (size_t)n_t * 8, offset, (size_t)n_t * 8, cudaMemcpyHostToDevice, st));
        d_offset = reinterpret_cast<const double*>(base + o_rss + (size_t)n_t * 8);
    }
    int launches = 0;
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[4], st));
    DDB_TRY(rss_core(ctx, c, nullptr, nullptr, nullptr, ctx->d_resid_terms, ctx->d_resid_partials,
                     reinterpret_cast<double*>(base + o_rss), &launches, nullptr, d_offset));
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[5], st));
    DDB_CUDA_TRY(cudaMemcpyAsync(rss_out, base + o_rss, (size_t)n_t * 8, cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaStreamSynchronize(st));  // off/nnz/... are stack-lifetime host buffers
    float ms = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    ctx->timings.rss_ms = ms;
    return DDB_OK;
}

}  // namespace ddb
