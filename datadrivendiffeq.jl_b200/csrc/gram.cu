// Fused candidate-library evaluation + normal-equation blocks on FP64 DMMA tensor cores.
//
// Replaces, for all targets at once, the library evaluation `(b::Basis)(prob)`
// (reference src/problem/type.jl:404-407 -> src/basis/build_function.jl:186-205) and the
// per-target `A*A'`, `B*A'` of lib/DataDrivenSparse/src/algorithms/STLSQ.jl:90-91,
// ADMM.jl:77-79, SR3.jl:108-109.
//
// Formulation.  Per sample s the augmented row vector is z(s) = [theta_1..theta_k, y_1..y_nt]
// (k library monomials of the states followed by the n_t targets).  The lower triangle of
// Z Z' = sum_s z z' holds G = Theta Theta' (k x k), R' = Y Theta' (n_t x k) and diag(Y Y').
// It is tiled into BT x BT block pairs (bi >= bj); the contraction index is the SAMPLE.
//
// Per CTA (persistent; the host plan hands it a list of (block pair, stage range) segments),
// warp-specialised, stage = 16 consecutive samples:
//   * TMA ring (3 deep): one producer thread issues cp.async.bulk copies of the X tile (16 x n
//     doubles, contiguous because the reference layout is sample-contiguous) and the Y tile into
//     shared memory, completion signalled on an mbarrier, running ahead across segment boundaries.
//   * Producer warps: thread p owns column p of the block row and of the block column; its two
//     monomial factor lists sit in registers; per stage it multiplies the factors out for the 16
//     samples and writes Theta_I' / Theta_J' (16 x BT, sample-major, pitch BT+4 -> conflict-free
//     fragment loads) into a 4-deep shared-memory ring.  Theta never reaches HBM.
//   * MMA warps: each owns a 64 x 32 patch of the BT x BT block (32 independent 8x8 accumulator
//     tiles in registers for the whole segment) and issues mma.sync.m8n8k4.f64 -- SASS DMMA.8x8x4,
//     the only FP64 tensor shape sm_100a implements -- with A/B fragments read from the ring.
//     Sub-tiles above the diagonal of a diagonal block pair and rows/columns past the end of the
//     library are skipped (bit mask per warp); the warp->patch map balances that over the four
//     SM sub-partitions.
//   Full/empty mbarriers connect the three roles; there is no CTA-wide barrier in steady state.
//   Accumulators are flushed once per segment to a partial-tile slot; a second kernel adds the
//   slots of each block pair in a fixed order (bit-reproducible: no floating-point atomics) and
//   scatters into the packed [G | R | yy].
#include "common.cuh"
#include "mma_common.cuh"
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <map>

namespace ddb {


constexpr int kNVS = 4;   // depth of the raw-sample (TMA) ring
constexpr int kNTS = 4;   // depth of the Theta-tile ring (tiles are built kAhead stages before use)
constexpr int kAhead = 2; // build distance: barrier arrivals precede the matching wait by a whole stage
template <int BT>
struct GramCfg {
    static constexpr int kWarpCols = BT / 32;            // warp grid (BT/64) x (BT/32), warp patch 64 x 32
    static constexpr int kWarpRows = BT / 64;
    static constexpr int kWarps = kWarpRows * kWarpCols;  // 8 (BT=128) or 2 (BT=64)
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kJobs = 2 * BT / kThreads;       // Theta columns built per thread: 1 or 2
    static constexpr int kPitch = BT + 4;  // doubles; pitch % 16 == 4 -> fragment loads hit 16 distinct banks
    // Theta_I' and Theta_J' tiles of one stage; BT = 64 only ever runs the single diagonal pair of a library of
    // at most 64 augmented rows (one tile), which lets four CTAs share an SM: their heavy (26 sub-tiles) and
    // light (10 sub-tiles) warps then pair up on the four sub-partitions
    static constexpr int kStageDoubles = (BT == 128 ? 2 : 1) * kKS * kPitch;
    static constexpr int kMinBlocks = (BT == 128) ? 1 : 4;
};

// One raw-sample slot (doubles): [ones(16) | 0.0 | pad | X tile 16 x n | Y tile 16 x n_t].
// ones[s] is 1.0 for a real sample and 0.0 for the zero padding of a ragged last stage, so a monomial
// with fewer than NF factors reads ones[s] for the missing ones (stride 1) and a padding row of the
// library reads the 0.0 (stride 0): every Theta entry is a product of exactly NF shared-memory values,
// no predicates in the hot loop.
__host__ __device__ inline size_t gram_vbytes(int n, int n_t) {
    size_t v = (size_t)(kVX + kKS * (n + n_t)) * sizeof(double);
    return (v + 127) & ~size_t(127);
}
template <int BT>
inline size_t gram_smem_bytes(int n, int n_t) {
    return 128 + kNVS * gram_vbytes(n, n_t) + (size_t)kNTS * GramCfg<BT>::kStageDoubles * sizeof(double);
}

// The per-thread share of building a later stage's Theta tiles: NJ columns (library terms), each a
// product of NF staged values per sample.
template <int NJ, int NF>
struct BuildJobs {
    int off[NJ][NF];  // offset of factor d at sample 0 inside a raw-sample slot
    int str[NJ][NF];  // per-sample stride of factor d
    int dst[NJ];      // offset of the column inside a Theta stage
    int s0;           // first sample this thread builds
    bool live;        // false: the plan does not need this thread's column (rows past 64 / 32 of a half / quarter pair)
};

// One stage (16 samples): 4 k-steps of {build SPK samples of a later stage, load fragments, DMMAs}.
// Straight-line code (the live sub-tiles are a template parameter) so that ptxas can software-
// pipeline the shared-memory loads under the DMMAs and no DMMA is predicated.
template <int BT, int NF, uint32_t MASK, bool SAME, bool BUILD = true>
__device__ __forceinline__ void stage_body(double (&acc)[8][4][2], const double* __restrict__ pa,
                                           const double* __restrict__ pb, const double* __restrict__ v,
                                           double* __restrict__ tnext,
                                           const BuildJobs<GramCfg<BT>::kJobs, NF>& jb) {
    constexpr int LDT = GramCfg<BT>::kPitch;
    constexpr int NJ = (BT == 128) ? 1 : (SAME ? 1 : 2);   // columns this thread builds
    constexpr int SPK = (BT == 128 && SAME) ? 2 : 4;        // samples per column per k-step
#pragma unroll
    for (int kk = 0; kk < kKS / 4; ++kk) {
        double fv[NJ][SPK][NF];
        if constexpr (BUILD) {
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int q = 0; q < SPK; ++q)
#pragma unroll
                    for (int d = 0; d < NF; ++d) fv[j][q][d] = v[jb.off[j][d] + (jb.s0 + kk * SPK + q) * jb.str[j][d]];
        }
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if ((MASK >> (4 * i)) & 0xFu) a[i] = pa[kk * 4 * LDT + 8 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((MASK >> j) & 0x11111111u) b[j] = pb[kk * 4 * LDT + 8 * j];
        if constexpr (BUILD) {
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int q = 0; q < SPK; ++q) {
                    double p = fv[j][q][0];
#pragma unroll
                    for (int d = 1; d < NF; ++d) p *= fv[j][q][d];
                    tnext[jb.dst[j] + (jb.s0 + kk * SPK + q) * LDT] = p;
                }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((MASK >> (i * 4 + j)) & 1u) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// OFFDIAG: every segment is an off-diagonal pair (the moment schedule's launches): only the FULL / FLAT / QUARTER patch
// shapes and the two-tile builder are instantiated -- the same stage bodies, but ptxas allocates the smaller kernel
// better (243 -> 237 registers; C3 Gram 696.3 -> 688.0 ms; a kernel with the FULL shape alone would add 0.3 %).
template <int BT, int NF, bool OFFDIAG = false>
__global__ void __launch_bounds__(GramCfg<BT>::kThreads, GramCfg<BT>::kMinBlocks) gram_kernel(const GramArgs args) {
    using Cfg = GramCfg<BT>;
    constexpr int LDT = Cfg::kPitch;
    constexpr int KS = kKS;
    constexpr int NJ = Cfg::kJobs;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int n = args.n, n_t = args.n_t;
    const uint32_t xbytes = (uint32_t)(KS * n * sizeof(double));
    const uint32_t ybytes = (uint32_t)(KS * n_t * sizeof(double));
    const int vdoubles = (int)(gram_vbytes(n, n_t) / sizeof(double));

    uint64_t* vfull = reinterpret_cast<uint64_t*>(smem_raw);  // [kNVS] raw samples landed
    uint64_t* tfull = vfull + kNVS;                           // [kNTS] Theta stage built by all warps
    double* vring = reinterpret_cast<double*>(smem_raw + 128);
    double* tring = vring + kNVS * vdoubles;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    if (tid == 0) {
        for (int i = 0; i < kNVS; ++i) mbar_init(&vfull[i], 1);
        for (int i = 0; i < kNTS; ++i) mbar_init(&tfull[i], Cfg::kWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < kNVS * kVX; e += Cfg::kThreads) {
        const int w = e % kVX;
        vring[(e / kVX) * vdoubles + w] = (w < kVZero) ? 1.0 : 0.0;
    }
    __syncthreads();

    const int64_t nstages_total = (args.m + KS - 1) / KS;
    const int seg_begin = args.cta_seg_begin[blockIdx.x];
    const int seg_end = args.cta_seg_begin[blockIdx.x + 1];
    int total = 0;
    for (int sg = seg_begin; sg < seg_end; ++sg) total += (int)(args.segs[sg].stage_end - args.segs[sg].stage_begin);
    if (total == 0) return;

    // ---- raw-sample cursor: TMA tiles in item order, across segment boundaries.  Every warp tracks the
    // (uniform) cursor; the warp whose turn it is (item mod #warps) issues the copies, so that no single
    // warp carries the issue cost every stage.
    int tma_seg = seg_begin;
    int tma_stage = (int)args.segs[seg_begin].stage_begin, tma_seg_end = (int)args.segs[seg_begin].stage_end;
    int tma_item = 0;
    uint32_t ragged_slots = 0;  // slots whose ones column currently carries zero padding
    const int ragged_stage = (args.m % KS) ? (int)(nstages_total - 1) : -1;
    auto request_samples = [&]() {  // all warps; requests the raw samples of item `tma_item`
        if (tma_item >= total) return;
        const int slot = tma_item & (kNVS - 1);
        const bool mine = (tma_item % Cfg::kWarps) == warp;
        if (tma_stage != ragged_stage) {
            if ((ragged_slots >> slot) & 1) {  // restore the ones column a ragged stage overwrote
                ragged_slots &= ~(1u << slot);
                if (mine) {
                    if (lane < KS) vring[slot * vdoubles + kVOnes + lane] = 1.0;
                    __syncwarp();
                }
            }
            if (mine && lane == 0) {
                double* dst = vring + slot * vdoubles + kVX;
                const int64_t s0 = (int64_t)tma_stage * KS;
                mbar_expect_tx(&vfull[slot], xbytes + ybytes);
                bulk_g2s(dst, args.X + s0 * n, xbytes, &vfull[slot]);
                bulk_g2s(dst + KS * n, args.Y + s0 * n_t, ybytes, &vfull[slot]);
            }
        } else {
            // ragged last stage of the shard: plain loads with zero fill (a bulk copy needs 16-byte
            // multiples), ones[s] = 0 for the padding samples
            ragged_slots |= 1u << slot;
            if (mine) {
                double* dst = vring + slot * vdoubles;
                const int64_t s0 = (int64_t)tma_stage * KS;
                const int valid = (int)(args.m - s0);
                for (int e = lane; e < KS * n; e += 32) dst[kVX + e] = (e < valid * n) ? args.X[s0 * n + e] : 0.0;
                for (int e = lane; e < KS * n_t; e += 32)
                    dst[kVX + KS * n + e] = (e < valid * n_t) ? args.Y[s0 * n_t + e] : 0.0;
                if (lane < KS) dst[kVOnes + lane] = (lane < valid) ? 1.0 : 0.0;
                __syncwarp();
                if (lane == 0) mbar_arrive(&vfull[slot]);
            }
        }
        ++tma_item;
        if (++tma_stage == tma_seg_end && ++tma_seg < seg_end) {
            tma_stage = (int)args.segs[tma_seg].stage_begin;
            tma_seg_end = (int)args.segs[tma_seg].stage_end;
        }
    };

    // ---- build cursor: this thread's share of the Theta tiles of the NEXT item -----------------------
    int bseg = seg_begin;
    int64_t bstage = 0, bseg_end = 0;
    bool bsame = false;
    BuildJobs<NJ, NF> jb;
    auto load_jobs = [&]() {
        const GramSegment seg = args.segs[bseg];
        bstage = seg.stage_begin;
        bseg_end = seg.stage_end;
        bsame = !OFFDIAG && (seg.bi == seg.bj);
        jb.live = true;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int tile, col;
            if (BT == 128) {  // 256 threads: one column each; diagonal pairs split the samples instead
                tile = bsame ? 0 : tid / BT;
                col = tid % BT;
                jb.s0 = bsame ? (tid / BT) * (KS / 2) : 0;
            } else {  // 64 threads: column tid of the block row (job 0) and of the block column (job 1)
                tile = j;
                col = tid;
                jb.s0 = 0;
            }
            jb.dst[j] = tile * KS * LDT + col;
            // half / quarter pairs of the moment schedule: rows past 64 / 32 of the row tile are never read
            if (BT == 128 && tile == 0 && (((seg.pad & 1) && col >= 64) || ((seg.pad & 2) && col >= 32))) jb.live = false;
            const int term = (tile == 0 ? seg.bi : seg.bj) * BT + col;
            const uint4 q = *reinterpret_cast<const uint4*>(args.factors + (size_t)term * kMaxFactors);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            const bool dead = term >= args.kaug;
#pragma unroll
            for (int d = 0; d < NF; ++d) {
                const uint32_t f = (w[d >> 1] >> (16 * (d & 1))) & 0xFFFFu;
                int off = kVOnes, str = 1;  // missing factor -> ones[s]
                if (f != kNoFactor) {
                    const bool isy = (f & kFactorIsY) != 0;
                    off = isy ? kVX + KS * n + (int)(f & 0x7FFF) : kVX + (int)f;
                    str = isy ? n_t : n;
                }
                if (dead) off = kVZero, str = 0;  // padding row of the library -> 0.0
                jb.off[j][d] = off;
                jb.str[j][d] = str;
            }
        }
    };
    auto advance_build = [&]() {
        if (++bstage == bseg_end && ++bseg < seg_end) load_jobs();
    };

    // ---- prologue: request the first raw tiles, build Theta(0..kAhead-1) completely -------------------
#pragma unroll 1
    for (int i = 0; i < kNVS; ++i) request_samples();
    load_jobs();
#pragma unroll 1
    for (int it = 0; it < kAhead && it < total; ++it) {
        mbar_wait(&vfull[it], 0);
        const double* v = vring + it * vdoubles;
        double* dstt = tring + it * Cfg::kStageDoubles;
        const int njobs = !jb.live ? 0 : (BT == 128 || bsame) ? 1 : 2;
        const int cnt = (BT == 128 && bsame) ? KS / 2 : KS;
        for (int j = 0; j < njobs; ++j)
            for (int s = jb.s0; s < jb.s0 + cnt; ++s) {
                double p = v[jb.off[j][0] + s * jb.str[j][0]];
#pragma unroll
                for (int d = 1; d < NF; ++d) p *= v[jb.off[j][d] + s * jb.str[j][d]];
                dstt[jb.dst[j] + s * LDT] = p;
            }
        advance_build();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tfull[it]);
    }

    // ---- MMA cursor --------------------------------------------------------------------------------------
    // warp -> (wm, wn): permuted so that on diagonal block pairs, where sub-tiles above the diagonal
    // are skipped, the live work is balanced over the four SM sub-partitions.
    int wm, wn;
    if (BT == 128) {
        wm = (0x4B >> warp) & 1;  // warps 0..7 -> (1,0)(1,1)(0,0)(1,2)(0,2)(0,3)(1,3)(0,1)
        wn = (0x7E84 >> (2 * warp)) & 3;
    } else {
        // two warps per CTA, four CTAs per SM: the heavy (26 sub-tiles of the diagonal pair) and the light (10) patch
        // are dealt by the hardware warp slot of warp 0, so that every SM sub-partition (slot mod 4) ends up with one
        // heavy and one light warp whichever CTAs share the SM (args.swap_mode 1; C2: 2.43 -> 2.28 ms)
        __shared__ int s_swap64;
        if (tid == 0) {
            unsigned wid;
            asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
            s_swap64 = args.swap_mode == 1 ? (int)((wid >> 2) & 1u) : 0;
        }
        __syncthreads();
        wm = 0;
        wn = warp ^ s_swap64;
    }
    int item = 0;
    bool ready_t = false, ready_v = false;
    for (int sg = seg_begin; sg < seg_end; ++sg) {
        const GramSegment seg = args.segs[sg];
        const bool same = !OFFDIAG && (seg.bi == seg.bj);
        const int rows_valid = min(BT, args.kaug - seg.bi * BT);
        // block rows with at most 64 live rows: lay all warps side by side (64 x 16 patches)
        // pad bit 0: the plan only needs rows 0..63 of this pair, bit 1: only rows 0..31 (moment schedule)
        const bool flat = (BT == 128) && (rows_valid <= 64 || (seg.pad & 3));
        const int row0 = flat ? 0 : wm * 64;
        const int col0 = flat ? warp * 16 : wn * 32;
        // shape of this warp's patch: 0 full, 1 diagonal offset 0, 2 diagonal offset -4, 3 empty, 4 flat, 5 quarter
        int shape = 0;
        if (flat) shape = (seg.pad & 2) ? 5 : 4;
        else if (same) {
            const int d = (row0 - col0) / 8;
            shape = (d >= 3) ? 0 : (d == 0 ? 1 : (d == -4 ? 2 : 3));
        }
        const uint32_t mask = shape == 0 ? kShapeFull
                              : shape == 1 ? kShapeDiag0
                              : shape == 2 ? kShapeDiagM4
                              : shape == 3 ? kShapeEmpty
                              : shape == 4 ? kShapeFlat
                                           : kShapeQuarter;

        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        const int nst = (int)(seg.stage_end - seg.stage_begin);
        for (int q = 0; q < nst; ++q, ++item) {
            const int tslot = item & (kNTS - 1);
            if (!ready_t) mbar_wait(&tfull[tslot], (uint32_t)((item / kNTS) & 1));
            // the raw tile of item+kNVS reuses the slot of item, which build(item) finished reading
            request_samples();
            const int nitem = item + kAhead;  // the item whose Theta tiles this stage builds
            const bool have_next = (nitem < total);
            if (have_next && !ready_v) mbar_wait(&vfull[nitem & (kNVS - 1)], (uint32_t)((nitem / kNVS) & 1));
            // test the NEXT stage's barriers now; the answers travel back while this stage computes.  A
            // warp that is ahead of the others sees "not yet" and blocks at the top of the next stage,
            // the slowest warp (the one that sets the pace) never waits for the round trip.
            ready_t = (item + 1 < total) ? mbar_test(&tfull[(item + 1) & (kNTS - 1)], (uint32_t)(((item + 1) / kNTS) & 1)) : true;
            ready_v = (nitem + 1 < total) ? mbar_test(&vfull[(nitem + 1) & (kNVS - 1)], (uint32_t)(((nitem + 1) / kNVS) & 1)) : true;

            const double* tI = tring + tslot * Cfg::kStageDoubles;
            const double* tJ = same ? tI : tI + KS * LDT;
            const double* pa = tI + t * LDT + row0 + g;
            const double* pb = tJ + t * LDT + col0 + g;
            const double* v = vring + (nitem & (kNVS - 1)) * vdoubles;          // raw samples of item+kAhead
            double* tnext = tring + (nitem & (kNTS - 1)) * Cfg::kStageDoubles;  // (stale but harmless at the end)
            // the build share belongs to a LATER item, whose pair may differ from this segment's
#define DDB_STAGE(M)                                                                     \
    do {                                                                                 \
        if (!OFFDIAG && bsame) stage_body<BT, NF, M, true>(acc, pa, pb, v, tnext, jb);   \
        else if (jb.live) stage_body<BT, NF, M, false>(acc, pa, pb, v, tnext, jb);       \
        else stage_body<BT, NF, M, false, false>(acc, pa, pb, v, tnext, jb);             \
    } while (0)
            if constexpr (OFFDIAG) {
                switch (shape) {
                    case 0: DDB_STAGE(kShapeFull); break;
                    case 4: DDB_STAGE(kShapeFlat); break;
                    default: DDB_STAGE(kShapeQuarter); break;
                }
            } else {
                switch (shape) {
                    case 0: DDB_STAGE(kShapeFull); break;
                    case 1: DDB_STAGE(kShapeDiag0); break;
                    case 2: DDB_STAGE(kShapeDiagM4); break;
                    case 3: DDB_STAGE(kShapeEmpty); break;
                    case 4: DDB_STAGE(kShapeFlat); break;
                    default: DDB_STAGE(kShapeQuarter); break;
                }
            }
#undef DDB_STAGE
            if (have_next) {
                advance_build();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tfull[nitem & (kNTS - 1)]);
            }
        }

        // flush the live sub-tiles to this segment's partial-tile slot (row-major BT x BT)
        double* out = args.partials + (size_t)seg.slot * BT * BT;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = row0 + 8 * i + g;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = col0 + 8 * j + 2 * t;
                if ((mask >> (i * 4 + j)) & 1)
                    *reinterpret_cast<double2*>(out + (size_t)row * BT + c) = make_double2(acc[i][j][0], acc[i][j][1]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Tensor-pipe builder (moment schedule, virtual terms of at most two factors: degree-2 libraries)
// ---------------------------------------------------------------------------------------------
// An FP64 multiply issued between DMMAs costs the tensor pipe almost a whole DMMA slot (tools/dmma_probe.cu:
// 36.6 -> 32.1 TFLOP/s with the four products per thread and k-step of gram_kernel).  Here the PRODUCTS THEMSELVES
// run on the tensor pipe: one DMMA with a one-hot A operand multiplies 8 first factors v_j with 2 second factors u_r
// for 4 samples,
//     A[(s, r)][k] = u_r(s) if k == s else 0,   B[k][j] = v_j(s_k),   C = 0   ->   D[(s, r)][j] = u_r(s) v_j(s),
// every output one correctly rounded product (all other addends are exact zeros), i.e. 64 library values per
// instruction instead of 32 per DMUL, and no FP64-pipe instruction at all in the loop (probe: 34.1 TFLOP/s with two
// such ops per k-step).  Which (u, v) rectangles cover the terms of a block is the host's job (cover_block in
// gram_moment.cu); the kernel only sees, per warp and stage, up to kMbOpSlots ops as per-lane tables of four 16-bit
// byte offsets: A source and B source inside the raw-sample slot (lanes that hold a zero of the one-hot operand
// point at the slot's 0.0), and the two destinations inside the Theta stage (unused outputs go to the pad columns).
// Library values that are a single raw value (targets, degree-1 terms, the constant) are not multiplied at all:
// up to kMbCopySlots copy slots per warp and stage move one double per lane.
// Everything else (TMA ring, tile ring, barriers, MMA patches, flush) is gram_kernel<128, 2> for off-diagonal pairs.
template <uint32_t MASK>
__device__ __forceinline__ void stage_body_mb(double (&acc)[8][4][2], const double* __restrict__ pa,
                                              const double* __restrict__ pb, const unsigned char* __restrict__ v,
                                              unsigned char* __restrict__ tn, const uint2* __restrict__ wtab,
                                              const uint32_t* __restrict__ ctab, const uint32_t ncopy) {
    constexpr int LDT = GramCfg<128>::kPitch;
    // straight-line code: three ops in k-steps 0..2, two in the last (unused slots hold ops that multiply zeros
    // into the pad columns), so that ptxas can hoist the table -> operand load chains under the DMMAs
#pragma unroll
    for (int kk = 0; kk < kKS / 4; ++kk) {
        constexpr int kMaxOps = 3;
        const int nops = (kk < 3) ? 3 : 2;
        uint2 e[kMaxOps];
        double ua[kMaxOps], vb[kMaxOps], d[kMaxOps][2];
#pragma unroll
        for (int o = 0; o < kMaxOps; ++o)
            if (o < nops) {
                e[o] = wtab[(kk * 3 + o) * 32];
                ua[o] = *reinterpret_cast<const double*>(v + (e[o].x & 0xFFFFu));
                vb[o] = *reinterpret_cast<const double*>(v + (e[o].x >> 16));
            }
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if ((MASK >> (4 * i)) & 0xFu) a[i] = pa[kk * 4 * LDT + 8 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((MASK >> j) & 0x11111111u) b[j] = pb[kk * 4 * LDT + 8 * j];
#pragma unroll
        for (int o = 0; o < kMaxOps; ++o)
            if (o < nops) {
                d[o][0] = d[o][1] = 0.0;
                dmma884(d[o][0], d[o][1], ua[o], vb[o]);
            }
        if (kk == 0) {  // one copy slot is always there (an unused one moves a zero into a pad column)
            const uint32_t c = ctab[0];
            *reinterpret_cast<double*>(tn + (c >> 16)) = *reinterpret_cast<const double*>(v + (c & 0xFFFFu));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((MASK >> (i * 4 + j)) & 1u) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            if (i == 1) {
#pragma unroll
                for (int o = 0; o < kMaxOps; ++o)
                    if (o < nops) {
                        *reinterpret_cast<double*>(tn + (e[o].y & 0xFFFFu)) = d[o][0];
                        *reinterpret_cast<double*>(tn + (e[o].y >> 16)) = d[o][1];
                    }
            }
        }
    }
    if (ncopy > 1) {  // warp-uniform; only pairs with a block of targets have more than one copy slot per warp
#pragma unroll 1
        for (uint32_t q = 1; q < ncopy; ++q) {
            const uint32_t c = ctab[q * 32];
            *reinterpret_cast<double*>(tn + (c >> 16)) = *reinterpret_cast<const double*>(v + (c & 0xFFFFu));
        }
    }
}

inline size_t gram_mb_smem_bytes(int n, int n_t) {
    return 128 + kNVS * gram_mb_vbytes(n, n_t) + (size_t)kNTS * GramCfg<128>::kStageDoubles * sizeof(double) +
           (size_t)kMbPairWords * sizeof(uint32_t);
}

__global__ void __launch_bounds__(GramCfg<128>::kThreads, 1) gram_mb_kernel(const GramArgs args) {
    using Cfg = GramCfg<128>;
    constexpr int BT = 128;
    constexpr int LDT = Cfg::kPitch;
    constexpr int KS = kKS;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int n = args.n, n_t = args.n_t;
    const uint32_t xrow_bytes = (uint32_t)(n * sizeof(double));
    const uint32_t ybytes = (uint32_t)(KS * n_t * sizeof(double));
    const int px = mb_pitch(n);
    const int vdoubles = (int)(gram_mb_vbytes(n, n_t) / sizeof(double));

    uint64_t* vfull = reinterpret_cast<uint64_t*>(smem_raw);  // [kNVS] raw samples landed
    uint64_t* tfull = vfull + kNVS;                           // [kNTS] Theta stage built by all warps
    double* vring = reinterpret_cast<double*>(smem_raw + 128);
    double* tring = vring + kNVS * vdoubles;
    uint32_t* otab = reinterpret_cast<uint32_t*>(tring + kNTS * Cfg::kStageDoubles);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    uint32_t* wwords = otab + warp * kMbWarpWords;                     // this warp's tables
    const uint2* wtab = reinterpret_cast<const uint2*>(wwords) + lane;  // ops, this lane's column
    const uint32_t* ctab = wwords + kMbOpSlots * 64 + lane;             // copy slots

    if (tid == 0) {
        for (int i = 0; i < kNVS; ++i) mbar_init(&vfull[i], 1);
        for (int i = 0; i < kNTS; ++i) mbar_init(&tfull[i], Cfg::kWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < kNVS * kVX; e += Cfg::kThreads) {
        const int w = e % kVX;
        vring[(e / kVX) * vdoubles + w] = (w < kVZero) ? 1.0 : 0.0;
    }
    // columns no op writes (terms past the end of the virtual library) stay 0.0 for the whole launch
    for (int e = tid; e < kNTS * Cfg::kStageDoubles; e += Cfg::kThreads) tring[e] = 0.0;
    __syncthreads();

    const int64_t nstages_total = (args.m + KS - 1) / KS;
    const int seg_begin = args.cta_seg_begin[blockIdx.x];
    const int seg_end = args.cta_seg_begin[blockIdx.x + 1];
    int total = 0;
    for (int sg = seg_begin; sg < seg_end; ++sg) total += (int)(args.segs[sg].stage_end - args.segs[sg].stage_begin);
    if (total == 0) return;

    // ---- raw-sample cursor (as gram_kernel) -----------------------------------------------------------------
    int tma_seg = seg_begin;
    int tma_stage = (int)args.segs[seg_begin].stage_begin, tma_seg_end = (int)args.segs[seg_begin].stage_end;
    int tma_item = 0;
    uint32_t ragged_slots = 0;
    const int ragged_stage = (args.m % KS) ? (int)(nstages_total - 1) : -1;
    auto request_samples = [&]() {
        if (tma_item >= total) return;
        const int slot = tma_item & (kNVS - 1);
        const bool mine = (tma_item % Cfg::kWarps) == warp;
        if (tma_stage != ragged_stage) {
            if ((ragged_slots >> slot) & 1) {
                ragged_slots &= ~(1u << slot);
                if (mine) {
                    if (lane < KS) vring[slot * vdoubles + kVOnes + lane] = 1.0;
                    __syncwarp();
                }
            }
            if (mine) {  // the X tile row by row (padded pitch), the Y tile in one piece
                double* dst = vring + slot * vdoubles + kVX;
                const int64_t s0 = (int64_t)tma_stage * KS;
                if (lane == 0) mbar_expect_tx(&vfull[slot], KS * xrow_bytes + ybytes);
                __syncwarp();
                if (lane < KS) bulk_g2s(dst + lane * px, args.X + (s0 + lane) * n, xrow_bytes, &vfull[slot]);
                if (lane == 0) bulk_g2s(dst + KS * px, args.Y + s0 * n_t, ybytes, &vfull[slot]);
            }
        } else {
            ragged_slots |= 1u << slot;
            if (mine) {
                double* dst = vring + slot * vdoubles;
                const int64_t s0 = (int64_t)tma_stage * KS;
                const int valid = (int)(args.m - s0);
                for (int e = lane; e < KS * n; e += 32)
                    dst[kVX + (e / n) * px + e % n] = (e < valid * n) ? args.X[s0 * n + e] : 0.0;
                for (int e = lane; e < KS * n_t; e += 32)
                    dst[kVX + KS * px + e] = (e < valid * n_t) ? args.Y[s0 * n_t + e] : 0.0;
                if (lane < KS) dst[kVOnes + lane] = (lane < valid) ? 1.0 : 0.0;
                __syncwarp();
                if (lane == 0) mbar_arrive(&vfull[slot]);
            }
        }
        ++tma_item;
        if (++tma_stage == tma_seg_end && ++tma_seg < seg_end) {
            tma_stage = (int)args.segs[tma_seg].stage_begin;
            tma_seg_end = (int)args.segs[tma_seg].stage_end;
        }
    };

    // ---- build cursor: this warp's op table of the pair whose Theta tiles are built next -----------------------
    int bseg = seg_begin;
    int64_t bstage = 0, bseg_end = 0;
    uint32_t bcnt = 0;  // copy slots in use
    auto load_ops = [&]() {
        const GramSegment seg = args.segs[bseg];
        bstage = seg.stage_begin;
        bseg_end = seg.stage_end;
        const int pair = seg.pad >> 8;
        const uint32_t* src = args.optab + (size_t)pair * kMbPairWords + warp * kMbWarpWords;
        __syncwarp();
#pragma unroll 1
        for (int e = lane; e < kMbWarpWords; e += 32) wwords[e] = src[e];
        bcnt = args.opcnt[pair * Cfg::kWarps + warp] >> 4;
        __syncwarp();
    };
    auto advance_build = [&]() {
        if (++bstage == bseg_end && ++bseg < seg_end) load_ops();
    };

    // ---- prologue: request the first raw tiles, build Theta(0..kAhead-1) completely ------------------------------
#pragma unroll 1
    for (int i = 0; i < kNVS; ++i) request_samples();
    load_ops();
#pragma unroll 1
    for (int it = 0; it < kAhead && it < total; ++it) {
        mbar_wait(&vfull[it], 0);
        const unsigned char* v = reinterpret_cast<const unsigned char*>(vring + it * vdoubles);
        unsigned char* tn = reinterpret_cast<unsigned char*>(tring + it * Cfg::kStageDoubles);
#pragma unroll 1
        for (int q = 0; q < kMbOpSlots; ++q) {
            const uint2 e = wtab[q * 32];
            const double ua = *reinterpret_cast<const double*>(v + (e.x & 0xFFFFu));
            const double vb = *reinterpret_cast<const double*>(v + (e.x >> 16));
            double d0 = 0.0, d1 = 0.0;
            dmma884(d0, d1, ua, vb);
            *reinterpret_cast<double*>(tn + (e.y & 0xFFFFu)) = d0;
            *reinterpret_cast<double*>(tn + (e.y >> 16)) = d1;
        }
#pragma unroll 1
        for (int q = 0; q < (int)max(bcnt, 1u); ++q) {
            const uint32_t c = ctab[q * 32];
            *reinterpret_cast<double*>(tn + (c >> 16)) = *reinterpret_cast<const double*>(v + (c & 0xFFFFu));
        }
        advance_build();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tfull[it]);
    }

    // ---- MMA cursor ---------------------------------------------------------------------------------------------
    const int wm = (0x4B >> warp) & 1;  // as gram_kernel
    const int wn = (0x7E84 >> (2 * warp)) & 3;
    int item = 0;
    bool ready_t = false, ready_v = false;
    for (int sg = seg_begin; sg < seg_end; ++sg) {
        const GramSegment seg = args.segs[sg];
        const bool flat = (seg.pad & 3) != 0;
        const int row0 = flat ? 0 : wm * 64;
        const int col0 = flat ? warp * 16 : wn * 32;
        const int shape = !flat ? 0 : (seg.pad & 2) ? 2 : 1;
        const uint32_t mask = shape == 0 ? kShapeFull : shape == 1 ? kShapeFlat : kShapeQuarter;

        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        const int nst = (int)(seg.stage_end - seg.stage_begin);
        for (int q = 0; q < nst; ++q, ++item) {
            const int tslot = item & (kNTS - 1);
            if (!ready_t) mbar_wait(&tfull[tslot], (uint32_t)((item / kNTS) & 1));
            request_samples();
            const int nitem = item + kAhead;
            const bool have_next = (nitem < total);
            if (have_next && !ready_v) mbar_wait(&vfull[nitem & (kNVS - 1)], (uint32_t)((nitem / kNVS) & 1));
            ready_t = (item + 1 < total) ? mbar_test(&tfull[(item + 1) & (kNTS - 1)], (uint32_t)(((item + 1) / kNTS) & 1)) : true;
            ready_v = (nitem + 1 < total) ? mbar_test(&vfull[(nitem + 1) & (kNVS - 1)], (uint32_t)(((nitem + 1) / kNVS) & 1)) : true;

            const double* tI = tring + tslot * Cfg::kStageDoubles;
            const double* tJ = tI + KS * LDT;
            const double* pa = tI + t * LDT + row0 + g;
            const double* pb = tJ + t * LDT + col0 + g;
            // past the last item the ops run once more on stale data into a stage nobody reads (harmless)
            const unsigned char* v = reinterpret_cast<const unsigned char*>(vring + (nitem & (kNVS - 1)) * vdoubles);
            unsigned char* tn = reinterpret_cast<unsigned char*>(tring + (nitem & (kNTS - 1)) * Cfg::kStageDoubles);
            switch (shape) {
                case 0: stage_body_mb<kShapeFull>(acc, pa, pb, v, tn, wtab, ctab, bcnt); break;
                case 1: stage_body_mb<kShapeFlat>(acc, pa, pb, v, tn, wtab, ctab, bcnt); break;
                default: stage_body_mb<kShapeQuarter>(acc, pa, pb, v, tn, wtab, ctab, bcnt); break;
            }
            if (have_next) {
                advance_build();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tfull[nitem & (kNTS - 1)]);
            }
        }

        double* out = args.partials + (size_t)seg.slot * BT * BT;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = row0 + 8 * i + g;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = col0 + 8 * j + 2 * t;
                if ((mask >> (i * 4 + j)) & 1)
                    *reinterpret_cast<double2*>(out + (size_t)row * BT + c) = make_double2(acc[i][j][0], acc[i][j][1]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Prefix-product builder (libraries of at most 64 augmented rows that are closed under "drop the last factor")
// ---------------------------------------------------------------------------------------------
// gram_kernel<64, NF> multiplies every library value out from its NF staged factors: NF - 1 FP64 multiplies per
// value, each costing the tensor pipe most of a DMMA slot (C2: 4 multiplies and 5 loads per value against 2.25
// DMMAs).  A term is its parent (the term without its last factor) times that factor: ONE multiply per value, the
// same association as the left-to-right product, i.e. the same bits.  The dependency is kept inside a thread: thread
// (sample s, quarter q) walks program q -- a contiguous piece of the depth-first order of the term tree, preceded by
// the ancestors of its first node -- through row s of the tile; a parent is always a value the same thread stored
// earlier (program order, no barrier).  Ancestors shared by two quarters are stored by both with identical bits.
// Everything else is gram_kernel<64> on the single diagonal pair.
constexpr int kTreeQuarters = 4;
inline size_t gram_tree_smem_bytes(int n, int n_t, int nsteps) {
    return gram_smem_bytes<64>(n, n_t) + (size_t)64 * nsteps * sizeof(uint2);
}

// Steps of a k-step form four groups (one per pair of sub-tile rows); the host guarantees that a step's parent is
// stored by an EARLIER group, so a group's loads are issued together, ahead of its multiplies and stores.
template <int G>
struct TreeGroup {
    uint2 e[G > 0 ? G : 1];
    double par[G > 0 ? G : 1], fac[G > 0 ? G : 1];
    __device__ __forceinline__ void load(const uint2* __restrict__ prog, int first, const unsigned char* __restrict__ v,
                                         const unsigned char* tn) {
#pragma unroll
        for (int j = 0; j < G; ++j) {
            e[j] = prog[(first + j) * 64];
            const unsigned char* pbase = (e[j].y >> 31) ? v : tn;  // no parent: the slot's ones column
            par[j] = *reinterpret_cast<const double*>(pbase + (e[j].x >> 16));
            fac[j] = *reinterpret_cast<const double*>(v + (e[j].y & 0xFFFFu));
        }
    }
    __device__ __forceinline__ void store(unsigned char* tn) const {
#pragma unroll
        for (int j = 0; j < G; ++j) *reinterpret_cast<double*>(tn + (e[j].x & 0xFFFFu)) = par[j] * fac[j];
    }
};

__host__ __device__ constexpr int tree_group_size(int spk, int g) { return spk / 4 + (g < spk % 4 ? 1 : 0); }
__host__ __device__ constexpr int tree_group_first(int spk, int g) {
    int f = 0;
    for (int h = 0; h < g; ++h) f += tree_group_size(spk, h);
    return f;
}

template <int NSTEP, uint32_t MASK>
__device__ __forceinline__ void stage_body_tree(double (&acc)[8][4][2], const double* __restrict__ pa,
                                                const double* __restrict__ pb, const unsigned char* __restrict__ v,
                                                unsigned char* tn, const uint2* __restrict__ prog) {
    constexpr int LDT = GramCfg<64>::kPitch;
    constexpr int SPK = NSTEP / 4;  // program steps per k-step, in four groups (one per pair of sub-tile rows)
#pragma unroll
    for (int kk = 0; kk < kKS / 4; ++kk) {
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if ((MASK >> (4 * i)) & 0xFu) a[i] = pa[kk * 4 * LDT + 8 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((MASK >> j) & 0x11111111u) b[j] = pb[kk * 4 * LDT + 8 * j];
        auto rows = [&](int i0) {
#pragma unroll
            for (int i = i0; i < i0 + 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((MASK >> (i * 4 + j)) & 1u) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        };
        {
            TreeGroup<tree_group_size(SPK, 0)> grp;
            grp.load(prog, kk * SPK + tree_group_first(SPK, 0), v, tn);
            rows(0);
            grp.store(tn);
        }
        {
            TreeGroup<tree_group_size(SPK, 1)> grp;
            grp.load(prog, kk * SPK + tree_group_first(SPK, 1), v, tn);
            rows(2);
            grp.store(tn);
        }
        {
            TreeGroup<tree_group_size(SPK, 2)> grp;
            grp.load(prog, kk * SPK + tree_group_first(SPK, 2), v, tn);
            rows(4);
            grp.store(tn);
        }
        {
            TreeGroup<tree_group_size(SPK, 3)> grp;
            grp.load(prog, kk * SPK + tree_group_first(SPK, 3), v, tn);
            rows(6);
            grp.store(tn);
        }
    }
}

template <int NSTEP>
__global__ void __launch_bounds__(64, 4) gram_tree_kernel(const GramArgs args) {
    using Cfg = GramCfg<64>;
    constexpr int BT = 64;
    constexpr int LDT = Cfg::kPitch;
    constexpr int KS = kKS;
    static_assert(NSTEP % 4 == 0 && NSTEP / 4 <= 8, "steps are spread over the 8 sub-tile rows of a k-step");
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int n = args.n, n_t = args.n_t;
    const uint32_t xbytes = (uint32_t)(KS * n * sizeof(double));
    const uint32_t ybytes = (uint32_t)(KS * n_t * sizeof(double));
    const int vdoubles = (int)(gram_vbytes(n, n_t) / sizeof(double));

    uint64_t* vfull = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* tfull = vfull + kNVS;
    double* vring = reinterpret_cast<double*>(smem_raw + 128);
    double* tring = vring + kNVS * vdoubles;
    uint2* progs = reinterpret_cast<uint2*>(tring + kNTS * Cfg::kStageDoubles);  // [NSTEP][64 threads]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    if (tid == 0) {
        for (int i = 0; i < kNVS; ++i) mbar_init(&vfull[i], 1);
        for (int i = 0; i < kNTS; ++i) mbar_init(&tfull[i], Cfg::kWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < kNVS * kVX; e += Cfg::kThreads) {
        const int w = e % kVX;
        vring[(e / kVX) * vdoubles + w] = (w < kVZero) ? 1.0 : 0.0;
    }
    // rows past the end of the library: no step writes them
    for (int e = tid; e < kNTS * Cfg::kStageDoubles; e += Cfg::kThreads) tring[e] = 0.0;
    {   // this thread's program, resolved for its sample: byte offsets inside a Theta stage / a raw-sample slot
        const int s = tid & 15, q = tid >> 4;
        for (int j = 0; j < NSTEP; ++j) {
            const uint32_t w = args.tree[q * NSTEP + j];
            const uint32_t dst = w & 0xFFu, par = (w >> 8) & 0xFFu, code = w >> 16;
            uint32_t foff;
            if (code == kNoFactor) foff = (uint32_t)(kVOnes + s) * 8;
            else if (code & kFactorIsY) foff = (uint32_t)(kVX + KS * n + s * n_t + (int)(code & 0x7FFF)) * 8;
            else foff = (uint32_t)(kVX + s * n + (int)code) * 8;
            const bool root = (par == 0xFFu);
            const uint32_t poff = root ? (uint32_t)(kVOnes + s) * 8 : (uint32_t)(s * LDT + (int)par) * 8;
            progs[j * 64 + tid] = make_uint2((uint32_t)(s * LDT + (int)dst) * 8 | (poff << 16), foff | (root ? 0x80000000u : 0u));
        }
    }
    __syncthreads();
    const uint2* prog = progs + tid;

    const int64_t nstages_total = (args.m + KS - 1) / KS;
    const int seg_begin = args.cta_seg_begin[blockIdx.x];
    const int seg_end = args.cta_seg_begin[blockIdx.x + 1];
    int total = 0;
    for (int sg = seg_begin; sg < seg_end; ++sg) total += (int)(args.segs[sg].stage_end - args.segs[sg].stage_begin);
    if (total == 0) return;

    int tma_seg = seg_begin;
    int tma_stage = (int)args.segs[seg_begin].stage_begin, tma_seg_end = (int)args.segs[seg_begin].stage_end;
    int tma_item = 0;
    uint32_t ragged_slots = 0;
    const int ragged_stage = (args.m % KS) ? (int)(nstages_total - 1) : -1;
    auto request_samples = [&]() {
        if (tma_item >= total) return;
        const int slot = tma_item & (kNVS - 1);
        const bool mine = (tma_item % Cfg::kWarps) == warp;
        if (tma_stage != ragged_stage) {
            if ((ragged_slots >> slot) & 1) {
                ragged_slots &= ~(1u << slot);
                if (mine) {
                    if (lane < KS) vring[slot * vdoubles + kVOnes + lane] = 1.0;
                    __syncwarp();
                }
            }
            if (mine && lane == 0) {
                double* dst = vring + slot * vdoubles + kVX;
                const int64_t s0 = (int64_t)tma_stage * KS;
                mbar_expect_tx(&vfull[slot], xbytes + ybytes);
                bulk_g2s(dst, args.X + s0 * n, xbytes, &vfull[slot]);
                bulk_g2s(dst + KS * n, args.Y + s0 * n_t, ybytes, &vfull[slot]);
            }
        } else {
            ragged_slots |= 1u << slot;
            if (mine) {
                double* dst = vring + slot * vdoubles;
                const int64_t s0 = (int64_t)tma_stage * KS;
                const int valid = (int)(args.m - s0);
                for (int e = lane; e < KS * n; e += 32) dst[kVX + e] = (e < valid * n) ? args.X[s0 * n + e] : 0.0;
                for (int e = lane; e < KS * n_t; e += 32)
                    dst[kVX + KS * n + e] = (e < valid * n_t) ? args.Y[s0 * n_t + e] : 0.0;
                if (lane < KS) dst[kVOnes + lane] = (lane < valid) ? 1.0 : 0.0;
                __syncwarp();
                if (lane == 0) mbar_arrive(&vfull[slot]);
            }
        }
        ++tma_item;
        if (++tma_stage == tma_seg_end && ++tma_seg < seg_end) {
            tma_stage = (int)args.segs[tma_seg].stage_begin;
            tma_seg_end = (int)args.segs[tma_seg].stage_end;
        }
    };

    // ---- prologue: the first raw tiles, Theta(0..kAhead-1) completely ---------------------------------------------
#pragma unroll 1
    for (int i = 0; i < kNVS; ++i) request_samples();
#pragma unroll 1
    for (int it = 0; it < kAhead && it < total; ++it) {
        mbar_wait(&vfull[it], 0);
        const unsigned char* v = reinterpret_cast<const unsigned char*>(vring + it * vdoubles);
        unsigned char* tn = reinterpret_cast<unsigned char*>(tring + it * Cfg::kStageDoubles);
#pragma unroll 1
        for (int j = 0; j < NSTEP; ++j) {
            const uint2 e = prog[j * 64];
            const unsigned char* pbase = (e.y >> 31) ? v : tn;
            const double par = *reinterpret_cast<const volatile double*>(pbase + (e.x >> 16));
            const double fac = *reinterpret_cast<const double*>(v + (e.y & 0xFFFFu));
            *reinterpret_cast<volatile double*>(tn + (e.x & 0xFFFFu)) = par * fac;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&tfull[it]);
    }

    // ---- MMA cursor: warp w owns columns 32 w .. 32 w + 31 of the diagonal pair --------------------------------------
    const int col0 = warp * 32;
    const int shape = (warp == 0) ? 1 : 2;  // diagonal offset 0 / -4
    const uint32_t mask = shape == 1 ? kShapeDiag0 : kShapeDiagM4;
    int item = 0;
    bool ready_t = false, ready_v = false;
    for (int sg = seg_begin; sg < seg_end; ++sg) {
        const GramSegment seg = args.segs[sg];
        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        const int nst = (int)(seg.stage_end - seg.stage_begin);
        for (int q = 0; q < nst; ++q, ++item) {
            const int tslot = item & (kNTS - 1);
            if (!ready_t) mbar_wait(&tfull[tslot], (uint32_t)((item / kNTS) & 1));
            request_samples();
            const int nitem = item + kAhead;
            const bool have_next = (nitem < total);
            if (have_next && !ready_v) mbar_wait(&vfull[nitem & (kNVS - 1)], (uint32_t)((nitem / kNVS) & 1));
            ready_t = (item + 1 < total) ? mbar_test(&tfull[(item + 1) & (kNTS - 1)], (uint32_t)(((item + 1) / kNTS) & 1)) : true;
            ready_v = (nitem + 1 < total) ? mbar_test(&vfull[(nitem + 1) & (kNVS - 1)], (uint32_t)(((nitem + 1) / kNVS) & 1)) : true;

            const double* tI = tring + tslot * Cfg::kStageDoubles;
            const double* pa = tI + t * LDT + g;
            const double* pb = tI + t * LDT + col0 + g;
            const unsigned char* v = reinterpret_cast<const unsigned char*>(vring + (nitem & (kNVS - 1)) * vdoubles);
            unsigned char* tn = reinterpret_cast<unsigned char*>(tring + (nitem & (kNTS - 1)) * Cfg::kStageDoubles);
            if (shape == 1) stage_body_tree<NSTEP, kShapeDiag0>(acc, pa, pb, v, tn, prog);
            else stage_body_tree<NSTEP, kShapeDiagM4>(acc, pa, pb, v, tn, prog);
            if (have_next) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&tfull[nitem & (kNTS - 1)]);
            }
        }

        double* out = args.partials + (size_t)seg.slot * BT * BT;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = 8 * i + g;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = col0 + 8 * j + 2 * t;
                if ((mask >> (i * 4 + j)) & 1)
                    *reinterpret_cast<double2*>(out + (size_t)row * BT + c) = make_double2(acc[i][j][0], acc[i][j][1]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Power-table builder (libraries of at most 64 augmented rows over at most 3 states, degree >= 3: config C2)
// ---------------------------------------------------------------------------------------------
// gram_kernel<64, NF> multiplies every library value out from NF = max degree staged states (C2: 4 multiplies and 5
// loads per value against 2.25 DMMAs, and every FP64 multiply costs the tensor pipe most of a DMMA slot).  A monomial is
// also the product of ONE power per state, x^a y^b z^c: at most min(n, degree) factors.  Here every stage gets a table
// of powers [ones(16) | 0 | pad | x_i^e for 16 samples x (n states x deg powers) | Y tile] -- the raw-sample slot of a
// pseudo problem with n * deg states -- and the ordinary builder (stage_body<64, NF, ., true>) reads it with NF =
// min(n, degree): C2 2 multiplies and 3 loads per value.  The table of stage i + 4 is written while stage i is
// contracted (thread = (sample, state): deg - 1 chained multiplies; the fourth group of threads copies the targets and
// the ones column), two stages before the builder reads it, so the "tile built" barrier every warp passes in between
// orders the writes before the reads and no barrier is added; the raw ring is 8 deep for that.  Powers are left-to-right
// products (x^e has the bits of the factor-by-factor product); mixed terms associate as (x^a)(y^b)(z^c), so the blocks
// differ from gram_kernel<64>'s in the last bits.  Everything else is gram_kernel<64> on the single diagonal pair.
constexpr int kPowNVS = 8;  // depth of the raw-sample ring
constexpr int kPowNPS = 4;  // depth of the power-table ring
inline size_t gram_pow_smem_bytes(int n, int n_t, int deg) {
    return 128 + kPowNVS * gram_vbytes(n, n_t) + kPowNPS * gram_vbytes(n * deg, n_t) +
           (size_t)kNTS * GramCfg<64>::kStageDoubles * sizeof(double);
}

// (the kernel itself follows gram_w4_kernel, whose warp layout and stage body it shares)

// ---------------------------------------------------------------------------------------------
// Four balanced warps on the single 64 x 64 diagonal pair (libraries of at most 64 augmented rows: config C2)
// ---------------------------------------------------------------------------------------------
// gram_kernel<64, NF> gives its two warps the column halves of the pair: 26 and 10 live sub-tiles.  With four CTAs on an
// SM that is two in-order warps per sub-partition, one of them with a quarter of the work, and what limits the kernel
// is the latency each warp exposes, not the multiply count (three builders with 4, 2 and 1 multiplies per value all
// run in the same time).  Here the 36 live 8 x 8 sub-tiles of the lower triangle are dealt to FOUR warps by sub-tile
// rows: warp w owns rows 7 - w (8 - w sub-tiles) and w (w + 1 sub-tiles) -- 9 sub-tiles each whatever w is, 36
// accumulator registers instead of 104, so four CTAs of 128 threads share an SM: four warps per sub-partition, all with
// the same load.  Builder: thread = (column, half of the stage's 16 samples).  Rings, barriers and flush as in
// gram_kernel<64>.
template <int NF, int W>
__device__ __forceinline__ void stage_body_w4(double (&accA)[8][2], double (&accB)[4][2], const double* __restrict__ tI,
                                              const double* __restrict__ v, double* __restrict__ tnext,
                                              const int (&off)[NF], const int (&str)[NF], int dst, int g, int t) {
    // off / dst already point at this thread's first sample (the callers fold s0 in): the sample index below is a
    // compile-time constant
    constexpr int LDT = GramCfg<64>::kPitch;
    constexpr int RA = 7 - W, RB = W;
#pragma unroll
    for (int kk = 0; kk < kKS / 4; ++kk) {
        double fv[2][NF];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int d = 0; d < NF; ++d) fv[q][d] = v[off[d] + (kk * 2 + q) * str[d]];
        const double* row = tI + (kk * 4 + t) * LDT + g;
        const double aA = row[8 * RA], aB = row[8 * RB];
        double b[8];
#pragma unroll
        for (int j = 0; j <= RA; ++j) b[j] = row[8 * j];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            double p = fv[q][0];
#pragma unroll
            for (int d = 1; d < NF; ++d) p *= fv[q][d];
            tnext[dst + (kk * 2 + q) * LDT] = p;
        }
#pragma unroll
        for (int j = 0; j <= RA; ++j) dmma884(accA[j][0], accA[j][1], aA, b[j]);
#pragma unroll
        for (int j = 0; j <= RB; ++j) dmma884(accB[j][0], accB[j][1], aB, b[j]);
    }
}

constexpr int kW4Threads = 128;
template <int NF, int MB = 4>
__global__ void __launch_bounds__(kW4Threads, MB) gram_w4_kernel(const GramArgs args) {
    using Cfg = GramCfg<64>;
    constexpr int BT = 64;
    constexpr int LDT = Cfg::kPitch;
    constexpr int KS = kKS;
    constexpr int NW = kW4Threads / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int n = args.n, n_t = args.n_t;
    const uint32_t xbytes = (uint32_t)(KS * n * sizeof(double));
    const uint32_t ybytes = (uint32_t)(KS * n_t * sizeof(double));
    const int vdoubles = (int)(gram_vbytes(n, n_t) / sizeof(double));

    uint64_t* vfull = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* tfull = vfull + kNVS;
    double* vring = reinterpret_cast<double*>(smem_raw + 128);
    double* tring = vring + kNVS * vdoubles;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    if (tid == 0) {
        for (int i = 0; i < kNVS; ++i) mbar_init(&vfull[i], 1);
        for (int i = 0; i < kNTS; ++i) mbar_init(&tfull[i], NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < kNVS * kVX; e += kW4Threads) {
        const int w = e % kVX;
        vring[(e / kVX) * vdoubles + w] = (w < kVZero) ? 1.0 : 0.0;
    }
    __syncthreads();

    const int64_t nstages_total = (args.m + KS - 1) / KS;
    const int seg_begin = args.cta_seg_begin[blockIdx.x];
    const int seg_end = args.cta_seg_begin[blockIdx.x + 1];
    int total = 0;
    for (int sg = seg_begin; sg < seg_end; ++sg) total += (int)(args.segs[sg].stage_end - args.segs[sg].stage_begin);
    if (total == 0) return;

    int tma_seg = seg_begin;
    int tma_stage = (int)args.segs[seg_begin].stage_begin, tma_seg_end = (int)args.segs[seg_begin].stage_end;
    int tma_item = 0;
    uint32_t ragged_slots = 0;
    const int ragged_stage = (args.m % KS) ? (int)(nstages_total - 1) : -1;
    auto request_samples = [&]() {
        if (tma_item >= total) return;
        const int slot = tma_item & (kNVS - 1);
        const bool mine = (tma_item % NW) == warp;
        if (tma_stage != ragged_stage) {
            if ((ragged_slots >> slot) & 1) {
                ragged_slots &= ~(1u << slot);
                if (mine) {
                    if (lane < KS) vring[slot * vdoubles + kVOnes + lane] = 1.0;
                    __syncwarp();
                }
            }
            if (mine && lane == 0) {
                double* dst = vring + slot * vdoubles + kVX;
                const int64_t s0 = (int64_t)tma_stage * KS;
                mbar_expect_tx(&vfull[slot], xbytes + ybytes);
                bulk_g2s(dst, args.X + s0 * n, xbytes, &vfull[slot]);
                bulk_g2s(dst + KS * n, args.Y + s0 * n_t, ybytes, &vfull[slot]);
            }
        } else {
            ragged_slots |= 1u << slot;
            if (mine) {
                double* dst = vring + slot * vdoubles;
                const int64_t s0 = (int64_t)tma_stage * KS;
                const int valid = (int)(args.m - s0);
                for (int e = lane; e < KS * n; e += 32) dst[kVX + e] = (e < valid * n) ? args.X[s0 * n + e] : 0.0;
                for (int e = lane; e < KS * n_t; e += 32)
                    dst[kVX + KS * n + e] = (e < valid * n_t) ? args.Y[s0 * n_t + e] : 0.0;
                if (lane < KS) dst[kVOnes + lane] = (lane < valid) ? 1.0 : 0.0;
                __syncwarp();
                if (lane == 0) mbar_arrive(&vfull[slot]);
            }
        }
        ++tma_item;
        if (++tma_stage == tma_seg_end && ++tma_seg < seg_end) {
            tma_stage = (int)args.segs[tma_seg].stage_begin;
            tma_seg_end = (int)args.segs[tma_seg].stage_end;
        }
    };

    // ---- this thread's share of a Theta tile: column tid mod 64, samples (tid / 64) * 8 .. + 7 (all segments of the plan
    // are pair (0, 0))
    int off[NF], str[NF];
    const int dst = tid & 63, s0 = (tid >> 6) * (KS / 2);
    {
        const int term = tid & 63;
        const uint4 q = *reinterpret_cast<const uint4*>(args.factors + (size_t)term * kMaxFactors);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        const bool dead = term >= args.kaug;
#pragma unroll
        for (int d = 0; d < NF; ++d) {
            const uint32_t f = (w[d >> 1] >> (16 * (d & 1))) & 0xFFFFu;
            int o = kVOnes, st = 1;
            if (f != kNoFactor) {
                const bool isy = (f & kFactorIsY) != 0;
                o = isy ? kVX + KS * n + (int)(f & 0x7FFF) : kVX + (int)f;
                st = isy ? n_t : n;
            }
            if (dead) o = kVZero, st = 0;
            off[d] = o, str[d] = st;
        }
    }
    int off0[NF];
#pragma unroll
    for (int d = 0; d < NF; ++d) off0[d] = off[d] + s0 * str[d];
    const int dst0 = dst + s0 * LDT;

#pragma unroll 1
    for (int i = 0; i < kNVS; ++i) request_samples();
#pragma unroll 1
    for (int it = 0; it < kAhead && it < total; ++it) {
        mbar_wait(&vfull[it], 0);
        const double* v = vring + it * vdoubles;
        double* dstt = tring + it * Cfg::kStageDoubles;
        for (int s = s0; s < s0 + KS / 2; ++s) {
            double p = v[off[0] + s * str[0]];
#pragma unroll
            for (int d = 1; d < NF; ++d) p *= v[off[d] + s * str[d]];
            dstt[dst + s * LDT] = p;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&tfull[it]);
    }

    int item = 0;
    bool ready_t = false, ready_v = false;
    for (int sg = seg_begin; sg < seg_end; ++sg) {
        const GramSegment seg = args.segs[sg];
        double accA[8][2], accB[4][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) accA[j][0] = accA[j][1] = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) accB[j][0] = accB[j][1] = 0.0;

        const int nst = (int)(seg.stage_end - seg.stage_begin);
        for (int q = 0; q < nst; ++q, ++item) {
            const int tslot = item & (kNTS - 1);
            if (!ready_t) mbar_wait(&tfull[tslot], (uint32_t)((item / kNTS) & 1));
            request_samples();
            const int nitem = item + kAhead;
            const bool have_next = (nitem < total);
            if (have_next && !ready_v) mbar_wait(&vfull[nitem & (kNVS - 1)], (uint32_t)((nitem / kNVS) & 1));
            ready_t = (item + 1 < total) ? mbar_test(&tfull[(item + 1) & (kNTS - 1)], (uint32_t)(((item + 1) / kNTS) & 1)) : true;
            ready_v = (nitem + 1 < total) ? mbar_test(&vfull[(nitem + 1) & (kNVS - 1)], (uint32_t)(((nitem + 1) / kNVS) & 1)) : true;

            const double* tI = tring + tslot * Cfg::kStageDoubles;
            const double* v = vring + (nitem & (kNVS - 1)) * vdoubles;
            double* tnext = tring + (nitem & (kNTS - 1)) * Cfg::kStageDoubles;  // (stale but harmless at the end)
            switch (warp) {
                case 0: stage_body_w4<NF, 0>(accA, accB, tI, v, tnext, off0, str, dst0, g, t); break;
                case 1: stage_body_w4<NF, 1>(accA, accB, tI, v, tnext, off0, str, dst0, g, t); break;
                case 2: stage_body_w4<NF, 2>(accA, accB, tI, v, tnext, off0, str, dst0, g, t); break;
                default: stage_body_w4<NF, 3>(accA, accB, tI, v, tnext, off0, str, dst0, g, t); break;
            }
            if (have_next) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&tfull[nitem & (kNTS - 1)]);
            }
        }

        // flush: sub-tile rows 7 - warp (columns 0 .. 7 - warp) and warp (columns 0 .. warp)
        double* out = args.partials + (size_t)seg.slot * BT * BT;
        const int rA = 7 - warp, rB = warp;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j <= rA)
                *reinterpret_cast<double2*>(out + (size_t)(8 * rA + g) * BT + 8 * j + 2 * t) = make_double2(accA[j][0], accA[j][1]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j <= rB)
                *reinterpret_cast<double2*>(out + (size_t)(8 * rB + g) * BT + 8 * j + 2 * t) = make_double2(accB[j][0], accB[j][1]);
    }
}

// Power-table builder on the four-warp layout (see the comment above gram_pow_smem_bytes): gram_w4_kernel whose builder
// reads a per-stage table of powers with NF = min(n, degree) factors per term.
template <int NF, int DEG = 0>  // DEG: the library's degree when instantiated for it (3, 4, 5), 0 = read from args
__global__ void __launch_bounds__(kW4Threads, 4) gram_pow_kernel(const GramArgs args) {
    using Cfg = GramCfg<64>;
    constexpr int BT = 64;
    constexpr int LDT = Cfg::kPitch;
    constexpr int KS = kKS;
    constexpr int NW = kW4Threads / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int n = args.n, n_t = args.n_t, deg = DEG > 0 ? DEG : args.pow_deg;
    const int np = n * deg;  // pseudo states
    const uint32_t xbytes = (uint32_t)(KS * n * sizeof(double));
    const uint32_t ybytes = (uint32_t)(KS * n_t * sizeof(double));
    const int vdoubles = (int)(gram_vbytes(n, n_t) / sizeof(double));
    const int pdoubles = (int)(gram_vbytes(np, n_t) / sizeof(double));

    uint64_t* vfull = reinterpret_cast<uint64_t*>(smem_raw);  // [kPowNVS] raw samples landed
    uint64_t* tfull = vfull + kPowNVS;                        // [kNTS] Theta stage built by all warps
    double* vring = reinterpret_cast<double*>(smem_raw + 128);
    double* pring = vring + kPowNVS * vdoubles;
    double* tring = pring + kPowNPS * pdoubles;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    if (tid == 0) {
        for (int i = 0; i < kPowNVS; ++i) mbar_init(&vfull[i], 1);
        for (int i = 0; i < kNTS; ++i) mbar_init(&tfull[i], NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < kPowNVS * kVX; e += kW4Threads) {
        const int w = e % kVX;
        vring[(e / kVX) * vdoubles + w] = (w < kVZero) ? 1.0 : 0.0;
    }
    for (int e = tid; e < kPowNPS * kVX; e += kW4Threads) {
        const int w = e % kVX;
        pring[(e / kVX) * pdoubles + w] = (w < kVZero) ? 1.0 : 0.0;
    }
    __syncthreads();

    const int64_t nstages_total = (args.m + KS - 1) / KS;
    const int seg_begin = args.cta_seg_begin[blockIdx.x];
    const int seg_end = args.cta_seg_begin[blockIdx.x + 1];
    int total = 0;
    for (int sg = seg_begin; sg < seg_end; ++sg) total += (int)(args.segs[sg].stage_end - args.segs[sg].stage_begin);
    if (total == 0) return;

    int tma_seg = seg_begin;
    int tma_stage = (int)args.segs[seg_begin].stage_begin, tma_seg_end = (int)args.segs[seg_begin].stage_end;
    int tma_item = 0;
    uint32_t ragged_slots = 0;
    const int ragged_stage = (args.m % KS) ? (int)(nstages_total - 1) : -1;
    auto request_samples = [&]() {
        if (tma_item >= total) return;
        const int slot = tma_item & (kPowNVS - 1);
        const bool mine = (tma_item % NW) == warp;
        if (tma_stage != ragged_stage) {
            if ((ragged_slots >> slot) & 1) {
                ragged_slots &= ~(1u << slot);
                if (mine) {
                    if (lane < KS) vring[slot * vdoubles + kVOnes + lane] = 1.0;
                    __syncwarp();
                }
            }
            if (mine && lane == 0) {
                double* dst = vring + slot * vdoubles + kVX;
                const int64_t s0 = (int64_t)tma_stage * KS;
                mbar_expect_tx(&vfull[slot], xbytes + ybytes);
                bulk_g2s(dst, args.X + s0 * n, xbytes, &vfull[slot]);
                bulk_g2s(dst + KS * n, args.Y + s0 * n_t, ybytes, &vfull[slot]);
            }
        } else {
            ragged_slots |= 1u << slot;
            if (mine) {
                double* dst = vring + slot * vdoubles;
                const int64_t s0 = (int64_t)tma_stage * KS;
                const int valid = (int)(args.m - s0);
                for (int e = lane; e < KS * n; e += 32) dst[kVX + e] = (e < valid * n) ? args.X[s0 * n + e] : 0.0;
                for (int e = lane; e < KS * n_t; e += 32)
                    dst[kVX + KS * n + e] = (e < valid * n_t) ? args.Y[s0 * n_t + e] : 0.0;
                if (lane < KS) dst[kVOnes + lane] = (lane < valid) ? 1.0 : 0.0;
                __syncwarp();
                if (lane == 0) mbar_arrive(&vfull[slot]);
            }
        }
        ++tma_item;
        if (++tma_stage == tma_seg_end && ++tma_seg < seg_end) {
            tma_stage = (int)args.segs[tma_seg].stage_begin;
            tma_seg_end = (int)args.segs[tma_seg].stage_end;
        }
    };

    // ---- the table of powers of item `pit`: thread = (sample, group of 16 threads); groups 0 .. n - 1 one state each,
    // group 7 the targets and the ones column (zero for the padding samples of a ragged last stage)
    const int es = tid & 15, eq = tid >> 4;
    auto expand = [&](int pit) {
        const double* rv = vring + (pit & (kPowNVS - 1)) * vdoubles;
        double* pv = pring + (pit & (kPowNPS - 1)) * pdoubles;
        if (eq < n) {
            const double x = rv[kVX + es * n + eq];
            double* dstp = pv + kVX + es * np + eq * deg;
            double p = x;
            dstp[0] = p;
            if constexpr (DEG > 0) {
#pragma unroll
                for (int d = 1; d < DEG; ++d) {
                    p *= x;
                    dstp[d] = p;
                }
            } else {
                for (int d = 1; d < deg; ++d) {
                    p *= x;
                    dstp[d] = p;
                }
            }
        } else if (eq == 7) {
            for (int e = 0; e < n_t; ++e) pv[kVX + KS * np + es * n_t + e] = rv[kVX + KS * n + es * n_t + e];
            pv[kVOnes + es] = rv[kVOnes + es];
        }
    };

    // ---- this thread's share of a Theta tile: column tid mod 64, samples (tid / 64) * 8 .. + 7
    int off[NF], str[NF];
    const int dst = tid & 63, s0 = (tid >> 6) * (KS / 2);
    {
        const int term = tid & 63;
        const uint4 q = *reinterpret_cast<const uint4*>(args.pow_factors + (size_t)term * kMaxFactors);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        const bool dead = term >= args.kaug;
#pragma unroll
        for (int d = 0; d < NF; ++d) {
            const uint32_t f = (w[d >> 1] >> (16 * (d & 1))) & 0xFFFFu;
            int o = kVOnes, st = 1;
            if (f != kNoFactor) {
                const bool isy = (f & kFactorIsY) != 0;
                o = isy ? kVX + KS * np + (int)(f & 0x7FFF) : kVX + (int)f;
                st = isy ? n_t : np;
            }
            if (dead) o = kVZero, st = 0;
            off[d] = o, str[d] = st;
        }
    }
    int off0[NF];
#pragma unroll
    for (int d = 0; d < NF; ++d) off0[d] = off[d] + s0 * str[d];
    const int dst0 = dst + s0 * LDT;

    // ---- prologue: raw tiles, the tables of items 0 .. 3, Theta(0 .. kAhead - 1) completely
#pragma unroll 1
    for (int i = 0; i < kPowNVS; ++i) request_samples();
#pragma unroll 1
    for (int it = 0; it < 2 * kAhead && it < total; ++it) {
        mbar_wait(&vfull[it], 0);
        expand(it);
    }
    __syncthreads();
#pragma unroll 1
    for (int it = 0; it < kAhead && it < total; ++it) {
        const double* v = pring + it * pdoubles;
        double* dstt = tring + it * Cfg::kStageDoubles;
        for (int s = s0; s < s0 + KS / 2; ++s) {
            double p = v[off[0] + s * str[0]];
#pragma unroll
            for (int d = 1; d < NF; ++d) p *= v[off[d] + s * str[d]];
            dstt[dst + s * LDT] = p;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&tfull[it]);
    }

    int item = 0;
    bool ready_t = false, ready_v = false;
    for (int sg = seg_begin; sg < seg_end; ++sg) {
        const GramSegment seg = args.segs[sg];
        double accA[8][2], accB[4][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) accA[j][0] = accA[j][1] = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) accB[j][0] = accB[j][1] = 0.0;

        const int nst = (int)(seg.stage_end - seg.stage_begin);
        for (int q = 0; q < nst; ++q, ++item) {
            const int tslot = item & (kNTS - 1);
            if (!ready_t) mbar_wait(&tfull[tslot], (uint32_t)((item / kNTS) & 1));
            // every warp has finished stage item - 2 here: the raw slot of item (read by expand four stages ago) and
            // the table slot of item (read by the builder two stages ago) are free
            request_samples();
            const int nitem = item + kAhead;      // the item whose Theta tile this stage builds (table: two stages old)
            const int pitem = item + 2 * kAhead;  // the item whose table of powers this stage writes
            const bool have_next = (nitem < total);
            if (pitem < total) {
                if (!ready_v) mbar_wait(&vfull[pitem & (kPowNVS - 1)], (uint32_t)((pitem / kPowNVS) & 1));
                expand(pitem);
            }
            ready_t = (item + 1 < total) ? mbar_test(&tfull[(item + 1) & (kNTS - 1)], (uint32_t)(((item + 1) / kNTS) & 1)) : true;
            ready_v = (pitem + 1 < total) ? mbar_test(&vfull[(pitem + 1) & (kPowNVS - 1)], (uint32_t)(((pitem + 1) / kPowNVS) & 1)) : true;

            const double* tI = tring + tslot * Cfg::kStageDoubles;
            const double* v = pring + (nitem & (kPowNPS - 1)) * pdoubles;
            double* tnext = tring + (nitem & (kNTS - 1)) * Cfg::kStageDoubles;  // (stale but harmless at the end)
            switch (warp) {
                case 0: stage_body_w4<NF, 0>(accA, accB, tI, v, tnext, off0, str, dst0, g, t); break;
                case 1: stage_body_w4<NF, 1>(accA, accB, tI, v, tnext, off0, str, dst0, g, t); break;
                case 2: stage_body_w4<NF, 2>(accA, accB, tI, v, tnext, off0, str, dst0, g, t); break;
                default: stage_body_w4<NF, 3>(accA, accB, tI, v, tnext, off0, str, dst0, g, t); break;
            }
            if (have_next) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&tfull[nitem & (kNTS - 1)]);
            }
        }

        double* out = args.partials + (size_t)seg.slot * BT * BT;
        const int rA = 7 - warp, rB = warp;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j <= rA)
                *reinterpret_cast<double2*>(out + (size_t)(8 * rA + g) * BT + 8 * j + 2 * t) = make_double2(accA[j][0], accA[j][1]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j <= rB)
                *reinterpret_cast<double2*>(out + (size_t)(8 * rB + g) * BT + 8 * j + 2 * t) = make_double2(accB[j][0], accB[j][1]);
    }
}

// Adds the partial tiles of each block pair in slot order and scatters into [G | R | yy].
__global__ void __launch_bounds__(256) gram_reduce_kernel(const double* __restrict__ partials,
                                                          const int2* __restrict__ pairs,
                                                          const int32_t* __restrict__ pair_slot_begin, int bt,
                                                          int k, int n_t, double* __restrict__ packed) {
    const int p = blockIdx.x;
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    if (e >= bt * bt) return;
    const int i = e / bt, j = e % bt;
    const int gi = pairs[p].x * bt + i;
    const int gj = pairs[p].y * bt + j;
    const int kaug = k + n_t;
    if (gi >= kaug || gj >= kaug || gj > gi) return;
    double s = 0.0;
    for (int sl = pair_slot_begin[p]; sl < pair_slot_begin[p + 1]; ++sl) s += partials[(size_t)sl * bt * bt + e];
    double* G = packed;
    double* R = packed + (size_t)k * k;
    double* yy = R + (size_t)k * n_t;
    if (gi < k) {  // both library terms
        G[(size_t)gi + (size_t)k * gj] = s;
        G[(size_t)gj + (size_t)k * gi] = s;
    } else if (gj < k) {  // target row x library term
        R[(size_t)gj + (size_t)k * (gi - k)] = s;
    } else if (gi == gj) {
        yy[gi - k] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Host side: work plan
// ---------------------------------------------------------------------------------------------
void build_gram_plan(GramPlan& plan, int kaug, int64_t m, int nctas, int bt, int ks) {
    // pair index -> (bi, bj), row-major over the lower triangle, and its cost per stage relative to a
    // full block pair (tensor-pipe time of the slowest SM sub-partition, see the shapes in the kernel)
    const int nblk = (kaug + bt - 1) / bt;
    std::vector<std::pair<int, int>> pairs;
    std::vector<double> cost;
    for (int bi = 0; bi < nblk; ++bi)
        for (int bj = 0; bj <= bi; ++bj) {
            pairs.push_back({bi, bj});
            const int rows_valid = std::min(bt, kaug - bi * bt);
            double c = 1.0;
            if (bt == 128 && rows_valid <= 64) c = 0.53;       // flat: 64 x 128 patch
            else if (bi == bj) c = (bt == 128) ? 0.60 : 0.85;  // diagonal: 36/64 (+ build) resp. 26/32
            cost.push_back(c);
        }
    build_plan_from_pairs(plan, pairs, cost, m, nctas, bt, ks);
    plan.nblk = nblk;
}

void build_plan_from_pairs(GramPlan& plan, const std::vector<std::pair<int, int>>& pairs,
                           const std::vector<double>& cost, int64_t m, int nctas, int bt, int ks) {
    plan = GramPlan();
    plan.bt = bt;
    plan.ks = ks;
    plan.nblk = 0;
    plan.npairs = (int)pairs.size();
    plan.nctas = nctas;
    plan.nstages = (m + ks - 1) / ks;
    plan.pairs = pairs;
    const int64_t ns = plan.nstages;

    struct Tmp {
        int cta, pair;
        int64_t b, e;
    };
    std::vector<Tmp> tmp;
    double total = 0.0;
    for (double c : cost) total += c * (double)ns;
    const double budget = total / nctas;
    std::vector<double> used(nctas, 0.0);
    // (1) whole pairs, most expensive first, dealt round-robin to CTAs that can still fit one: these
    // CTAs walk the sample stream in lockstep, so each X/Y tile is fetched from HBM once and served
    // to the other CTAs from L2.
    std::vector<int> order(plan.npairs);
    for (int p = 0; p < plan.npairs; ++p) order[p] = p;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<int> rest;
    int next_cta = 0;
    for (int p : order) {
        const double need = cost[p] * (double)ns;
        int found = -1;
        for (int probe = 0; probe < nctas && ns > 0; ++probe) {
            const int c = (next_cta + probe) % nctas;
            if (used[c] + need <= budget * (1.0 + 1e-9)) {
                found = c;
                break;
            }
        }
        if (found < 0) {
            rest.push_back(p);
            continue;
        }
        tmp.push_back({found, p, 0, ns});
        used[found] += need;
        next_cta = (found + 1) % nctas;
    }
    // (2) the remaining (pair, stage) stream fills every CTA up to the common budget, contiguously:
    // CTA c owns the part of the stream whose cumulative cost lies in (cap[c-1], cap[c]].
    if (!rest.empty() && ns > 0) {
        double rest_total = 0.0;
        for (int p : rest) rest_total += cost[p] * (double)ns;
        std::vector<double> cap(nctas);
        double free_total = 0.0;
        for (int c = 0; c < nctas; ++c) free_total += std::max(0.0, budget - used[c]);
        double acc = 0.0;
        for (int c = 0; c < nctas; ++c) {
            acc += std::max(0.0, budget - used[c]) * (free_total > 0.0 ? rest_total / free_total : 0.0);
            cap[c] = acc;
        }
        cap[nctas - 1] = rest_total * (1.0 + 1e-9) + 1.0;
        int c = 0;
        double cum = 0.0;  // cost of the stream before the current pair
        for (int p : rest) {
            int64_t st = 0;
            while (st < ns) {
                // last stage (exclusive) of pair p whose midpoint still falls into CTA c's share
                int64_t end = (int64_t)std::floor((cap[c] - cum) / cost[p] + 0.5);
                end = std::min<int64_t>(std::max<int64_t>(end, st), ns);
                if (end > st) {
                    tmp.push_back({c, p, st, end});
                    st = end;
                }
                if (st < ns) ++c;  // CTA c is full
                if (c >= nctas) c = nctas - 1;
            }
            cum += cost[p] * (double)ns;
        }
    }
    // slots: grouped by pair, inside a pair ordered by CTA (fixed summation order)
    std::vector<int> sorder(tmp.size());
    for (size_t i = 0; i < tmp.size(); ++i) sorder[i] = (int)i;
    std::stable_sort(sorder.begin(), sorder.end(), [&](int a, int b) {
        if (tmp[a].pair != tmp[b].pair) return tmp[a].pair < tmp[b].pair;
        if (tmp[a].cta != tmp[b].cta) return tmp[a].cta < tmp[b].cta;
        return tmp[a].b < tmp[b].b;
    });
    std::vector<int> slot_of(tmp.size());
    plan.pair_slot_begin.assign(plan.npairs + 1, 0);
    for (size_t i = 0; i < sorder.size(); ++i) {
        slot_of[sorder[i]] = (int)i;
        plan.pair_slot_begin[tmp[sorder[i]].pair + 1]++;
    }
    for (int p = 0; p < plan.npairs; ++p) plan.pair_slot_begin[p + 1] += plan.pair_slot_begin[p];
    plan.nslots = (int)tmp.size();

    // segments sorted by CTA (tmp is already CTA-major for the full rounds, then CTA-major again for
    // the remainder -> stable sort by CTA keeps full rounds first)
    std::vector<int> by_cta(tmp.size());
    for (size_t i = 0; i < tmp.size(); ++i) by_cta[i] = (int)i;
    std::stable_sort(by_cta.begin(), by_cta.end(), [&](int a, int b) { return tmp[a].cta < tmp[b].cta; });
    plan.cta_seg_begin.assign(nctas + 1, 0);
    plan.segments.reserve(tmp.size());
    for (int idx : by_cta) {
        const Tmp& t = tmp[idx];
        GramSegment s;
        s.bi = pairs[t.pair].first;
        s.bj = pairs[t.pair].second;
        s.slot = slot_of[idx];
        s.pad = 0;
        s.stage_begin = t.b;
        s.stage_end = t.e;
        plan.segments.push_back(s);
        plan.cta_seg_begin[t.cta + 1]++;
    }
    for (int c = 0; c < nctas; ++c) plan.cta_seg_begin[c + 1] += plan.cta_seg_begin[c];
}

// ---------------------------------------------------------------------------------------------
// Host side: launch
// ---------------------------------------------------------------------------------------------

// Kernel for a block tile and the library's maximum total degree: a term is a product of NF staged values,
// NF = smallest instantiated count >= the degree (every FP64 multiply costs a DMMA issue slot).
using GramKernel = void (*)(const GramArgs);
template <int BT>
static GramKernel pick_gram_kernel(int max_degree) {
    if (max_degree <= 2) return gram_kernel<BT, 2>;
    if (max_degree <= 3) return gram_kernel<BT, 3>;
    if (max_degree <= 4) return gram_kernel<BT, 4>;
    if (max_degree <= 5) return gram_kernel<BT, 5>;
    return gram_kernel<BT, kMaxFactors>;
}
static GramKernel pick_gram_kernel(int bt, int max_degree) {
    return bt == 64 ? pick_gram_kernel<64>(max_degree) : pick_gram_kernel<128>(max_degree);
}

static int launch_gram(ddb_ctx* ctx, int bt, const GramArgs& args, size_t smem) {
    GramKernel kern = pick_gram_kernel(bt, ctx->max_degree);
    DDB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ctx->plan.nctas, bt == 64 ? GramCfg<64>::kThreads : GramCfg<128>::kThreads, smem, ctx->stream>>>(args);
    DDB_CUDA_TRY(cudaGetLastError());
    return DDB_OK;
}

// DMMA flops per sample of the plain schedules: live 8 x 8 sub-tiles of every lower-triangular block pair
// (diagonal pairs skip the sub-tiles above the diagonal, a last block row of at most 64 rows is a 64-row patch)
double gram_plain_dmma_flops(int kaug, int bt) {
    const int nblk = (kaug + bt - 1) / bt, sub = bt / 8;
    double tiles = 0.0;
    for (int bi = 0; bi < nblk; ++bi) {
        const int rows_valid = std::min(bt, kaug - bi * bt);
        for (int bj = 0; bj <= bi; ++bj) {
            if (bt == 128 && rows_valid <= 64) tiles += 8.0 * sub;
            else if (bi == bj) tiles += 0.5 * sub * (sub + 1);
            else tiles += (double)sub * sub;
        }
    }
    return tiles * 2.0 * 64.0;
}

int gram_ensure_factors(ddb_ctx* ctx, int kaug_pad);
static int ensure_factor_table(ddb_ctx* ctx, int kaug_pad) { return gram_ensure_factors(ctx, kaug_pad); }
int gram_ensure_factors(ddb_ctx* ctx, int kaug_pad) {
    // (k + n_t) rows of kMaxFactors uint16: library terms then one single-factor row per target
    const int k = ctx->k, n = ctx->n, n_t = ctx->n_t;
    std::vector<uint16_t> tab((size_t)kaug_pad * kMaxFactors, kNoFactor);
    for (int j = 0; j < k; ++j) {
        int d = 0;
        for (int i = 0; i < n; ++i)
            for (int e = 0; e < ctx->exps[(size_t)i + (size_t)n * j]; ++e) tab[(size_t)j * kMaxFactors + d++] = (uint16_t)i;
    }
    for (int i = 0; i < n_t; ++i) tab[(size_t)(k + i) * kMaxFactors] = (uint16_t)(kFactorIsY | i);
    DDB_TRY(ctx->d_factors.reserve(tab.size() * sizeof(uint16_t)));
    DDB_CUDA_TRY(cudaMemcpyAsync(ctx->d_factors.p, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice,
                                 ctx->stream));
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // tab is a stack-lifetime host buffer
    ctx->factors_nt = n_t;
    return DDB_OK;
}

// Runs the self-building kernel (BT = 128) over an arbitrary segment list; used by gram_ring.cu for the
// block pairs that do not get a whole CTA in the ring pass.
template <int BT>
static GramKernel pick_gram_kernel_offdiag(int max_degree) {
    if (max_degree <= 2) return gram_kernel<BT, 2, true>;
    if (max_degree <= 3) return gram_kernel<BT, 3, true>;
    if (max_degree <= 4) return gram_kernel<BT, 4, true>;
    if (max_degree <= 5) return gram_kernel<BT, 5, true>;
    return gram_kernel<BT, kMaxFactors, true>;
}
int gram_launch_bt128(ddb_ctx* ctx, const GramArgs& args, int nctas, bool offdiag_only) {
    const size_t smem = gram_smem_bytes<128>(args.n, args.n_t);
    const char* oenv = getenv("DDB_GRAM_OFFDIAG");
    if (oenv && atoi(oenv) == 0) offdiag_only = false;
    GramKernel kern = offdiag_only ? pick_gram_kernel_offdiag<128>(ctx->max_degree) : pick_gram_kernel<128>(ctx->max_degree);
    DDB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nctas, GramCfg<128>::kThreads, smem, ctx->stream>>>(args);
    DDB_CUDA_TRY(cudaGetLastError());
    return DDB_OK;
}
// The tensor-pipe builder variant (args.optab / args.opcnt set by gram_moment.cu); false if it does not fit.
bool gram_mb_fits(int n, int n_t) {
    // rows of the X tile are separate bulk copies: 16-byte multiples
    return n % 2 == 0 && gram_mb_smem_bytes(n, n_t) <= 227 * 1024 && gram_mb_vbytes(n, n_t) <= 65528;
}
int gram_launch_mb128(ddb_ctx* ctx, const GramArgs& args, int nctas) {
    const size_t smem = gram_mb_smem_bytes(args.n, args.n_t);
    DDB_CUDA_TRY(cudaFuncSetAttribute(gram_mb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gram_mb_kernel<<<nctas, GramCfg<128>::kThreads, smem, ctx->stream>>>(args);
    DDB_CUDA_TRY(cudaGetLastError());
    return DDB_OK;
}
void gram_reduce_launch(ddb_ctx* ctx, const double* partials, const int2* pairs, const int32_t* pair_slot_begin,
                        int npairs, int bt, int k, int n_t, double* packed) {
    dim3 rgrid(npairs, (bt * bt + 255) / 256);
    gram_reduce_kernel<<<rgrid, 256, 0, ctx->stream>>>(partials, pairs, pair_slot_begin, bt, k, n_t, packed);
}
int gram_ensure_factors(ddb_ctx* ctx, int rows);
int gram_ring_run(ddb_ctx* ctx, double* dPacked, bool* handled);
int gram_moment_run(ddb_ctx* ctx, double* dPacked, bool* handled);

// Programs of gram_tree_kernel for the installed library and n_t targets: the term tree (parent = the term without
// its last factor; the targets hang under the constant) in depth-first order, cut into four contiguous quarters,
// each preceded by the ancestors of its first node.  Returns the steps per program (a multiple of 4), 0 if the
// library has no constant, is not closed under "drop the last factor", or repeats a term.  words: 4 x steps of
// dst column | parent column << 8 (255: none) | factor code << 16.
int gram_tree_program(const uint8_t* exps, int n, int k, int n_t, std::vector<uint32_t>* words) {
    const int kaug = k + n_t;
    if (kaug > 64 || k < 1) return 0;
    std::vector<std::vector<uint16_t>> fac(k);
    std::map<std::vector<uint16_t>, int> col_of;
    for (int j = 0; j < k; ++j) {
        for (int i = 0; i < n; ++i)
            for (int e = 0; e < exps[(size_t)i + (size_t)n * j]; ++e) fac[j].push_back((uint16_t)i);
        if (!col_of.emplace(fac[j], j).second) return 0;
    }
    const auto rootit = col_of.find(std::vector<uint16_t>());
    if (rootit == col_of.end()) return 0;
    const int root = rootit->second;
    std::vector<int> parent(kaug, -1);
    std::vector<uint16_t> code(kaug, kNoFactor);
    std::vector<std::vector<int>> children(kaug);
    for (int j = 0; j < k; ++j) {
        if (j == root) continue;
        std::vector<uint16_t> pf(fac[j].begin(), fac[j].end() - 1);
        const auto it = col_of.find(pf);
        if (it == col_of.end()) return 0;
        parent[j] = it->second;
        code[j] = fac[j].back();
        children[it->second].push_back(j);
    }
    for (int i = 0; i < n_t; ++i) {
        parent[k + i] = root;
        code[k + i] = (uint16_t)(kFactorIsY | i);
        children[root].push_back(k + i);
    }
    std::vector<int> order, depth(kaug, 0), stack = {root};
    while (!stack.empty()) {
        const int u = stack.back();
        stack.pop_back();
        order.push_back(u);
        for (auto it = children[u].rbegin(); it != children[u].rend(); ++it) {
            depth[*it] = depth[u] + 1;
            stack.push_back(*it);
        }
    }
    if ((int)order.size() != kaug) return 0;
    const int seglen = (kaug + kTreeQuarters - 1) / kTreeQuarters;
    std::vector<std::vector<int>> nodes(kTreeQuarters);  // ancestor-closed node sets, by depth then depth-first index
    std::vector<int> dfs_pos(kaug, 0);
    for (int i = 0; i < kaug; ++i) dfs_pos[order[i]] = i;
    for (int q = 0; q < kTreeQuarters; ++q) {
        const int b = q * seglen, e = std::min(kaug, b + seglen);
        if (b >= e) continue;
        for (int u = parent[order[b]]; u >= 0; u = parent[u]) nodes[q].push_back(u);
        for (int i = b; i < e; ++i) nodes[q].push_back(order[i]);
        std::sort(nodes[q].begin(), nodes[q].end(), [&](int x, int y) {
            return depth[x] != depth[y] ? depth[x] < depth[y] : dfs_pos[x] < dfs_pos[y];
        });
    }
    // the kernel runs the steps of a k-step in four groups (tree_group_size); a node goes into the first group after
    // its parent's.  Smallest step count (a multiple of 4, at most 24) that takes all four programs.
    const uint32_t idle = 64u | (255u << 8) | ((uint32_t)kNoFactor << 16);  // multiplies ones into a pad column
    for (int steps = 4; steps <= 24; steps += 4) {
        const int spk = steps / 4;
        words->assign((size_t)kTreeQuarters * steps, idle);
        bool fits = true;
        for (int q = 0; q < kTreeQuarters && fits; ++q) {
            std::vector<int> group_of(kaug, -1);
            int grp = 0, used = 0;  // group index (four per k-step), slots used in it
            for (int u : nodes[q]) {
                const int need = parent[u] < 0 ? 0 : group_of[parent[u]] + 1;
                while (grp < 16 && (grp < need || used >= tree_group_size(spk, grp % 4))) ++grp, used = 0;
                if (grp >= 16) {
                    fits = false;
                    break;
                }
                group_of[u] = grp;
                (*words)[(size_t)q * steps + (grp / 4) * spk + tree_group_first(spk, grp % 4) + used++] =
                    (uint32_t)u | ((uint32_t)(parent[u] < 0 ? 255 : parent[u]) << 8) | ((uint32_t)code[u] << 16);
            }
        }
        if (fits) return steps;
    }
    return 0;
}

static GramKernel pick_tree_kernel(int steps) {
    switch (steps) {
        case 4: return gram_tree_kernel<4>;
        case 8: return gram_tree_kernel<8>;
        case 12: return gram_tree_kernel<12>;
        case 16: return gram_tree_kernel<16>;
        case 20: return gram_tree_kernel<20>;
        default: return gram_tree_kernel<24>;
    }
}

// (Re)builds the programs when the library or the number of targets changed; ctx->tree_nsteps = 0: not applicable.
static int ensure_tree_program(ddb_ctx* ctx) {
    if (ctx->tree_lib_version == ctx->lib_version && ctx->tree_nt == ctx->n_t) return DDB_OK;
    std::vector<uint32_t> words;
    const int steps = gram_tree_program(ctx->exps.data(), ctx->n, ctx->k, ctx->n_t, &words);
    if (steps > 0) {
        DDB_TRY(ctx->d_tree.reserve(words.size() * sizeof(uint32_t)));
        DDB_CUDA_TRY(cudaMemcpyAsync(ctx->d_tree.p, words.data(), words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    if (getenv("DDB_GRAM_DEBUG")) fprintf(stderr, "[ddb] prefix-product builder: %d steps per program (0 = not applicable)\n", steps);
    ctx->tree_nsteps = steps;
    ctx->tree_lib_version = ctx->lib_version;
    ctx->tree_nt = ctx->n_t;
    return DDB_OK;
}

template <int NF>
static GramKernel pick_pow_kernel_deg(int deg) {
    switch (deg) {
        case 3: return gram_pow_kernel<NF, 3>;
        case 4: return gram_pow_kernel<NF, 4>;
        case 5: return gram_pow_kernel<NF, 5>;
        default: return gram_pow_kernel<NF, 0>;
    }
}
static GramKernel pick_pow_kernel(int nf, int deg) {
    return nf == 1 ? pick_pow_kernel_deg<1>(deg) : nf == 2 ? pick_pow_kernel_deg<2>(deg) : pick_pow_kernel_deg<3>(deg);
}

// Power-table builder: (re)builds the factor codes when the library or the number of targets changed;
// ctx->pow_nf = 0: not applicable (more than 3 states -- the table is written by (sample, state) threads, the fourth
// group copies the targets --, or degree < 3 where a power per state saves no multiply).
// Factor codes of the power-table builder for a monomial library: per augmented row (64 rows x kMaxFactors codes) one
// code per state with a non-zero exponent, state * deg + exponent - 1 (an index into the stage's table of powers), the
// targets as kFactorIsY | t, kNoFactor elsewhere.  Returns the number of factors per term the kernel needs (1 .. 3), 0
// when the builder does not apply.  Host only (also behind ddb_pow_factors_describe for the CPU tests).
int gram_pow_factors(const uint8_t* exps, int n, int k, int n_t, int deg, std::vector<uint16_t>* tab) {
    if (n > 3 || deg < 3 || deg > kMaxFactors || k + n_t > 64) return 0;
    tab->assign((size_t)64 * kMaxFactors, kNoFactor);
    int nf = 1;
    for (int j = 0; j < k; ++j) {
        int d = 0;
        for (int i = 0; i < n; ++i) {
            const int e = exps[(size_t)i + (size_t)n * j];
            if (e > 0) (*tab)[(size_t)j * kMaxFactors + d++] = (uint16_t)(i * deg + e - 1);
        }
        nf = std::max(nf, d);
    }
    for (int i = 0; i < n_t; ++i) (*tab)[(size_t)(k + i) * kMaxFactors] = (uint16_t)(kFactorIsY | i);
    return nf;
}

static int ensure_pow_factors(ddb_ctx* ctx) {
    if (ctx->pow_lib_version == ctx->lib_version && ctx->pow_nt == ctx->n_t) return DDB_OK;
    std::vector<uint16_t> tab;
    const int nf = gram_pow_factors(ctx->exps.data(), ctx->n, ctx->k, ctx->n_t, ctx->max_degree, &tab);
    if (nf > 0) {
        DDB_TRY(ctx->d_pow_factors.reserve(tab.size() * sizeof(uint16_t)));
        DDB_CUDA_TRY(cudaMemcpyAsync(ctx->d_pow_factors.p, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice,
                                     ctx->stream));
        DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    ctx->pow_nf = nf;
    ctx->pow_lib_version = ctx->lib_version;
    ctx->pow_nt = ctx->n_t;
    return DDB_OK;
}

int gram_run(ddb_ctx* ctx, double* dPacked) {
    if (ctx->k <= 0) {
        set_error("ddb_gram: no library installed (call ddb_set_library_monomial first)");
        return DDB_BAD_STATE;
    }
    if (ctx->view_empty && ctx->n_t > 0) {  // no sample in this rank's view: zero blocks (the other ranks hold the data)
        const int64_t count = ddb_gram_count(ctx);
        if (!dPacked) {
            DDB_TRY(ctx->d_packed.reserve((size_t)count * sizeof(double)));
            dPacked = ctx->d_packed.as<double>();
        }
        DDB_CUDA_TRY(cudaMemsetAsync(dPacked, 0, (size_t)count * sizeof(double), ctx->stream));
        if (ctx->factors_nt != ctx->n_t) DDB_TRY(gram_ensure_factors(ctx, ((ctx->k + ctx->n_t + 127) / 128) * 128));
        DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->timings.gram_ms = 0.0;
        ctx->timings.gram_launches = 0;
        ctx->have_gram = (dPacked == ctx->d_packed.as<double>());
        ctx->imp_k = 0;
        return DDB_OK;
    }
    if (!ctx->dX || !ctx->dY || ctx->m <= 0) {
        set_error("ddb_gram: no data bound (call ddb_upload / ddb_bind_device / ddb_generate first)");
        return DDB_BAD_STATE;
    }
    const int k = ctx->k, n = ctx->n, n_t = ctx->n_t;
    const int kaug = k + n_t;
    const int bt = (kaug <= 64) ? 64 : 128;
    const int ks = kKS;
    if (n >= 0x7FFF || n_t >= 0x7FFF) {
        set_error("ddb_gram: more than 32766 states/targets are not supported by the monomial evaluator");
        return DDB_UNSUPPORTED;
    }
    const size_t smem = (bt == 64) ? gram_smem_bytes<64>(n, n_t) : gram_smem_bytes<128>(n, n_t);
    if (smem > 227 * 1024) {
        set_error("ddb_gram: n + n_t = %d is too wide to stage a sample tile in shared memory (%zu bytes)", n + n_t,
                  smem);
        return DDB_UNSUPPORTED;
    }
    if (((uintptr_t)ctx->dX & 15) || ((uintptr_t)ctx->dY & 15)) {
        set_error("ddb_gram: device data pointers must be 16-byte aligned");
        return DDB_INVALID_ARG;
    }
    if (bt == 128) {
        bool handled = false;
        DDB_TRY(gram_moment_run(ctx, dPacked, &handled));
        if (handled) return DDB_OK;
        DDB_TRY(gram_ring_run(ctx, dPacked, &handled));
        if (handled) return DDB_OK;
    }
    // single diagonal pair of a library that is closed under "drop the last factor": one multiply per library value
    // (gram_tree_kernel).  Opt-in (DDB_GRAM_TREE=1): bit-identical and 24 instead of 64 multiplies per thread and stage
    // at C2, but the parent -> child chains through shared memory cost more than the multiplies saved (2.57 ms in
    // depth-first order, 2.67 ms in groups of independent steps, against 2.45 ms; profiles/gram_tree_experiment_r2.txt)
    GramKernel tree_kern = nullptr;
    if (bt == 64 && ctx->max_degree >= 2) {
        const char* tenv = getenv("DDB_GRAM_TREE");
        if (tenv && atoi(tenv) == 1) {
            DDB_TRY(ensure_tree_program(ctx));
            if (ctx->tree_nsteps > 0 && gram_tree_smem_bytes(n, n_t, ctx->tree_nsteps) <= 56 * 1024)
                tree_kern = pick_tree_kernel(ctx->tree_nsteps);
        }
    }
    // at most 3 states, degree >= 3 (C2): one power per state instead of one factor per degree (gram_pow_kernel).
    // On the two-warp layout the multiply count did not matter (2.28 against 2.25 ms); with four warps per sub-partition
    // it does: C2 2.015 -> 1.947 ms (profiles/gram_c2_experiments_r2.txt).  DDB_GRAM_POW=0 keeps the factor builder.
    GramKernel pow_kern = nullptr;
    if (bt == 64 && !tree_kern) {
        const char* penv = getenv("DDB_GRAM_POW");
        const char* wenv0 = getenv("DDB_GRAM_W4");
        if (!(penv && atoi(penv) == 0) && !(wenv0 && atoi(wenv0) == 0)) {
            DDB_TRY(ensure_pow_factors(ctx));
            if (ctx->pow_nf > 0 && gram_pow_smem_bytes(n, n_t, ctx->max_degree) <= 56 * 1024)
                pow_kern = pick_pow_kernel(ctx->pow_nf, ctx->max_degree);
        }
    }
    // default for single-pair libraries: four balanced warps (gram_w4_kernel); DDB_GRAM_W4=0 keeps gram_kernel<64>
    int kern_threads = GramCfg<64>::kThreads;
    bool w4 = false;
    if (pow_kern) {
        tree_kern = pow_kern;  // launched the same way (its own thread count and shared-memory size)
        kern_threads = kW4Threads;
    }
    if (bt == 64 && !tree_kern) {
        const char* wenv = getenv("DDB_GRAM_W4");
        if (!(wenv && atoi(wenv) == 0)) {
            const int dg = ctx->max_degree;
            tree_kern = dg <= 2 ? gram_w4_kernel<2> : dg <= 3 ? gram_w4_kernel<3> : dg <= 4 ? gram_w4_kernel<4>
                        : dg <= 5 ? gram_w4_kernel<5> : gram_w4_kernel<kMaxFactors>;
            // (five CTAs per SM -- 96 registers, no spills -- is slower: 2.14 against 2.04 ms at C2)
            kern_threads = kW4Threads;
            w4 = true;
        }
    }
    const size_t smem_launch = pow_kern ? gram_pow_smem_bytes(n, n_t, ctx->max_degree)
                               : w4 ? smem
                               : tree_kern ? gram_tree_smem_bytes(n, n_t, ctx->tree_nsteps) : smem;
    // persistent grid: CTAs per SM from the occupancy calculator (1 for BT=128, up to 3 for BT=64)
    int occ = 1;
    {
        GramKernel kern = tree_kern ? tree_kern : pick_gram_kernel(bt, ctx->max_degree);
        const size_t smem = smem_launch;
        DDB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, kern, bt == 64 ? kern_threads : GramCfg<128>::kThreads, smem);
        DDB_CUDA_TRY(oe);
        if (occ < 1) {
            set_error("ddb_gram: kernel does not fit on an SM (smem %zu bytes)", smem);
            return DDB_UNSUPPORTED;
        }
    }
    const int nctas = ctx->sm_count * occ;
    // plan (cached per shape)
    if (ctx->plan_m != ctx->m || ctx->plan_kaug != kaug || ctx->plan.bt != bt || ctx->plan.nctas != nctas) {
        build_gram_plan(ctx->plan, kaug, ctx->m, nctas, bt, ks);
        const GramPlan& pl = ctx->plan;
        std::vector<int2> pairs;
        for (int bi = 0; bi < pl.nblk; ++bi)
            for (int bj = 0; bj <= bi; ++bj) pairs.push_back(make_int2(bi, bj));
        size_t off_seg = 0;
        size_t off_cta = off_seg + pl.segments.size() * sizeof(GramSegment);
        size_t off_pairs = (off_cta + pl.cta_seg_begin.size() * sizeof(int32_t) + 15) & ~size_t(15);
        size_t off_psb = off_pairs + pairs.size() * sizeof(int2);
        size_t total = off_psb + pl.pair_slot_begin.size() * sizeof(int32_t);
        std::vector<unsigned char> host(total);
        memcpy(host.data() + off_seg, pl.segments.data(), pl.segments.size() * sizeof(GramSegment));
        memcpy(host.data() + off_cta, pl.cta_seg_begin.data(), pl.cta_seg_begin.size() * sizeof(int32_t));
        memcpy(host.data() + off_pairs, pairs.data(), pairs.size() * sizeof(int2));
        memcpy(host.data() + off_psb, pl.pair_slot_begin.data(), pl.pair_slot_begin.size() * sizeof(int32_t));
        DDB_TRY(ctx->d_plan.reserve(total));
        DDB_CUDA_TRY(cudaMemcpyAsync(ctx->d_plan.p, host.data(), total, cudaMemcpyHostToDevice, ctx->stream));
        DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        DDB_TRY(ctx->d_partials.reserve((size_t)pl.nslots * bt * bt * sizeof(double)));
        ctx->plan_m = ctx->m;
        ctx->plan_kaug = kaug;
        ctx->factors_nt = -1;
    }
    const GramPlan& pl = ctx->plan;
    if (ctx->factors_nt != n_t) DDB_TRY(ensure_factor_table(ctx, pl.nblk * bt));
    if (!dPacked) {
        DDB_TRY(ctx->d_packed.reserve((size_t)ddb_gram_count(ctx) * sizeof(double)));
        dPacked = ctx->d_packed.as<double>();
    }

    unsigned char* base = static_cast<unsigned char*>(ctx->d_plan.p);
    size_t off_cta = pl.segments.size() * sizeof(GramSegment);
    size_t off_pairs = (off_cta + pl.cta_seg_begin.size() * sizeof(int32_t) + 15) & ~size_t(15);
    size_t off_psb = off_pairs + (size_t)pl.npairs * sizeof(int2);

    GramArgs args;
    args.X = ctx->dX;
    args.Y = ctx->dY;
    args.n = n;
    args.n_t = n_t;
    args.m = ctx->m;
    args.kaug = kaug;
    args.factors = ctx->d_factors.as<uint16_t>();
    args.segs = reinterpret_cast<const GramSegment*>(base);
    args.cta_seg_begin = reinterpret_cast<const int32_t*>(base + off_cta);
    args.partials = ctx->d_partials.as<double>();

    args.swap_mode = getenv("DDB_GRAM_SWAP") ? atoi(getenv("DDB_GRAM_SWAP")) : 1;
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (tree_kern) {
        args.tree = ctx->d_tree.as<uint32_t>();
        args.pow_factors = ctx->d_pow_factors.as<uint16_t>();
        args.pow_deg = ctx->max_degree;
        tree_kern<<<ctx->plan.nctas, kern_threads, smem_launch, ctx->stream>>>(args);
        DDB_CUDA_TRY(cudaGetLastError());
    } else {
        const int st = launch_gram(ctx, bt, args, smem);
        DDB_TRY(st);
    }
    dim3 rgrid(pl.npairs, (bt * bt + 255) / 256);
    gram_reduce_kernel<<<rgrid, 256, 0, ctx->stream>>>(ctx->d_partials.as<double>(),
                                                       reinterpret_cast<const int2*>(base + off_pairs),
                                                       reinterpret_cast<const int32_t*>(base + off_psb), bt, k, n_t,
                                                       dPacked);
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->timings.gram_ms = ms;
    ctx->timings.gram_launches = 2;
    ctx->timings.gram_schedule = pow_kern ? 3 : (w4 ? 4 : 0);
    ctx->timings.gram_dmma_flops = gram_plain_dmma_flops(kaug, bt);
    ctx->have_gram = (dPacked == ctx->d_packed.as<double>());
    ctx->imp_k = 0;
    return DDB_OK;
}

}  // namespace ddb
