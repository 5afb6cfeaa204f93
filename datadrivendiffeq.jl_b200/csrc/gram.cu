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
// Per CTA (persistent, one per SM; the host plan hands it a list of (block pair, stage range)):
//   stage = KS consecutive samples
//   1. TMA (cp.async.bulk) brings the X tile (KS x n doubles, contiguous in memory because the
//      reference layout is sample-contiguous) and the Y tile into shared memory, double buffered,
//      signalled through an mbarrier.
//   2. Every thread owns one library term (its monomial factor list sits in registers) and
//      multiplies it out for a few samples: Theta_I' and Theta_J' tiles (KS x BT, sample-major,
//      padded to a conflict-free pitch) are built in shared memory only -- Theta never reaches HBM.
//   3. Each warp owns a 32 x 32 patch of the BT x BT block and accumulates it with
//      mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 -- the only FP64 tensor shape sm_100a implements),
//      16 independent accumulator tiles per warp, A and B fragments read from the two Theta tiles.
//   Accumulators stay in registers across all stages of a segment and are flushed once to a
//   partial-tile slot; a second kernel adds the slots of each block pair in a fixed order
//   (bit-reproducible: no floating-point atomics) and scatters into the packed [G | R | yy].
#include "common.cuh"
#include <algorithm>
#include <cstring>

namespace ddb {

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            "  .reg .pred p;\n"
            "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "  selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk asynchronous copy global -> shared through the TMA unit (SASS UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct GramArgs {
    const double* X;
    const double* Y;
    int n, n_t;
    int64_t m;
    int kaug;
    const uint16_t* factors;  // kaug_pad x kMaxFactors
    const GramSegment* segs;
    const int32_t* cta_seg_begin;
    double* partials;
};

template <int BT>
struct GramCfg {
    static constexpr int kWarpsPerDim = BT / 32;
    static constexpr int kWarps = kWarpsPerDim * kWarpsPerDim;
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kPitch = BT + 4;  // doubles; pitch % 16 == 4 -> fragment loads hit 16 distinct banks
};

__host__ __device__ inline size_t gram_smem_bytes(int bt, int ks, int n, int n_t) {
    size_t v = (size_t)ks * (n + n_t) * sizeof(double);
    v = (v + 127) & ~size_t(127);
    size_t th = (size_t)ks * (bt + 4) * sizeof(double);
    return 128 + 2 * v + 4 * th;
}

template <int BT, int KS, int NF>
__global__ void __launch_bounds__(GramCfg<BT>::kThreads, 1) gram_kernel(const GramArgs args) {
    using Cfg = GramCfg<BT>;
    constexpr int NT = Cfg::kThreads;
    constexpr int LDT = Cfg::kPitch;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int n = args.n, n_t = args.n_t;
    const size_t xbytes = (size_t)KS * n * sizeof(double);
    const size_t ybytes = (size_t)KS * n_t * sizeof(double);
    size_t vbytes = (xbytes + ybytes + 127) & ~size_t(127);

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // 2 mbarriers
    unsigned char* vbase = smem_raw + 128;
    double* theta = reinterpret_cast<double*>(vbase + 2 * vbytes);  // [buf 2][tile 2][KS][LDT]
    constexpr int kTileDoubles = KS * LDT;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wm = warp / Cfg::kWarpsPerDim;
    const int wn = warp % Cfg::kWarpsPerDim;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase0 = 0, phase1 = 0;  // parity to wait for on each buffer's barrier

    const int64_t nstages_total = (args.m + KS - 1) / KS;
    const int seg_begin = args.cta_seg_begin[blockIdx.x];
    const int seg_end = args.cta_seg_begin[blockIdx.x + 1];

    for (int sg = seg_begin; sg < seg_end; ++sg) {
        const GramSegment seg = args.segs[sg];
        const bool same = (seg.bi == seg.bj);
        const int nslots = same ? BT : 2 * BT;
        const int groups = NT / nslots;
        const int spg = KS / groups;  // samples per thread per stage
        const int slot = tid % nslots;
        const int grp = tid / nslots;
        const int tile = slot / BT;   // 0 -> block row bi, 1 -> block column bj
        const int col = slot % BT;
        const int term = (tile == 0 ? seg.bi : seg.bj) * BT + col;

        // factor list of this thread's term: up to NF offsets into the staged sample row
        uint16_t fac[NF];
        {
            const uint16_t* f = args.factors + (size_t)term * kMaxFactors;
#pragma unroll
            for (int d = 0; d < NF; ++d) fac[d] = f[d];
        }
        const double init = (term < args.kaug) ? 1.0 : 0.0;

        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // ---- staging helpers --------------------------------------------------------------
        auto load_stage = [&](int64_t st, int buf) {
            const int64_t s0 = st * KS;
            unsigned char* dst = vbase + (size_t)buf * vbytes;
            const bool full = (st + 1 < nstages_total) || (args.m - s0 == KS);
            if (full) {
                if (tid == 0) {
                    mbar_expect_tx(&bars[buf], (uint32_t)(xbytes + ybytes));
                    bulk_g2s(dst, args.X + s0 * n, (uint32_t)xbytes, &bars[buf]);
                    bulk_g2s(dst + xbytes, args.Y + s0 * n_t, (uint32_t)ybytes, &bars[buf]);
                }
            } else {
                // ragged last stage of the shard: plain loads, zero fill (TMA needs 16-byte multiples)
                const int valid = (int)(args.m - s0);
                double* vx = reinterpret_cast<double*>(dst);
                double* vy = reinterpret_cast<double*>(dst + xbytes);
                for (int e = tid; e < KS * n; e += NT) vx[e] = (e < valid * n) ? args.X[s0 * n + e] : 0.0;
                for (int e = tid; e < KS * n_t; e += NT) vy[e] = (e < valid * n_t) ? args.Y[s0 * n_t + e] : 0.0;
            }
        };
        auto wait_stage = [&](int64_t st, int buf) {
            const int64_t s0 = st * KS;
            const bool full = (st + 1 < nstages_total) || (args.m - s0 == KS);
            if (full) {
                if (buf == 0) {
                    mbar_wait(&bars[0], phase0);
                    phase0 ^= 1;
                } else {
                    mbar_wait(&bars[1], phase1);
                    phase1 ^= 1;
                }
            }
        };
        auto build_stage = [&](int64_t st, int buf) {
            const int64_t s0 = st * KS;
            const int valid = (int)min((int64_t)KS, args.m - s0);
            const double* vx = reinterpret_cast<const double*>(vbase + (size_t)buf * vbytes);
            const double* vy = reinterpret_cast<const double*>(vbase + (size_t)buf * vbytes + xbytes);
            double* out = theta + (size_t)(buf * 2 + tile) * kTileDoubles + col;
            for (int q = 0; q < spg; ++q) {
                const int s = grp * spg + q;
                double p = (s < valid) ? init : 0.0;
#pragma unroll
                for (int d = 0; d < NF; ++d) {
                    const uint16_t f = fac[d];
                    if (f != kNoFactor) {
                        const double v = (f & kFactorIsY) ? vy[s * n_t + (f & 0x7FFF)] : vx[s * n + f];
                        p *= v;
                    }
                }
                out[s * LDT] = p;
            }
        };
        auto mma_stage = [&](int buf) {
            const double* tI = theta + (size_t)(buf * 2 + 0) * kTileDoubles;
            const double* tJ = same ? tI : theta + (size_t)(buf * 2 + 1) * kTileDoubles;
            const double* pa = tI + (lane & 3) * LDT + wm * 32 + (lane >> 2);
            const double* pb = tJ + (lane & 3) * LDT + wn * 32 + (lane >> 2);
#pragma unroll
            for (int kk = 0; kk < KS / 4; ++kk) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = pa[kk * 4 * LDT + 8 * i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = pb[kk * 4 * LDT + 8 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        };

        // ---- software pipeline: TMA two stages ahead, Theta build one stage ahead of the MMAs --
        const int64_t b = seg.stage_begin, e = seg.stage_end;
        __syncthreads();  // previous segment's smem fully consumed
        load_stage(b, 0);
        if (b + 1 < e) load_stage(b + 1, 1);
        wait_stage(b, 0);
        __syncthreads();  // manual (ragged) loads visible
        build_stage(b, 0);
        __syncthreads();
        for (int64_t st = b; st < e; ++st) {
            const int cur = (int)((st - b) & 1);
            if (st + 2 < e) load_stage(st + 2, cur);  // v buffer `cur` was consumed by build(st)
            if (st + 1 < e) {
                wait_stage(st + 1, cur ^ 1);
                build_stage(st + 1, cur ^ 1);
            }
            mma_stage(cur);
            __syncthreads();
        }

        // ---- flush the 32x32 warp patch to this segment's partial-tile slot -----------------
        double* out = args.partials + (size_t)seg.slot * BT * BT;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = wm * 32 + 8 * i + (lane >> 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = wn * 32 + 8 * j + 2 * (lane & 3);
                *reinterpret_cast<double2*>(out + (size_t)row * BT + c) = make_double2(acc[i][j][0], acc[i][j][1]);
            }
        }
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
    plan = GramPlan();
    plan.bt = bt;
    plan.ks = ks;
    plan.nblk = (kaug + bt - 1) / bt;
    plan.npairs = plan.nblk * (plan.nblk + 1) / 2;
    plan.nctas = nctas;
    plan.nstages = (m + ks - 1) / ks;
    const int64_t ns = plan.nstages;

    struct Tmp {
        int cta, pair;
        int64_t b, e;
    };
    std::vector<Tmp> tmp;
    const int q = plan.npairs / nctas, r = plan.npairs % nctas;
    // full rounds: every CTA owns one block pair and all CTAs walk the sample stream in lockstep,
    // so each X/Y tile is fetched from HBM once and served to the other CTAs from L2.
    for (int c = 0; c < nctas; ++c)
        for (int i = 0; i < q; ++i) tmp.push_back({c, i * nctas + c, 0, ns});
    // remainder pairs: the (pair, stage) space is cut into nctas equal contiguous ranges.
    if (r > 0 && ns > 0) {
        const int64_t total = (int64_t)r * ns;
        for (int c = 0; c < nctas; ++c) {
            int64_t lo = total * c / nctas, hi = total * (c + 1) / nctas;
            while (lo < hi) {
                const int pr = (int)(lo / ns);
                const int64_t sb = lo % ns;
                const int64_t se = std::min<int64_t>(ns, sb + (hi - lo));
                tmp.push_back({c, q * nctas + pr, sb, se});
                lo += se - sb;
            }
        }
    }
    // slots: grouped by pair, inside a pair ordered by CTA (fixed summation order)
    std::vector<int> order(tmp.size());
    for (size_t i = 0; i < tmp.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (tmp[a].pair != tmp[b].pair) return tmp[a].pair < tmp[b].pair;
        if (tmp[a].cta != tmp[b].cta) return tmp[a].cta < tmp[b].cta;
        return tmp[a].b < tmp[b].b;
    });
    std::vector<int> slot_of(tmp.size());
    plan.pair_slot_begin.assign(plan.npairs + 1, 0);
    for (size_t i = 0; i < order.size(); ++i) {
        slot_of[order[i]] = (int)i;
        plan.pair_slot_begin[tmp[order[i]].pair + 1]++;
    }
    for (int p = 0; p < plan.npairs; ++p) plan.pair_slot_begin[p + 1] += plan.pair_slot_begin[p];
    plan.nslots = (int)tmp.size();

    // pair index -> (bi, bj), row-major over the lower triangle
    std::vector<std::pair<int, int>> pairs;
    for (int bi = 0; bi < plan.nblk; ++bi)
        for (int bj = 0; bj <= bi; ++bj) pairs.push_back({bi, bj});

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
template <int BT, int KS, int NF>
static int launch_gram(ddb_ctx* ctx, const GramArgs& args, size_t smem) {
    auto kern = gram_kernel<BT, KS, NF>;
    DDB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ctx->plan.nctas, GramCfg<BT>::kThreads, smem, ctx->stream>>>(args);
    DDB_CUDA_TRY(cudaGetLastError());
    return DDB_OK;
}

static int ensure_factor_table(ddb_ctx* ctx, int kaug_pad) {
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

int gram_run(ddb_ctx* ctx, double* dPacked) {
    if (ctx->k <= 0) {
        set_error("ddb_gram: no library installed (call ddb_set_library_monomial first)");
        return DDB_BAD_STATE;
    }
    if (!ctx->dX || !ctx->dY || ctx->m <= 0) {
        set_error("ddb_gram: no data bound (call ddb_upload / ddb_bind_device / ddb_generate first)");
        return DDB_BAD_STATE;
    }
    const int k = ctx->k, n = ctx->n, n_t = ctx->n_t;
    const int kaug = k + n_t;
    const int bt = (kaug <= 64) ? 64 : 128;
    const int ks = 16;
    if (n >= 0x7FFF || n_t >= 0x7FFF) {
        set_error("ddb_gram: more than 32766 states/targets are not supported by the monomial evaluator");
        return DDB_UNSUPPORTED;
    }
    const size_t smem = gram_smem_bytes(bt, ks, n, n_t);
    if (smem > 227 * 1024) {
        set_error("ddb_gram: n + n_t = %d is too wide to stage a sample tile in shared memory (%zu bytes)", n + n_t,
                  smem);
        return DDB_UNSUPPORTED;
    }
    if (((uintptr_t)ctx->dX & 15) || ((uintptr_t)ctx->dY & 15)) {
        set_error("ddb_gram: device data pointers must be 16-byte aligned");
        return DDB_INVALID_ARG;
    }
    // plan (cached per shape)
    if (ctx->plan_m != ctx->m || ctx->plan_kaug != kaug || ctx->plan.bt != bt) {
        build_gram_plan(ctx->plan, kaug, ctx->m, ctx->sm_count, bt, ks);
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

    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    const bool small_deg = ctx->max_degree <= 2;
    int st;
    if (bt == 64)
        st = small_deg ? launch_gram<64, 16, 2>(ctx, args, smem) : launch_gram<64, 16, 8>(ctx, args, smem);
    else
        st = small_deg ? launch_gram<128, 16, 2>(ctx, args, smem) : launch_gram<128, 16, 8>(ctx, args, smem);
    DDB_TRY(st);
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
    ctx->have_gram = (dPacked == ctx->d_packed.as<double>());
    return DDB_OK;
}

}  // namespace ddb
