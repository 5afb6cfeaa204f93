// Fused Theta + Gram, "ring" schedule: producer SMs feed consumer SMs through an L2-resident ring.
//
// Why: in gram.cu every CTA multiplies out its own Theta tiles, so each library value is recomputed by
// the ~18 CTAs that share its block, and the FP64 multiplies interleaved with the DMMAs cost the tensor
// pipe ~13 % (profiles/dmma_probe_r1.json) on top of per-stage bookkeeping.  The TMA-fed kernel of dmd.cu
// shows the same MMA loop runs at 96 % of the DMMA peak when nobody multiplies on the side.  Here the two
// jobs are split inside ONE persistent launch (all CTAs co-resident, one per SM):
//   * a few PRODUCER CTAs walk the sample stream window by window (64 samples), stage X/Y in shared
//     memory, multiply every library term out exactly once and write the augmented rows
//     z(s) = [theta_1..theta_k, y_1..y_nt, 0-padding] into a ring in global memory that is small enough
//     to stay in L2 (ring rows are rewritten every few microseconds; Theta as a whole -- 184 GB at config
//     C3 -- is never materialised);
//   * every CONSUMER CTA owns one whole block pair for all samples: its TMA warp copies the two 16 x 128
//     tiles of each stage from the ring straight into the fragment layout, its eight MMA warps do nothing
//     but LDS + DMMA.8x8x4 (the dmd.cu pipeline).
// Flow control is two sets of monotone counters in global memory: ready[slot] (window index + 1, written
// by the producer with release semantics after its stores) and progress[consumer] (windows whose tiles
// have landed in that consumer's shared memory).  A producer reuses a ring slot only when every consumer
// has passed it.
// Consumers come in two kinds: PAIR consumers own one 128 x 128 block pair (2 x 4 warp patches of 64 x 32);
// STRIP consumers share a short last block row (<= 40 live rows, e.g. the 33 rows 2176..2208 of config C3)
// -- each takes the 40-row strip against up to 320 columns (8 warp patches of 40 x 40), so the ragged tail
// costs its true 5/16 of a block row instead of one consumer per block pair.  Block pairs that do not get
// a consumer (the half-empty diagonal ones) are finished by the self-building kernel of gram.cu in a
// second launch; one reduce kernel adds all partial tiles.
// The launch is cooperative: the spin-waits need every CTA co-resident, and the runtime refuses the
// launch instead of dead-locking when that cannot be guaranteed.
#include "common.cuh"
#include "mma_common.cuh"
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace ddb {

constexpr int kRingBT = 128;
constexpr int kRingPitch = kRingBT + 4;              // pair tile: [16 samples][132]
constexpr int kPairSamples = 32;                     // samples per stage of a pair consumer (one barrier round trip)
constexpr int kPairStage = 2 * kPairSamples * kRingPitch;  // doubles per stage of a pair consumer
constexpr int kPairDepth = 3;                        // its shared-memory stage ring
constexpr int kStripRows = 40;                       // live rows of a strip (5 row groups of 8)
constexpr int kStripPitch = kStripRows + 12;         // 52 = 4 (mod 16): conflict-free fragment loads
constexpr int kStripBlocks = 3;                      // block columns per strip consumer (8 warps x 48 columns)
constexpr int kStripStage = kKS * kStripPitch + kStripBlocks * kKS * kRingPitch;
constexpr int kStripDepth = 4;
constexpr int kRingSmemDoubles =
    kPairDepth * kPairStage > kStripDepth * kStripStage ? kPairDepth * kPairStage : kStripDepth * kStripStage;
constexpr int kMaxDepth = 8;
constexpr int kRingWarps = 8;                        // every warp multiplies; warp 0 also drives the TMA unit
constexpr int kRingThreads = kRingWarps * 32;
constexpr int kWin = 64;                             // samples per producer window (4 stages)
constexpr int kSPW = kWin / kKS;                     // stages per window
constexpr int kRingWindows = 16;                     // windows in the global ring
constexpr int kMaxProducers = 8;
constexpr int kMaxTermsPerThread = 12;               // ceil(rows / kRingThreads) supported by the producer

// what one consumer CTA does: kind 0 = block pair (a = block row, b = block column), kind 1 = strip
// (a = the short last block row, b = first block column, nb = number of block columns (<= kStripBlocks),
// slot0 = partial-tile slot of block column 0 of the strip's block row)
struct RingCons {
    int32_t kind, a, b, nb, slot0, pad0, pad1, pad2;
};

struct RingArgs {
    const double* X;
    const double* Y;
    int n, n_t, kaug, krow;   // krow = padded row count of the augmented library (multiple of 128)
    int nblk, last_pitch;     // block count; row pitch of the last block's chunk (kStripPitch in strip mode)
    int win_doubles;          // doubles per window of the global ring
    int64_t m;
    int nprod, ncons;
    int prod_bufs, prod_stride;  // producer staging tiles (1 or 2) and their stride in doubles
    const uint16_t* factors;
    const RingCons* cons;
    // global ring, kRingWindows windows of [block][sample 0..63][pitch]: one (block, stage) operand tile is
    // ONE contiguous chunk of 16 x pitch doubles already in the padded shared-memory layout -> one bulk copy
    double* zring;
    unsigned long long* done;      // [kMaxProducers] windows finished by producer p (it owns w = p, p + nprod, ...)
    unsigned long long* progress;  // [ncons] windows copied out of the global ring by consumer c
    double* partials;         // pair consumer c -> slot c; strips -> slot0 + block column
};

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}

// The TMA side of a consumer, run by one warp at a time between its DMMAs: lane l < nops copies operand tile l of a
// stage (one contiguous chunk of the global ring) into the stage ring.
struct RingFeed {
    const double* zring;
    const unsigned long long* done;
    uint64_t* full;
    uint64_t* empty;
    double* ring;
    int64_t nstages;
    uint32_t* s_ready;         // shared watermark: every window below it is complete in the global ring
    uint32_t ready;            // this warp's copy of the watermark
    int nprod, lane, win_doubles;
    uint32_t spw;              // stages per producer window (4 for 16-sample stages, 2 for 32-sample stages)
    int src_off, src_stage, dst_off;  // this lane's operand: offset in a window, per-stage stride, offset in a stage
    uint32_t bytes, stage_bytes;
    long long t_ready, t_empty;

    template <int DEPTH, int STAGE>
    __device__ __forceinline__ void issue(int64_t st) {
        if (st >= nstages) return;
        const int slot = (int)((uint32_t)st % (uint32_t)DEPTH);  // st < 2^32 (checked on the host)
        const uint32_t w = (uint32_t)st / spw;
        if (w >= ready) {
            // Windows [0, ready) are known to be complete.  The watermark is shared through shared memory
            // (any warp may have refreshed it); whoever finds it short polls the producers' counters:
            // producer p has finished windows p, p + nprod, ... (done[p] of them), so every window below
            // min_p (done[p] * nprod + p) is complete.
#ifdef DDB_RING_TIMING
            const long long t0 = clock64();
#endif
            uint32_t r;
            asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(r) : "r"(smem_u32(s_ready)) : "memory");
            if (w >= r) {
                do {
                    r = lane < nprod ? (uint32_t)ld_acquire(done + lane) * (uint32_t)nprod + (uint32_t)lane : 0xFFFFFFFFu;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) r = min(r, __shfl_xor_sync(0xFFFFFFFFu, r, o));
                } while (w >= r);
                if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(s_ready)), "r"(r) : "memory");
            }
            ready = r;
            // the producers wrote through the generic proxy, the bulk copies below read through the async proxy
            asm volatile("fence.proxy.async;" ::: "memory");
#ifdef DDB_RING_TIMING
            t_ready += clock64() - t0;
#endif
        }
        if (st >= DEPTH) {
#ifdef DDB_RING_TIMING
            const long long t0 = clock64();
#endif
            mbar_wait(&empty[slot], (((uint32_t)st / (uint32_t)DEPTH) & 1u) ^ 1u);
#ifdef DDB_RING_TIMING
            t_empty += clock64() - t0;
#endif
        }
        if (lane == 0) mbar_expect_tx(&full[slot], stage_bytes);
        __syncwarp();
        if (bytes) {
            const double* src = zring + (size_t)(w % kRingWindows) * win_doubles + src_off + ((uint32_t)st % spw) * src_stage;
            bulk_g2s(ring + (size_t)slot * STAGE + dst_off, src, bytes, &full[slot]);
        }
    }
};

// One consumer warp: RI x CJ accumulator sub-tiles (live ones in MASK, bit i * CJ + j), fragments
// double-buffered across k-steps AND across stages: the loads of the next stage's first k-step are issued
// before the last DMMAs of the current stage whenever its tile has already landed.
// BSPLIT: column sub-tile j lives in operand tile j / 2 (strip consumers: two 8-column groups per block)
template <int RI, int CJ, uint32_t MASK, int DEPTH, int STAGE, bool BSPLIT, int KSTEPS>
struct RingWarp {
    static_assert(KSTEPS % 2 == 0, "fragments alternate between two buffers");
    double acc[RI][CJ][2];
    __device__ __forceinline__ static constexpr bool row_live(int i) {
        uint32_t m = 0;
        for (int j = 0; j < CJ; ++j) m |= (MASK >> (i * CJ + j)) & 1u;
        return m != 0;
    }
    __device__ __forceinline__ static constexpr bool col_live(int j) {
        uint32_t m = 0;
        for (int i = 0; i < RI; ++i) m |= (MASK >> (i * CJ + j)) & 1u;
        return m != 0;
    }
    __device__ __forceinline__ void load(double (&a)[RI], double (&b)[CJ], const double* pa, const double* pb) {
#pragma unroll
        for (int i = 0; i < RI; ++i)
            if (row_live(i)) a[i] = pa[8 * i];
#pragma unroll
        for (int j = 0; j < CJ; ++j)
            if (col_live(j)) b[j] = BSPLIT ? pb[(j >> 1) * (kKS * kRingPitch) + (j & 1) * 8] : pb[8 * j];
    }
    __device__ __forceinline__ void mma(const double (&a)[RI], const double (&b)[CJ]) {
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
            for (int j = 0; j < CJ; ++j)
                if ((MASK >> (i * CJ + j)) & 1u) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    // offA/offB: this lane's fragment offsets inside a stage; pitch4A/B = 4 sample rows of the operand tile.
    // The warps take turns driving the TMA unit (warp st % 8 requests stage st + DEPTH - 2 at the top of its
    // stage st): a fixed feeder warp would always be the slowest one and, being alone in setting the pace,
    // expose its bookkeeping on the tensor pipe.  The slot being refilled was released a whole stage ago by
    // every warp, so the wait inside issue() does not normally stall.
    __device__ __forceinline__ void run(const double* ring, uint64_t* full, uint64_t* empty, int64_t nstages,
                                        int offA, int offB, int pitch4A, int pitch4B, const int warp, RingFeed& fd,
                                        unsigned long long* progress, int lane, long long& t_wait) {
        constexpr int kAhead = DEPTH - 2;
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
            for (int j = 0; j < CJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        if (warp == 0)
            for (int q = 0; q < kAhead; ++q) fd.template issue<DEPTH, STAGE>(q);
        double a0[RI], b0[CJ], a1[RI], b1[CJ];
        mbar_wait(&full[0], 0u);
        if (MASK != 0u) load(a0, b0, ring + offA, ring + offB);
        int slot = 0;
        uint32_t par = 0;
        for (int64_t st = 0; st < nstages; ++st) {
            if (((uint32_t)st & (kRingWarps - 1)) == (uint32_t)warp) {  // this warp's turn to drive the TMA unit
                // the last stage of a window has landed in shared memory: the global ring slot may be recycled
                if (lane == 0 && ((uint32_t)st % fd.spw) == fd.spw - 1) st_release(progress, (unsigned long long)((uint32_t)st / fd.spw + 1));
                fd.template issue<DEPTH, STAGE>(st + kAhead);
            }
            const double* base = ring + (size_t)slot * STAGE;
            int nslot = slot + 1;
            uint32_t npar = par;
            if (nslot == DEPTH) {
                nslot = 0;
                npar ^= 1u;
            }
            const double* nbase = ring + (size_t)nslot * STAGE;
            if (MASK != 0u) {
#pragma unroll
                for (int kk = 0; kk < KSTEPS - 1; ++kk) {
                    if (kk & 1) {
                        load(a0, b0, base + offA + (kk + 1) * pitch4A, base + offB + (kk + 1) * pitch4B);
                        mma(a1, b1);
                    } else {
                        load(a1, b1, base + offA + (kk + 1) * pitch4A, base + offB + (kk + 1) * pitch4B);
                        mma(a0, b0);
                    }
                }
            }
            if (st + 1 < nstages) {
                if (MASK != 0u && mbar_test(&full[nslot], npar)) {
                    load(a0, b0, nbase + offA, nbase + offB);
                    mma(a1, b1);
                } else {
                    if (MASK != 0u) mma(a1, b1);
#ifdef DDB_RING_TIMING
                    const long long t0 = clock64();
#endif
                    mbar_wait(&full[nslot], npar);
#ifdef DDB_RING_TIMING
                    t_wait += clock64() - t0;
#endif
                    if (MASK != 0u) load(a0, b0, nbase + offA, nbase + offB);
                }
            } else if (MASK != 0u) {
                mma(a1, b1);
            }
            // every fragment of this stage is in registers (the DMMAs above consumed them)
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            slot = nslot;
            par = npar;
        }
    }
};

template <int NF>
__global__ void __launch_bounds__(kRingThreads, 1) gram_ring_kernel(const RingArgs args) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nwin = (args.m + kWin - 1) / kWin;
    const int krow = args.krow;

    if ((int)blockIdx.x < args.nprod) {
        // =============================== PRODUCER CTA ==============================================
        // staging tile (doubles): [ones(kWin) | 0.0 | pad | X (kWin x n) | Y (kWin x n_t)], sample-major
        // exactly as in HBM.  A missing factor reads ones[s] (stride 1), a padding row reads the 0.0
        // (stride 0): every entry is a branch-free product of NF staged values.
        const int n = args.n, n_t = args.n_t;
        double* v = reinterpret_cast<double*>(smem_raw);
        constexpr int kOnes = 0, kZero = kWin, kX = kWin + 2;
        // this thread owns the adjacent term pairs 2*tid + {0,1} + 2*kRingThreads*q -> 16-byte stores
        constexpr int kPairs = kMaxTermsPerThread / 2;
        int off[kPairs][2][NF], str[kPairs][2][NF];
        int dsto[kPairs], dstp[kPairs];  // destination of the pair inside a window: offset of sample 0, row pitch
        const int npair_rows = krow / 2;
        const int my_pairs = (npair_rows - tid + kRingThreads - 1) / kRingThreads;
#pragma unroll
        for (int q = 0; q < kPairs; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int term = 2 * (tid + q * kRingThreads) + h;
                if (h == 0) {
                    const int blk = term / kRingBT;
                    dstp[q] = (blk == args.nblk - 1) ? args.last_pitch : kRingPitch;
                    dsto[q] = blk * (kWin * kRingPitch) + (term % kRingBT);
                }
#pragma unroll
                for (int d = 0; d < NF; ++d) {
                    off[q][h][d] = kOnes;
                    str[q][h][d] = 1;
                }
                if (q < my_pairs) {
                    if (term < args.kaug) {
                        const uint4 w4 = *reinterpret_cast<const uint4*>(args.factors + (size_t)term * kMaxFactors);
                        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int d = 0; d < NF; ++d) {
                            const uint32_t f = (w[d >> 1] >> (16 * (d & 1))) & 0xFFFFu;
                            if (f != kNoFactor) {
                                const bool isy = (f & kFactorIsY) != 0;
                                off[q][h][d] = isy ? kX + kWin * n + (int)(f & 0x7FFF) : kX + (int)f;
                                str[q][h][d] = isy ? n_t : n;
                            }
                        }
                    } else {
                        off[q][h][0] = kZero;
                        str[q][h][0] = 0;
                    }
                }
            }
        // two staging tiles: the X/Y rows of this producer's next window are fetched with cp.async while
        // the current one is multiplied out (one tile when two do not fit in shared memory)
        const int vstride = args.prod_stride;
        for (int e = tid; e < kWin + 2; e += kRingThreads) {
            v[e] = (e < kWin) ? 1.0 : 0.0;
            if (args.prod_bufs == 2) v[vstride + e] = (e < kWin) ? 1.0 : 0.0;
        }
        auto stage_window = [&](int64_t w, double* vb) {
            const int64_t s0 = w * kWin;
            const int valid = (int)min((int64_t)kWin, args.m - s0);
            const double* gx = args.X + s0 * n;
            const double* gy = args.Y + s0 * n_t;
            const int nx = valid * n, ny = valid * n_t;
            if ((((uintptr_t)gx | (uintptr_t)gy) & 15) == 0 && ((nx | ny) & 1) == 0) {
                for (int e = 2 * tid; e < nx; e += 2 * kRingThreads) cp_async16(vb + kX + e, gx + e);
                for (int e = 2 * tid; e < ny; e += 2 * kRingThreads) cp_async16(vb + kX + kWin * n + e, gy + e);
            } else {
                for (int e = tid; e < nx; e += kRingThreads) cp_async8(vb + kX + e, gx + e);
                for (int e = tid; e < ny; e += kRingThreads) cp_async8(vb + kX + kWin * n + e, gy + e);
            }
            if (valid < kWin) {  // ragged last window: zero rows, and a zero "ones" column for them
                for (int e = nx + tid; e < kWin * n; e += kRingThreads) vb[kX + e] = 0.0;
                for (int e = ny + tid; e < kWin * n_t; e += kRingThreads) vb[kX + kWin * n + e] = 0.0;
                for (int e = tid; e < kWin; e += kRingThreads) vb[kOnes + e] = (e < valid) ? 1.0 : 0.0;
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#ifdef DDB_RING_TIMING
        long long tp_wait = 0, tp_all = clock64(), tp0;
#endif
        int cur = 0;
        if ((int64_t)blockIdx.x < nwin) stage_window(blockIdx.x, v);
        for (int64_t w = blockIdx.x; w < nwin; w += args.nprod) {
            const double* vb = v + cur * vstride;
            const bool ahead = args.prod_bufs == 2 && w + args.nprod < nwin;
            if (ahead) stage_window(w + args.nprod, v + (cur ^ 1) * vstride);
#ifdef DDB_RING_TIMING
            tp0 = clock64();
#endif
            // the ring slot is free once every consumer has copied window w - kRingWindows out of it
            if (w >= kRingWindows) {
                const unsigned long long need = (unsigned long long)(w - kRingWindows + 1);
                while (true) {
                    int ok = 1;
                    for (int c = tid; c < args.ncons; c += kRingThreads) ok &= (ld_acquire(args.progress + c) >= need);
                    if (__syncthreads_and(ok)) break;
                    __nanosleep(100);
                }
            }
#ifdef DDB_RING_TIMING
            tp_wait += clock64() - tp0;
#endif
            if (ahead)
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            else
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            double* zwin = args.zring + (size_t)(w % kRingWindows) * args.win_doubles;
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                if (q < my_pairs) {
                    // a short last block only keeps its first last_pitch - 12 columns (the rest is padding)
                    const bool keep = dstp[q] == kRingPitch || (2 * (tid + q * kRingThreads)) % kRingBT < dstp[q] - 12;
                    double* dst = zwin + dsto[q];
#pragma unroll 4
                    for (int s = 0; s < kWin; ++s) {
                        double p0 = vb[off[q][0][0] + s * str[q][0][0]];
                        double p1 = vb[off[q][1][0] + s * str[q][1][0]];
#pragma unroll
                        for (int d = 1; d < NF; ++d) {
                            p0 *= vb[off[q][0][d] + s * str[q][0][d]];
                            p1 *= vb[off[q][1][d] + s * str[q][1][d]];
                        }
                        if (keep) *reinterpret_cast<double2*>(dst + s * dstp[q]) = make_double2(p0, p1);
                    }
                }
            }
            __syncthreads();  // all rows of the window written (and the staging tile free for the next window)
            if (tid == 0) {
                __threadfence();
                st_release(args.done + blockIdx.x, (unsigned long long)(w / args.nprod + 1));
            }
            if (args.prod_bufs == 2)
                cur ^= 1;
            else if (w + args.nprod < nwin)
                stage_window(w + args.nprod, v);
        }
#ifdef DDB_RING_TIMING
        if (tid == 0) printf("producer %d: all %lld wait %lld\n", blockIdx.x, clock64() - tp_all, tp_wait);
#endif
        return;
    }

    // =================================== CONSUMER CTA =================================================
    const int c = blockIdx.x - args.nprod;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // [kMaxDepth] stage landed
    uint64_t* empty = full + kMaxDepth;                      // [kMaxDepth] every warp is done with the slot
    uint32_t* s_ready = reinterpret_cast<uint32_t*>(smem_raw + 128);
    double* ring = reinterpret_cast<double*>(smem_raw + 256);
    const RingCons cd = args.cons[c];
    const bool strip = cd.kind == 1;
    const bool same = !strip && cd.a == cd.b;
    if (tid == 0) {
        for (int i = 0; i < kMaxDepth; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kRingWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *s_ready = 0u;
    }
    if (strip)  // block columns past the strip's range are never copied: they must read as zero, not as garbage
        for (int e = tid; e < kStripDepth * kStripStage; e += kRingThreads) ring[e] = 0.0;
    __syncthreads();
    const int64_t nstages = nwin * (kWin / (cd.kind == 1 ? kKS : kPairSamples));  // producers zero-fill the tail

    RingFeed feed = {};
    {
        // operand tiles of a stage: pair = {block a, block b}; strip = {short last block, nb block columns}
        const int chunk = kWin * kRingPitch;  // doubles per full block of a window
        const int nops = strip ? 1 + cd.nb : (same ? 1 : 2);
        const int sps = strip ? kKS : kPairSamples;  // samples per stage
        if (lane < nops) {
            const int blk = strip ? (lane == 0 ? cd.a : cd.b + lane - 1) : (lane == 0 ? cd.a : cd.b);
            const int pitch = (blk == args.nblk - 1) ? args.last_pitch : kRingPitch;
            feed.src_off = blk * chunk;
            feed.src_stage = sps * pitch;
            feed.bytes = (uint32_t)(sps * pitch * sizeof(double));
            feed.dst_off = strip ? (lane == 0 ? 0 : kKS * kStripPitch + (lane - 1) * kKS * kRingPitch) : lane * kPairSamples * kRingPitch;
        }
        feed.spw = (uint32_t)(kWin / sps);
        uint32_t tot = feed.bytes;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
        feed.stage_bytes = tot;
        feed.zring = args.zring;
        feed.done = args.done;
        feed.full = full;
        feed.empty = empty;
        feed.ring = ring;
        feed.nstages = nstages;
        feed.s_ready = s_ready;
        feed.ready = 0u;
        feed.win_doubles = args.win_doubles;
        feed.nprod = args.nprod;
        feed.lane = lane;
        feed.t_ready = feed.t_empty = 0;
    }

    const int g = lane >> 2, t = lane & 3;
    long long t_wait = 0;
#ifdef DDB_RING_TIMING
    const long long tm_all = clock64();
#endif
    if (!strip) {
        const int wm = (0x4B >> warp) & 1;  // warps 0..7 -> (1,0)(1,1)(0,0)(1,2)(0,2)(0,3)(1,3)(0,1), see gram.cu
        const int wn = (0x7E84 >> (2 * warp)) & 3;
        const int row0 = wm * 64, col0 = wn * 32;
        int shape = 0;
        if (same) {
            const int d = (row0 - col0) / 8;
            shape = (d >= 3) ? 0 : (d == 0 ? 1 : (d == -4 ? 2 : 3));
        }
        const int offA = t * kRingPitch + row0 + g;
        const int offB = (same ? 0 : kPairSamples * kRingPitch) + t * kRingPitch + col0 + g;
        double* out = args.partials + (size_t)c * kRingBT * kRingBT;
        auto go = [&](auto tag) {
            constexpr uint32_t MASK = decltype(tag)::value;
            RingWarp<8, 4, MASK, kPairDepth, kPairStage, false, kPairSamples / 4> w;
            w.run(ring, full, empty, nstages, offA, offB, 4 * kRingPitch, 4 * kRingPitch, warp, feed, args.progress + c, lane,
                  t_wait);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((MASK >> (i * 4 + j)) & 1u)
                        *reinterpret_cast<double2*>(out + (size_t)(row0 + 8 * i + g) * kRingBT + col0 + 8 * j + 2 * t) =
                            make_double2(w.acc[i][j][0], w.acc[i][j][1]);
        };
        switch (shape) {
            case 0: go(std::integral_constant<uint32_t, kShapeFull>()); break;
            case 1: go(std::integral_constant<uint32_t, kShapeDiag0>()); break;
            case 2: go(std::integral_constant<uint32_t, kShapeDiagM4>()); break;
            default: go(std::integral_constant<uint32_t, 0u>()); break;
        }
    } else {
        // strip: 5 row groups x 6 column groups per warp -- warp w owns the 8-column groups 2w, 2w + 1 of each
        // of the (up to) three block columns, so its B fragments are affine in the sub-tile index
        const int offA = t * kStripPitch + g;
        const int offB = kKS * kStripPitch + t * kRingPitch + warp * 16 + g;
        RingWarp<5, 6, 0x3FFFFFFFu, kStripDepth, kStripStage, true, kKS / 4> w;
        w.run(ring, full, empty, nstages, offA, offB, 4 * kStripPitch, 4 * kRingPitch, warp, feed, args.progress + c, lane,
              t_wait);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if ((j >> 1) < cd.nb) {
                double* out = args.partials + (size_t)(cd.slot0 + cd.b + (j >> 1)) * kRingBT * kRingBT + warp * 16 +
                              (j & 1) * 8 + 2 * t;
#pragma unroll
                for (int i = 0; i < 5; ++i)
                    *reinterpret_cast<double2*>(out + (size_t)(8 * i + g) * kRingBT) =
                        make_double2(w.acc[i][j][0], w.acc[i][j][1]);
            }
        }
    }
#ifdef DDB_RING_TIMING
    if (lane == 0 && ((c % 37) == 0 || strip) && (warp == 0 || warp == 5))
        printf("consumer %d kind %d warp %d: all %lld full-wait %lld ready-wait %lld empty-wait %lld\n", c, cd.kind, warp,
               clock64() - tm_all, t_wait, feed.t_ready, feed.t_empty);
#endif
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
int gram_launch_bt128(ddb_ctx* ctx, const GramArgs& args, int nctas, bool offdiag_only);
void gram_reduce_launch(ddb_ctx* ctx, const double* partials, const int2* pairs, const int32_t* pair_slot_begin,
                        int npairs, int bt, int k, int n_t, double* packed);
int gram_ensure_factors(ddb_ctx* ctx, int rows);

struct RingState {
    DevBuf zring, flags, plan, partials;
    int plan_kaug = -1, plan_nctas = -1;
    int64_t plan_m = -1;
    int nprod = 0, ncons = 0, npairs = 0, nslots = 0;
    int strip_slot0 = 0, last_pitch = 0, win_doubles = 0;
    size_t off_pairs = 0, off_psb = 0, off_cons = 0, off_segs = 0, off_cta = 0;
    bool have_rest = false;
};

double gram_plain_dmma_flops(int kaug, int bt);

int gram_ring_run(ddb_ctx* ctx, double* dPacked, bool* handled) {
    *handled = false;
    const char* env = getenv("DDB_GRAM_RING");
    if (env && atoi(env) == 0) return DDB_OK;
    const int k = ctx->k, n = ctx->n, n_t = ctx->n_t, kaug = k + n_t;
    const int nblk = (kaug + kRingBT - 1) / kRingBT;
    const int krow = nblk * kRingBT;
    int nprod = 4;
    if (const char* e = getenv("DDB_RING_NPROD")) nprod = std::max(1, std::min(kMaxProducers, atoi(e)));  // tuning knob
    const int ncons = ctx->sm_count - nprod;
    const int npairs = nblk * (nblk + 1) / 2;
    // eligible when every consumer SM gets work for the whole stream, the producer can hold its terms in
    // registers and its staging tile fits
    if (npairs < ncons || krow > kMaxTermsPerThread * kRingThreads) return DDB_OK;
    if ((size_t)(kWin + 2 + kWin * (n + n_t)) * sizeof(double) > 200 * 1024) return DDB_OK;
    if (ctx->m < 4 * kWin * kRingWindows) return DDB_OK;  // not worth the pipeline fill
    if (ctx->m / kKS >= (int64_t)1 << 31) return DDB_OK;  // the consumers count stages in 32 bits
    int coop = 0;
    DDB_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
    if (!coop) return DDB_OK;

    if (!ctx->ring) ctx->ring = new RingState();
    RingState* R = ctx->ring;
    const bool small_deg = ctx->max_degree <= 2;
    if (R->plan_kaug != kaug || R->plan_m != ctx->m || R->plan_nctas != ctx->sm_count) {
        // a short last block row is shared by strip consumers instead of costing one consumer per pair
        const int last_rows = kaug - (nblk - 1) * kRingBT;
        const bool use_strip = last_rows <= kStripRows && nblk >= 2;
        const int nstrip = use_strip ? (nblk - 1 + kStripBlocks - 1) / kStripBlocks : 0;
        // pairs by decreasing cost: the first `ncons - nstrip` get a whole consumer CTA
        std::vector<std::pair<int, int>> pairs;
        std::vector<double> cost;
        for (int bi = 0; bi < nblk; ++bi)
            for (int bj = 0; bj <= bi; ++bj) {
                if (use_strip && bi == nblk - 1 && bj < bi) continue;  // covered by the strips
                pairs.push_back({bi, bj});
                const int rows_valid = std::min(kRingBT, kaug - bi * kRingBT);
                cost.push_back(rows_valid <= 64 ? 0.53 : (bi == bj ? 0.60 : 1.0));
            }
        const int np = (int)pairs.size();
        const int npaircons = std::min(ncons - nstrip, np);
        if (npaircons + nstrip != ncons) return DDB_OK;  // would leave consumer SMs idle
        std::vector<int> order(np);
        for (int p = 0; p < np; ++p) order[p] = p;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
        std::vector<int2> all_pairs;
        std::vector<int32_t> psb(1, 0);
        std::vector<RingCons> cons;
        for (int i = 0; i < npaircons; ++i) {
            const auto pr = pairs[order[i]];
            all_pairs.push_back(make_int2(pr.first, pr.second));
            psb.push_back(i + 1);
            cons.push_back({0, pr.first, pr.second, 0, 0, 0, 0, 0});
        }
        int nslots = npaircons;
        R->strip_slot0 = nslots;
        if (use_strip) {
            for (int bj = 0; bj < nblk - 1; ++bj) {  // one slot per block column of the last block row
                all_pairs.push_back(make_int2(nblk - 1, bj));
                psb.push_back(++nslots);
            }
            for (int b0 = 0; b0 < nblk - 1; b0 += kStripBlocks)
                cons.push_back({1, nblk - 1, b0, std::min(kStripBlocks, nblk - 1 - b0), R->strip_slot0, 0, 0, 0});
        }
        R->last_pitch = use_strip ? kStripPitch : kRingPitch;
        R->win_doubles = (nblk - 1) * kWin * kRingPitch + kWin * R->last_pitch;
        std::vector<std::pair<int, int>> rest;
        std::vector<double> rest_cost;
        for (int i = npaircons; i < np; ++i) {
            rest.push_back(pairs[order[i]]);
            rest_cost.push_back(cost[order[i]]);
        }
        GramPlan pb;
        R->have_rest = !rest.empty();
        if (R->have_rest) {
            const int nrest = (int)rest.size();
            const int nranges = ctx->sm_count / nrest;
            const int64_t ns = (ctx->m + kKS - 1) / kKS;
            if (nranges >= 2 && ns >= nranges) {
                // Lock-step plan: the sample stream is cut into `nranges` ranges and CTA r * nrest + p runs pair p
                // over range r, so the nrest CTAs of a range walk the same samples at the same time and every
                // X / Y tile comes from HBM once and from L2 for the other pairs (the cost-balanced stream split
                // of build_plan_from_pairs made each pair's CTAs read the whole input again: 16 x the traffic).
                pb = GramPlan();
                pb.bt = kRingBT;
                pb.ks = kKS;
                pb.npairs = nrest;
                pb.nctas = ctx->sm_count;
                pb.nstages = ns;
                pb.pairs = rest;
                pb.cta_seg_begin.assign(ctx->sm_count + 1, 0);
                pb.pair_slot_begin.assign(nrest + 1, 0);
                for (int p = 0; p <= nrest; ++p) pb.pair_slot_begin[p] = p * nranges;
                for (int c = 0; c < ctx->sm_count; ++c) {
                    pb.cta_seg_begin[c] = (int32_t)pb.segments.size();
                    if (c < nranges * nrest) {
                        const int r = c / nrest, p = c % nrest;
                        GramSegment sg;
                        sg.bi = rest[p].first;
                        sg.bj = rest[p].second;
                        sg.slot = p * nranges + r;
                        sg.pad = 0;
                        sg.stage_begin = ns * r / nranges;
                        sg.stage_end = ns * (r + 1) / nranges;
                        pb.segments.push_back(sg);
                    }
                }
                pb.cta_seg_begin[ctx->sm_count] = (int32_t)pb.segments.size();
                pb.nslots = nrest * nranges;
            } else {
                build_plan_from_pairs(pb, rest, rest_cost, ctx->m, ctx->sm_count, kRingBT, kKS);
            }
            for (auto& sgm : pb.segments) sgm.slot += nslots;
            for (size_t j = 0; j < rest.size(); ++j) {
                all_pairs.push_back(make_int2(rest[j].first, rest[j].second));
                psb.push_back(nslots + pb.pair_slot_begin[j + 1]);
            }
            nslots += pb.nslots;
        }
        R->npairs = (int)all_pairs.size();
        R->nslots = nslots;
        auto align16 = [](size_t x) { return (x + 15) & ~size_t(15); };
        R->off_pairs = 0;
        R->off_psb = all_pairs.size() * sizeof(int2);
        R->off_cons = align16(R->off_psb + psb.size() * sizeof(int32_t));
        R->off_segs = align16(R->off_cons + cons.size() * sizeof(RingCons));
        R->off_cta = R->off_segs + pb.segments.size() * sizeof(GramSegment);
        const size_t total = R->off_cta + (pb.cta_seg_begin.size() + 1) * sizeof(int32_t);
        std::vector<unsigned char> host(total, 0);
        memcpy(host.data(), all_pairs.data(), all_pairs.size() * sizeof(int2));
        memcpy(host.data() + R->off_psb, psb.data(), psb.size() * sizeof(int32_t));
        memcpy(host.data() + R->off_cons, cons.data(), cons.size() * sizeof(RingCons));
        if (R->have_rest) {
            memcpy(host.data() + R->off_segs, pb.segments.data(), pb.segments.size() * sizeof(GramSegment));
            memcpy(host.data() + R->off_cta, pb.cta_seg_begin.data(), pb.cta_seg_begin.size() * sizeof(int32_t));
        }
        DDB_TRY(R->plan.reserve(total));
        DDB_CUDA_TRY(cudaMemcpyAsync(R->plan.p, host.data(), total, cudaMemcpyHostToDevice, ctx->stream));
        DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        DDB_TRY(R->partials.reserve((size_t)R->nslots * kRingBT * kRingBT * sizeof(double)));
        DDB_TRY(R->zring.reserve((size_t)kRingWindows * R->win_doubles * sizeof(double)));
        DDB_TRY(R->flags.reserve((size_t)(kMaxProducers + ncons) * sizeof(unsigned long long)));
        R->plan_kaug = kaug;
        R->plan_m = ctx->m;
        R->plan_nctas = ctx->sm_count;
        R->nprod = nprod;
        R->ncons = ncons;
        ctx->factors_nt = -1;
    }
    if (ctx->factors_nt != n_t) DDB_TRY(gram_ensure_factors(ctx, krow));
    if (!dPacked) {
        DDB_TRY(ctx->d_packed.reserve((size_t)ddb_gram_count(ctx) * sizeof(double)));
        dPacked = ctx->d_packed.as<double>();
    }
    unsigned char* base = static_cast<unsigned char*>(R->plan.p);
    RingArgs a;
    a.X = ctx->dX;
    a.Y = ctx->dY;
    a.n = n;
    a.n_t = n_t;
    a.kaug = kaug;
    a.krow = krow;
    a.nblk = nblk;
    a.last_pitch = R->last_pitch;
    a.win_doubles = R->win_doubles;
    a.m = ctx->m;
    a.nprod = R->nprod;
    a.ncons = R->ncons;
    a.factors = ctx->d_factors.as<uint16_t>();
    a.cons = reinterpret_cast<const RingCons*>(base + R->off_cons);
    a.zring = R->zring.as<double>();
    a.done = R->flags.as<unsigned long long>();
    a.progress = a.done + kMaxProducers;
    a.partials = R->partials.as<double>();
    const size_t smem_cons = 256 + (size_t)kRingSmemDoubles * sizeof(double);
    const int prod_stride = (kWin + 2 + kWin * (n + n_t) + 1) & ~1;
    a.prod_stride = prod_stride;
    a.prod_bufs = (size_t)2 * prod_stride * sizeof(double) <= smem_cons ? 2 : 1;
    const size_t smem_prod = (size_t)a.prod_bufs * prod_stride * sizeof(double);
    const size_t smem = std::max(smem_cons, smem_prod);
    const void* kern = small_deg ? (const void*)gram_ring_kernel<2> : (const void*)gram_ring_kernel<8>;
    DDB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    DDB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRingThreads, smem));
    if (per_sm < 1) return DDB_OK;
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    DDB_CUDA_TRY(cudaMemsetAsync(R->flags.p, 0, (size_t)(kMaxProducers + R->ncons) * sizeof(unsigned long long), ctx->stream));
    void* kargs[] = {&a};
    DDB_CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(R->nprod + R->ncons), dim3(kRingThreads), kargs, smem, ctx->stream));
    int launches = 2;
    if (R->have_rest) {
        GramArgs g;
        g.X = ctx->dX;
        g.Y = ctx->dY;
        g.n = n;
        g.n_t = n_t;
        g.m = ctx->m;
        g.kaug = kaug;
        g.factors = ctx->d_factors.as<uint16_t>();
        g.segs = reinterpret_cast<const GramSegment*>(base + R->off_segs);
        g.cta_seg_begin = reinterpret_cast<const int32_t*>(base + R->off_cta);
        g.partials = R->partials.as<double>();
        DDB_TRY(gram_launch_bt128(ctx, g, ctx->sm_count, false));
        ++launches;
    }
    gram_reduce_launch(ctx, R->partials.as<double>(), reinterpret_cast<const int2*>(base + R->off_pairs),
                       reinterpret_cast<const int32_t*>(base + R->off_psb), R->npairs, kRingBT, k, n_t, dPacked);
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->timings.gram_ms = ms;
    ctx->timings.gram_launches = launches;
    ctx->timings.gram_schedule = 1;
    ctx->timings.gram_dmma_flops = gram_plain_dmma_flops(kaug, kRingBT);
    ctx->have_gram = (dPacked == ctx->d_packed.as<double>());
    ctx->imp_k = 0;
    *handled = true;
    return DDB_OK;
}

void ring_free(ddb_ctx* ctx) {
    if (ctx->ring) {
        ctx->ring->zring.release();
        ctx->ring->flags.release();
        ctx->ring->plan.release();
        ctx->ring->partials.release();
        delete ctx->ring;
        ctx->ring = nullptr;
    }
}

}  // namespace ddb
