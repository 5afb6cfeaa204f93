// PTX helpers and accumulator-patch shapes shared by the FP64 DMMA kernels (gram.cu, dmd.cu).
#pragma once
#include "common.cuh"

namespace ddb {

constexpr int kKS = 16;  // samples per stage
// raw-sample slot of the Gram kernels (doubles): [ones(16) | 0.0 | pad | X tile | Y tile], see gram.cu
constexpr int kVOnes = 0, kVZero = 16, kVX = 18;

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
    // tensor-pipe builder (gram_mb_kernel, moment schedule of degree-2 libraries): per block pair and warp a table of
    // per-lane operand / destination offsets (kMbPairWords words) and a flag byte per warp
    const uint32_t* optab = nullptr;
    const uint8_t* opcnt = nullptr;
    // prefix-product builder (gram_tree_kernel, single 64 x 64 diagonal pair): 4 programs x nsteps x {dst column,
    // parent column (255 = none: the slot's ones column), factor code}
    const uint32_t* tree = nullptr;
    // power-table builder (gram_pow_kernel, single 64 x 64 diagonal pair, at most 3 states): factor codes that point
    // into a per-stage table of powers x_i^1 .. x_i^pow_deg (pseudo-state i * pow_deg + e - 1), kaug_pad x kMaxFactors
    const uint16_t* pow_factors = nullptr;
    int pow_deg = 0;
    int swap_mode = 0;  // which of a CTA's two warps takes the heavy (26 sub-tiles) patch, see gram_pow_kernel
};

// Tensor-pipe builder, per warp and 16-sample stage: product ops (two in every k-step + a third in k-steps 0..2) as
// 32 lanes x {A offset | B offset << 16, destination 0 | destination 1 << 16}, then copy slots (library values that are
// a single raw value: one double per lane) as 32 lanes x {source | destination << 16}; all byte offsets.
constexpr int kMbOpSlots = 11;
constexpr int kMbCopySlots = 5;
constexpr int kMbWarpWords = kMbOpSlots * 64 + kMbCopySlots * 32;
constexpr int kMbPairWords = 8 * kMbWarpWords;
// Row pitch (doubles) of the X tile inside gram_mb_kernel's raw-sample slot: = 4 (mod 16), so that an op's operand
// load (8 consecutive states x 4 consecutive samples) touches every shared-memory bank once.  The Y tile keeps pitch
// n_t (targets are only ever copied).  Slot: [ones(16) | 0.0 | pad | X 16 x pitch | Y 16 x n_t].
__host__ __device__ inline int mb_pitch(int n) { return n + ((4 - n % 16) + 16) % 16; }
__host__ __device__ inline size_t gram_mb_vbytes(int n, int n_t) {
    size_t v = (size_t)(kVX + kKS * (mb_pitch(n) + n_t)) * sizeof(double);
    return (v + 127) & ~size_t(127);
}

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
// non-blocking phase test (the result comes back through the scoreboard: issue early, consume late)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        "  .reg .pred p;\n"
        "  mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "  selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
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

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Static shapes of live 8x8 accumulator sub-tiles inside a warp's 64x32 patch (bit i*4+j: rows 8i..,
// columns 8j..).  A diagonal block pair only needs sub-tiles on or below the diagonal; with 64x32
// patches every warp's patch is one of: entirely below (FULL), straddling with offset 0 (sub-tile
// column j live for rows i >= j), straddling with offset -4 (live for i >= j + 4), or entirely above
// (EMPTY).  FLAT is the 64x16 patch used when a block row has at most 64 live rows.
constexpr uint32_t shape_mask(int d, int jmax) {  // live iff j <= i + d and j < jmax
    uint32_t m = 0;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j)
            if (j <= i + d && j < jmax) m |= 1u << (i * 4 + j);
    return m;
}
constexpr uint32_t kShapeFull = shape_mask(8, 4);
constexpr uint32_t kShapeDiag0 = shape_mask(0, 4);
constexpr uint32_t kShapeDiagM4 = shape_mask(-4, 4);
constexpr uint32_t kShapeEmpty = 0u;
constexpr uint32_t kShapeFlat = shape_mask(8, 2);
// QUARTER: the first four sub-tile rows of a FLAT patch -> rows 0..31 of the block pair over all eight warps
constexpr uint32_t kShapeQuarter = kShapeFlat & 0xFFFFu;


}  // namespace ddb
