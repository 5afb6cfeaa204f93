// Internal declarations shared by the translation units of libddb200.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/ddb.h"

namespace ddb {

void set_error(const char* fmt, ...);

#define DDB_CUDA_TRY(expr)                                                                        \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            ::ddb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,   \
                             __LINE__);                                                           \
            return (e__ == cudaErrorMemoryAllocation) ? DDB_OUT_OF_MEMORY : DDB_CUDA_ERROR;       \
        }                                                                                         \
    } while (0)

#define DDB_TRY(expr)                 \
    do {                              \
        int s__ = (expr);             \
        if (s__ != DDB_OK) return s__; \
    } while (0)

// A device allocation that grows on demand and is released with the context.  With DDB_GUARD=1 in the environment
// every allocation is bracketed by 256-byte guard zones filled with a pattern; ddb_debug_check_guards() verifies them
// (the pool's compute-sanitizer is closed, so out-of-bounds writes are caught this way in the test suite).
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int reserve(size_t need);
    void release();
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ---------------------------------------------------------------------------------------------
// Gram stage
// ---------------------------------------------------------------------------------------------
constexpr int kMaxFactors = 8;       // maximum total degree of a library term
constexpr uint16_t kNoFactor = 0xFFFF;
constexpr uint16_t kFactorIsY = 0x8000;

// One contiguous piece of work of the Gram kernel: a block pair (bi >= bj) over a stage range.
struct GramSegment {
    int32_t bi, bj;          // block-row / block-column of the augmented library
    int32_t slot;            // partial-tile slot the accumulators are flushed to
    int32_t pad;             // bit 0: only rows 0..63 of the pair are needed (64 x 128 patch; moment schedule)
    int64_t stage_begin, stage_end;
};

struct GramPlan {
    int bt = 0;              // block tile (64 or 128)
    int ks = 0;              // samples per stage
    int nblk = 0;            // blocks per dimension of the augmented library
    int npairs = 0;          // lower-triangular block pairs
    int nctas = 0;
    int nslots = 0;
    int64_t nstages = 0;
    std::vector<GramSegment> segments;     // sorted by CTA
    std::vector<int32_t> cta_seg_begin;    // nctas + 1
    std::vector<int32_t> pair_slot_begin;  // npairs + 1  (slots of one pair are contiguous)
    std::vector<int32_t> slot_order;       // slots grouped by pair, fixed summation order
    std::vector<std::pair<int, int>> pairs; // pair index -> (block row, block column)
};

void build_gram_plan(GramPlan& plan, int kaug, int64_t m, int nctas, int bt, int ks);
void build_plan_from_pairs(GramPlan& plan, const std::vector<std::pair<int, int>>& pairs,
                           const std::vector<double>& cost, int64_t m, int nctas, int bt, int ks);

struct SweepState;
struct DmdState;
struct RingState;
struct TrsmState;
struct MomentState;
struct UserFeat;

}  // namespace ddb

struct ddb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaStream_t copy_stream = nullptr;  // host->device copies of ddb_upload_gram (overlap with the Gram kernels)
    std::vector<cudaEvent_t> chunk_ev;   // chunk c has landed
    void* stage_host = nullptr;          // page-locked bounce buffers of ddb_upload_gram for pageable callers (kStagePieces x kStageBytes)
    cudaEvent_t stage_ev[4] = {};        // bounce buffer b may be rewritten
    ddb::DevBuf d_chunk_packed;          // blocks of one chunk before they are added to the total

    // library
    int n = 0, k = 0, max_degree = 0;
    uint64_t lib_version = 0;            // bumped by every ddb_set_library_monomial
    std::vector<uint8_t> exps;           // n x k
    ddb::DevBuf d_factors;               // (k + n_t) x kMaxFactors uint16, built when n_t is known
    int factors_nt = -1;

    // feature map (ddb_set_features): the library's "states" are n features g_f(c_f * x[src_f]) of n_raw raw
    // states; uploads / binds take RAW states and a small kernel expands them once (n x m doubles, resident)
    int n_raw = 0;                       // 0 = no feature map (the states are used as they are)
    ddb::DevBuf d_feat;                  // n x {int kind, int src, double coef}
    ddb::DevBuf raw_X;                   // staging of the raw states for host uploads
    ddb::UserFeat* userfeat = nullptr;   // NVRTC-compiled feature kernel (ddb_set_features_source); overrides d_feat

    // data
    const double* dX = nullptr;
    const double* dY = nullptr;
    int n_t = 0;
    int64_t m = 0;
    ddb::DevBuf own_X, own_Y;

    // gram
    ddb::DevBuf d_packed;                // [G | R | yy]
    ddb::DevBuf d_partials;              // partial tiles
    ddb::DevBuf d_plan;                  // device copy of segments etc.
    ddb::GramPlan plan;
    int64_t plan_m = -1;
    int plan_kaug = -1;
    bool have_gram = false;
    // prefix-product builder of the single-pair Gram kernel (gram.cu: gram_tree_kernel)
    ddb::DevBuf d_tree;                  // 4 programs x tree_nsteps words
    uint64_t tree_lib_version = ~0ull;
    int tree_nt = -1, tree_nsteps = 0;   // 0 = the library is not closed under "drop the last factor"
    ddb::DevBuf d_pow_factors;           // power-table builder (gram_pow_kernel): factor codes into the table of powers
    uint64_t pow_lib_version = ~0ull;
    int pow_nt = -1, pow_nf = 0;         // 0 = not applicable (more than 3 states, degree < 3)

    // term normalisation (ddb_gram_scale_terms): the blocks in d_packed are those of Theta_j / scale_j
    ddb::DevBuf d_term_scale;            // divisor per term
    bool term_scale_on = false;
    ddb::DevBuf d_term_shift;            // affine normalisation (ddb_gram_affine_terms): theta~_j = (theta_j - shift_j) / scale_j
    bool term_shift_on = false;
    ddb::DevBuf d_cand_shift;            // per-candidate constants of the exact-rss pass under an affine normalisation
    ddb::DevBuf d_stats;                 // ddb_target_stats scratch
    // ddb_residual scratch
    ddb::DevBuf d_resid, d_resid_terms, d_resid_partials;
    // sample view (ddb_set_sample_range): dX / dY / m above describe the view; the whole data set is kept here
    bool range_on = false;
    bool view_empty = false;             // the current view holds no sample (ddb_set_sample_range(first, first))
    const double* full_X = nullptr;
    const double* full_Y = nullptr;
    int64_t full_m = 0;
    ddb::DevBuf view_X, view_Y;          // aligned staging of a view that starts on an odd double

    // implicit view (ddb_implicit_select): the sweep runs on G[idx, idx] with EVERY selected library row as a
    // target regressed on the other selected rows (ImplicitOptimizer, reference Implicit.jl:57-108)
    ddb::DevBuf imp_packed;              // [G' (kk x kk) | R' = G' | yy' = diag G']
    ddb::DevBuf imp_idx;                 // kk int32: selected rows of the installed library
    int imp_k = 0;                       // 0 = explicit mode

    // sweep / rss / select state lives in sweep.cu
    ddb::SweepState* sweep = nullptr;
    ddb::DmdState* dmd = nullptr;
    ddb::RingState* ring = nullptr;
    ddb::TrsmState* trsm = nullptr;
    ddb::MomentState* moment = nullptr;

    // multi-GPU (comm.cu): NCCL communicator of this context's rank, owned by the context
    void* comm = nullptr;                // ncclComm_t
    int nranks = 1, rank = 0;
    cudaEvent_t ev_ar[2] = {};           // brackets of the last queued all-reduce
    bool allreduce_pending = false;

    ddb_timings timings = {};
};
