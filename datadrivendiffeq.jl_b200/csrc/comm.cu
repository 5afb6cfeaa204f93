// Multi-GPU inside the C ABI: NCCL communicators owned by the contexts, the two exchange steps of the
// row-sharded path on the compute stream, and the one-call sharded solves.
//
// The path shards by rows: every GPU contracts its own contiguous sample range, then
//   1. ONE sum all-reduce of the packed normal-equation blocks [G | R | yy] (37.9 MB at k = 2145, n_t = 64),
//   2. the thresholded sweep runs replicated (deterministic, identical blocks on every rank),
//   3. ONE sum all-reduce of the candidate-rss table (a few KB; candidates are numbered canonically, see
//      canon_base_kernel in sweep.cu, so entry c means the same model on every rank),
// and every rank holds the full result.  Both collectives are `ncclAllReduce(sum, ncclDouble)` in place, queued
// on the context's compute stream behind the kernel that produces the buffer -- no host synchronisation between
// the Gram kernel, the collective and the sweep.
//
// Two ways to get a communicator, both ending in the same per-context state:
//   * one process (or thread) per GPU: ddb_comm_unique_id on one rank, the 128 bytes travel to the others by any
//     host transport (MPI, torch.distributed, a file), then ddb_comm_init_rank on every rank;
//   * one process driving N GPUs (what a Julia session does): ddb_group_create = N contexts + ncclCommInitAll;
//     ddb_group_solve_sparse cuts a host X / Y into N sample ranges and runs one host thread per GPU.
//
// NCCL is bound at run time (dlopen "libnccl.so.2") the first time a communicator is requested, so the library
// loads and the single-GPU path works on machines without NCCL; in a process that already loaded NCCL (e.g.
// through torch) the same copy is reused.
#include "common.cuh"
#include "sweep.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>

namespace ddb {

int sweep_run(ddb_ctx* ctx, const ddb_sweep_params* prm, const double* lambdas, int L, int64_t m_total);
int select_run(ddb_ctx* ctx, double* coef, double* lambda_opt, int32_t* iters, double* rss, int32_t* dof,
               uint32_t* support_path, double* coef_path, int32_t* iters_path, int32_t* flags);
const char* last_error_message();

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mutex;

static int nccl_load() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.handle) return DDB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("NCCL is not available (dlopen libnccl.so.2: %s); the multi-GPU entry points need it", dlerror());
        return DDB_NCCL_ERROR;
    }
    NcclApi a;
    a.handle = h;
#define DDB_NCCL_SYM(field, name)                                              \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));             \
    if (!a.field) {                                                            \
        set_error("NCCL symbol %s is missing from the loaded libnccl", name);   \
        return DDB_NCCL_ERROR;                                                 \
    }
    DDB_NCCL_SYM(GetVersion, "ncclGetVersion")
    DDB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    DDB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    DDB_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    DDB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    DDB_NCCL_SYM(CommAbort, "ncclCommAbort")
    DDB_NCCL_SYM(AllReduce, "ncclAllReduce")
    DDB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef DDB_NCCL_SYM
    g_nccl = a;
    return DDB_OK;
}

#define DDB_NCCL_TRY(expr)                                                                              \
    do {                                                                                                \
        ncclResult_t r__ = (expr);                                                                      \
        if (r__ != ncclSuccess) {                                                                       \
            ::ddb::set_error("%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
            return DDB_NCCL_ERROR;                                                                      \
        }                                                                                               \
    } while (0)

void comm_free(ddb_ctx* ctx) {
    if (ctx->comm && g_nccl.handle) g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
}

// in-place sum over the ranks of `count` doubles at d, queued on the compute stream (no host synchronisation)
int comm_allreduce(ddb_ctx* ctx, double* d, int64_t count) {
    if (!ctx->comm || ctx->nranks <= 1) return DDB_OK;
    if (!d || count <= 0) {
        set_error("all-reduce: null buffer or empty count");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev_ar[0], ctx->stream));
    DDB_NCCL_TRY(g_nccl.AllReduce(d, d, (size_t)count, ncclDouble, ncclSum, static_cast<ncclComm_t>(ctx->comm), ctx->stream));
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev_ar[1], ctx->stream));
    ctx->allreduce_pending = true;
    ctx->timings.allreduce_bytes += (double)count * 8.0;
    ctx->timings.allreduce_calls += 1;
    return DDB_OK;
}

// adds the duration of the last queued all-reduce to the timings (call after a stream synchronisation)
void comm_collect_time(ddb_ctx* ctx) {
    if (!ctx->allreduce_pending) return;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_ar[0], ctx->ev_ar[1]) == cudaSuccess) ctx->timings.allreduce_ms += ms;
    ctx->allreduce_pending = false;
}

// The sharded solve of ONE rank: host shard in, full result out.  Collective: every rank of the communicator
// calls it with the same library, parameters and thresholds.
int solve_sharded_rank(ddb_ctx* ctx, const double* X, const double* Y, int n_t, int64_t m_local, int64_t m_total,
                       const ddb_sweep_params* prm, const double* lambdas, int L, double* coef, double* lambda_opt,
                       int32_t* iters, double* rss, int32_t* dof, uint32_t* support_path, int32_t* flags) {
    DDB_TRY(ddb_upload_gram(ctx, X, Y, n_t, m_local, nullptr));  // copies overlapped by the Gram kernels
    DDB_TRY(comm_allreduce(ctx, ctx->d_packed.as<double>(), ddb_gram_count(ctx)));
    DDB_TRY(sweep_run(ctx, prm, lambdas, L, m_total));           // synchronises and collects the all-reduce time
    DDB_TRY(rss_run(ctx, ctx->sweep, nullptr));
    DDB_TRY(comm_allreduce(ctx, ctx->sweep->rss.as<double>(), ctx->sweep->ncand));
    return select_run(ctx, coef, lambda_opt, iters, rss, dof, support_path, nullptr, nullptr, flags);
}

}  // namespace ddb

using namespace ddb;

// One process driving several GPUs: contexts + communicators created together.
struct ddb_group {
    std::vector<ddb_ctx*> ctxs;
};

extern "C" {

int ddb_comm_unique_id(void* id128) {
    if (!id128) {
        set_error("ddb_comm_unique_id: null output");
        return DDB_INVALID_ARG;
    }
    DDB_TRY(nccl_load());
    static_assert(NCCL_UNIQUE_ID_BYTES == DDB_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId id;
    DDB_NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return DDB_OK;
}

int ddb_comm_init_rank(ddb_ctx* ctx, int nranks, int rank, const void* id128) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) {
        set_error("ddb_comm_init_rank: need a context, 0 <= rank < nranks and the 128-byte unique id");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    DDB_TRY(nccl_load());
    comm_free(ctx);
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    ncclComm_t comm = nullptr;
    DDB_NCCL_TRY(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return DDB_OK;
}

int ddb_comm_info(const ddb_ctx* ctx, int* nranks, int* rank, int* nccl_version) {
    if (!ctx) {
        set_error("ddb_comm_info: null context");
        return DDB_INVALID_ARG;
    }
    if (nranks) *nranks = ctx->comm ? ctx->nranks : 1;
    if (rank) *rank = ctx->comm ? ctx->rank : 0;
    if (nccl_version) {
        *nccl_version = 0;
        if (g_nccl.handle) g_nccl.GetVersion(nccl_version);
    }
    return DDB_OK;
}

int ddb_allreduce_sum(ddb_ctx* ctx, double* d, int64_t count) {
    if (!ctx) {
        set_error("ddb_allreduce_sum: null context");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    return comm_allreduce(ctx, d, count);
}

int ddb_allreduce_gram(ddb_ctx* ctx) {
    if (!ctx) {
        set_error("ddb_allreduce_gram: null context");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (!ctx->have_gram || !ctx->d_packed.p) {
        set_error("ddb_allreduce_gram: no normal-equation blocks in the context's own buffer (ddb_gram with dPacked = NULL)");
        return DDB_BAD_STATE;
    }
    return comm_allreduce(ctx, ctx->d_packed.as<double>(), ddb_gram_count(ctx));
}

int ddb_allreduce_rss(ddb_ctx* ctx) {
    if (!ctx) {
        set_error("ddb_allreduce_rss: null context");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    SweepState* S = ctx->sweep;
    if (!S || !S->valid || !S->have_rss || S->rss_ptr != S->rss.as<double>()) {
        set_error("ddb_allreduce_rss: run ddb_rss with dRss = NULL first (the context's own table)");
        return DDB_BAD_STATE;
    }
    return comm_allreduce(ctx, S->rss.as<double>(), S->ncand);
}

int ddb_solve_sparse_sharded(ddb_ctx* ctx, const double* X, const double* Y, int n_t, int64_t m_local, int64_t m_total,
                             const ddb_sweep_params* params, const double* lambdas, int L, double* coef, double* lambda_opt,
                             int32_t* iters, double* rss, int32_t* dof, uint32_t* support_path, int32_t* flags) {
    if (!ctx) {
        set_error("ddb_solve_sparse_sharded: null context");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (m_total < m_local || m_local <= 0) {
        set_error("ddb_solve_sparse_sharded: need 0 < m_local <= m_total (every rank owns at least one sample)");
        return DDB_INVALID_ARG;
    }
    return solve_sharded_rank(ctx, X, Y, n_t, m_local, m_total, params, lambdas, L, coef, lambda_opt, iters, rss, dof,
                              support_path, flags);
}

// ---- one process, N GPUs -------------------------------------------------------------------------------------

int ddb_group_create(int ngpus, const int* devices, ddb_group** out) {
    if (!out || ngpus < 1) {
        set_error("ddb_group_create: need ngpus >= 1 and a non-null output");
        return DDB_INVALID_ARG;
    }
    *out = nullptr;
    std::vector<int> devs(ngpus);
    for (int i = 0; i < ngpus; ++i) devs[i] = devices ? devices[i] : i;
    for (int i = 0; i < ngpus; ++i)
        for (int j = 0; j < i; ++j)
            if (devs[i] == devs[j]) {
                set_error("ddb_group_create: device %d is listed twice (one rank per GPU)", devs[i]);
                return DDB_INVALID_ARG;
            }
    ddb_group* g = new ddb_group();
    int status = DDB_OK;
    for (int i = 0; i < ngpus && status == DDB_OK; ++i) {
        ddb_ctx* c = nullptr;
        status = ddb_ctx_create(devs[i], &c);
        if (status == DDB_OK) g->ctxs.push_back(c);
    }
    if (status == DDB_OK && ngpus > 1) {
        status = nccl_load();
        if (status == DDB_OK) {
            std::vector<ncclComm_t> comms(ngpus, nullptr);
            ncclResult_t r = g_nccl.CommInitAll(comms.data(), ngpus, devs.data());
            if (r != ncclSuccess) {
                set_error("ncclCommInitAll(%d GPUs) failed: %s", ngpus, g_nccl.GetErrorString(r));
                status = DDB_NCCL_ERROR;
            } else {
                for (int i = 0; i < ngpus; ++i) {
                    g->ctxs[i]->comm = comms[i];
                    g->ctxs[i]->nranks = ngpus;
                    g->ctxs[i]->rank = i;
                }
            }
        }
    }
    if (status != DDB_OK) {
        for (ddb_ctx* c : g->ctxs) ddb_ctx_destroy(c);
        delete g;
        return status;
    }
    *out = g;
    return DDB_OK;
}

int ddb_group_destroy(ddb_group* g) {
    if (!g) return DDB_OK;
    for (ddb_ctx* c : g->ctxs) ddb_ctx_destroy(c);
    delete g;
    return DDB_OK;
}

int ddb_group_size(const ddb_group* g) { return g ? (int)g->ctxs.size() : 0; }

ddb_ctx* ddb_group_ctx(ddb_group* g, int i) {
    if (!g || i < 0 || i >= (int)g->ctxs.size()) return nullptr;
    return g->ctxs[i];
}

int ddb_group_set_library_monomial(ddb_group* g, int n, int k, const uint8_t* exps) {
    if (!g) {
        set_error("ddb_group_set_library_monomial: null group");
        return DDB_INVALID_ARG;
    }
    for (ddb_ctx* c : g->ctxs) DDB_TRY(ddb_set_library_monomial(c, n, k, exps));
    return DDB_OK;
}

int ddb_group_set_features(ddb_group* g, int n_raw, int n_feat, const int32_t* kind, const int32_t* source, const double* coef) {
    if (!g) {
        set_error("ddb_group_set_features: null group");
        return DDB_INVALID_ARG;
    }
    for (ddb_ctx* c : g->ctxs) DDB_TRY(ddb_set_features(c, n_raw, n_feat, kind, source, coef));
    return DDB_OK;
}

int ddb_group_shard(const ddb_group* g, int64_t m, int rank, int64_t* first, int64_t* last) {
    const int N = ddb_group_size(g);
    if (N <= 0 || rank < 0 || rank >= N || m <= 0 || !first || !last) {
        set_error("ddb_group_shard: bad group / rank / sample count");
        return DDB_INVALID_ARG;
    }
    // contiguous ranges whose interior boundaries are multiples of 16 samples (the Gram kernel's stage)
    const int64_t blocks = (m + 15) / 16;
    *first = std::min<int64_t>(m, blocks * rank / N * 16);
    *last = std::min<int64_t>(m, blocks * (rank + 1) / N * 16);
    return DDB_OK;
}

int ddb_group_solve_sparse(ddb_group* g, const double* X, const double* Y, int n_t, int64_t m,
                           const ddb_sweep_params* params, const double* lambdas, int L, double* coef, double* lambda_opt,
                           int32_t* iters, double* rss, int32_t* dof, uint32_t* support_path, int32_t* flags) {
    if (!g || g->ctxs.empty()) {
        set_error("ddb_group_solve_sparse: null group");
        return DDB_INVALID_ARG;
    }
    const int N = (int)g->ctxs.size();
    if (!X || !Y || n_t <= 0 || m <= 0) {
        set_error("ddb_group_solve_sparse: X, Y must be non-null and n_t, m positive");
        return DDB_INVALID_ARG;
    }
    if (m < (int64_t)16 * N) {
        set_error("ddb_group_solve_sparse: %lld samples cannot be cut into %d ranges of whole 16-sample stages; use fewer GPUs",
                  (long long)m, N);
        return DDB_INVALID_ARG;
    }
    ddb_ctx* c0 = g->ctxs[0];
    const int n_in = c0->n_raw > 0 ? c0->n_raw : c0->n;
    if (n_in <= 0) {
        set_error("ddb_group_solve_sparse: install the library first");
        return DDB_BAD_STATE;
    }
    std::vector<int> status(N, DDB_OK);
    std::vector<std::string> message(N);
    auto work = [&](int r) {
        int64_t lo = 0, hi = 0;
        ddb_group_shard(g, m, r, &lo, &hi);
        const bool root = r == 0;
        status[r] = solve_sharded_rank(g->ctxs[r], X + lo * n_in, Y + lo * n_t, n_t, hi - lo, m, params, lambdas, L,
                                       root ? coef : nullptr, root ? lambda_opt : nullptr, root ? iters : nullptr,
                                       root ? rss : nullptr, root ? dof : nullptr, root ? support_path : nullptr,
                                       root ? flags : nullptr);
        if (status[r] != DDB_OK) message[r] = last_error_message();
    };
    if (N == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int r = 0; r < N; ++r)
            th.emplace_back([&, r] {
                cudaSetDevice(g->ctxs[r]->device);
                work(r);
            });
        for (auto& t : th) t.join();
    }
    for (int r = 0; r < N; ++r)
        if (status[r] != DDB_OK) {
            set_error("rank %d (device %d): %s", r, g->ctxs[r]->device, message[r].c_str());
            return status[r];
        }
    return DDB_OK;
}

int ddb_group_target_stats(ddb_group* g, double* mean, double* css) {
    if (!g || g->ctxs.empty() || !mean || !css) {
        set_error("ddb_group_target_stats: null argument");
        return DDB_INVALID_ARG;
    }
    const int N = (int)g->ctxs.size();
    int64_t m_total = 0;
    for (ddb_ctx* c : g->ctxs) m_total += c->m;
    const int n_t = g->ctxs[0]->n_t;
    std::vector<int> status(N, DDB_OK);
    std::vector<std::string> message(N);
    std::vector<std::vector<double>> mn(N, std::vector<double>(std::max(1, n_t))), cs(N, std::vector<double>(std::max(1, n_t)));
    auto work = [&](int r) {
        cudaSetDevice(g->ctxs[r]->device);
        status[r] = ddb_target_stats(g->ctxs[r], m_total, mn[r].data(), cs[r].data());
        if (status[r] != DDB_OK) message[r] = last_error_message();
    };
    std::vector<std::thread> th;
    for (int r = 1; r < N; ++r) th.emplace_back(work, r);
    work(0);
    for (auto& t : th) t.join();
    for (int r = 0; r < N; ++r)
        if (status[r] != DDB_OK) {
            set_error("rank %d (device %d): %s", r, g->ctxs[r]->device, message[r].c_str());
            return status[r];
        }
    for (int e = 0; e < n_t; ++e) mean[e] = mn[0][e], css[e] = cs[0][e];
    return DDB_OK;
}

int ddb_group_get_timings(const ddb_group* g, ddb_timings* out) {
    if (!g || g->ctxs.empty() || !out) {
        set_error("ddb_group_get_timings: null argument");
        return DDB_INVALID_ARG;
    }
    *out = g->ctxs[0]->timings;
    for (const ddb_ctx* c : g->ctxs) {  // the slowest rank of every stage
        out->h2d_ms = std::max(out->h2d_ms, c->timings.h2d_ms);
        out->gram_ms = std::max(out->gram_ms, c->timings.gram_ms);
        out->rss_ms = std::max(out->rss_ms, c->timings.rss_ms);
        out->allreduce_ms = std::max(out->allreduce_ms, c->timings.allreduce_ms);
    }
    return DDB_OK;
}

}  // extern "C"
