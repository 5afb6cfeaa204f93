// extern "C" entry points of libddb200.so (declared in include/ddb.h): context, library, data and
// the stage wrappers.  Kernels live in gram.cu / sweep.cu / rss.cu / generate.cu / dmd.cu.
#include "common.cuh"
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

namespace ddb {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// ---- guard zones (DDB_GUARD=1): see DevBuf in common.cuh ----------------------------------------------------------
constexpr size_t kGuard = 256;
constexpr int kGuardByte = 0xA5;
static bool guard_on() {
    static const bool on = [] {
        const char* e = getenv("DDB_GUARD");
        return e && atoi(e) != 0;
    }();
    return on;
}
static std::mutex g_guard_mutex;
static std::map<void*, size_t> g_guard_live;  // user pointer -> user bytes

int DevBuf::reserve(size_t need) {
    if (need <= bytes && p) return DDB_OK;
    release();
    if (need == 0) need = 16;
    const size_t pad = guard_on() ? kGuard : 0;
    void* raw = nullptr;
    cudaError_t e = cudaMalloc(&raw, need + 2 * pad);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", need, cudaGetErrorString(e));
        p = nullptr;
        return e == cudaErrorMemoryAllocation ? DDB_OUT_OF_MEMORY : DDB_CUDA_ERROR;
    }
    if (pad) {
        cudaMemset(raw, kGuardByte, pad);
        cudaMemset(static_cast<char*>(raw) + pad + need, kGuardByte, pad);
        std::lock_guard<std::mutex> lk(g_guard_mutex);
        g_guard_live[static_cast<char*>(raw) + pad] = need;
    }
    p = static_cast<char*>(raw) + pad;
    bytes = need;
    return DDB_OK;
}
void DevBuf::release() {
    if (p) {
        if (guard_on()) {
            {
                std::lock_guard<std::mutex> lk(g_guard_mutex);
                g_guard_live.erase(p);
            }
            cudaFree(static_cast<char*>(p) - kGuard);
        } else {
            cudaFree(p);
        }
    }
    p = nullptr;
    bytes = 0;
}

// number of live allocations whose guard zones were overwritten (0 without DDB_GUARD); the first one is described
// in the error message
static int check_guards() {
    if (!guard_on()) return 0;
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_guard_mutex);
    int bad = 0;
    std::vector<unsigned char> h(2 * kGuard);
    for (const auto& kv : g_guard_live) {
        char* user = static_cast<char*>(kv.first);
        if (cudaMemcpy(h.data(), user - kGuard, kGuard, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(h.data() + kGuard, user + kv.second, kGuard, cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();
            continue;  // an allocation of another device's context: checked when that device is current
        }
        for (size_t i = 0; i < 2 * kGuard; ++i)
            if (h[i] != kGuardByte) {
                if (!bad)
                    set_error("guard zone %s a %zu-byte allocation was overwritten (offset %zu)", i < kGuard ? "before" : "after",
                              kv.second, i < kGuard ? i : i - kGuard);
                ++bad;
                break;
            }
    }
    return bad;
}

int gram_run(ddb_ctx* ctx, double* dPacked);
int residual_run(ddb_ctx* ctx, const double* coef, const double* offset, double* rss_out);
int gram_ensure_factors(ddb_ctx* ctx, int kaug_pad);
int generate_run(ddb_ctx* ctx, int system_id, int n_states, int64_t m, int64_t first_sample, uint64_t seed,
                 double noise_std);
void sweep_free(ddb_ctx* ctx);
void dmd_free(ddb_ctx* ctx);
void ring_free(ddb_ctx* ctx);
void moment_free(ddb_ctx* ctx);
int64_t moment_wave_describe(const uint8_t* exps, int n, int k, int n_t, int64_t m, int nctas, int64_t* segments,
                             int64_t cap, double* info);
int64_t moment_plan_describe(const uint8_t* exps, int n, int k, int n_t, int32_t* src, int64_t src_cap, uint16_t* tab,
                             int64_t tab_cap, int32_t* pairs, int64_t pairs_cap, int64_t* info);
int64_t moment_builder_describe(const uint8_t* exps, int n, int k, int n_t, uint32_t* optab, int64_t optab_cap,
                                uint8_t* opcnt, int64_t opcnt_cap, int32_t* block_ops, int64_t block_cap);
void trsm_free(ddb_ctx* ctx);
void comm_free(ddb_ctx* ctx);
int gram_pow_factors(const uint8_t* exps, int n, int k, int n_t, int deg, std::vector<uint16_t>* tab);  // gram.cu
void userfeat_free(ddb_ctx* ctx);
int userfeat_nfeat(const ddb_ctx* ctx);
int userfeat_expand(ddb_ctx* ctx, const double* d_raw, int64_t mc, double* out, cudaStream_t st);
const char* last_error_message() { return g_error; }

}  // namespace ddb

using namespace ddb;

#define CHECK_CTX(ctx)                                     \
    do {                                                   \
        if (!(ctx)) {                                      \
            set_error("%s: null context", __func__);       \
            return DDB_INVALID_ARG;                        \
        }                                                  \
        DDB_CUDA_TRY(cudaSetDevice((ctx)->device));        \
    } while (0)

extern "C" {

const char* ddb_last_error(void) { return g_error; }
const char* ddb_version(void) { return "ddb200 0.2 (sm_100a)"; }

int ddb_debug_check_guards(void) { return check_guards(); }

int ddb_ctx_create(int device, ddb_ctx** out) {
    if (!out) {
        set_error("ddb_ctx_create: out is null");
        return DDB_INVALID_ARG;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("ddb_ctx_create: no CUDA device available (%s); this backend has no CPU fallback",
                  cudaGetErrorString(e));
        return DDB_CUDA_ERROR;
    }
    if (device < 0 || device >= count) {
        set_error("ddb_ctx_create: device %d out of range (have %d)", device, count);
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    DDB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("ddb_ctx_create: device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major,
                  prop.minor);
        return DDB_UNSUPPORTED;
    }
    ddb_ctx* ctx = new ddb_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    DDB_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (auto& ev : ctx->ev) DDB_CUDA_TRY(cudaEventCreate(&ev));
    for (auto& ev : ctx->ev_ar) DDB_CUDA_TRY(cudaEventCreate(&ev));
    *out = ctx;
    return DDB_OK;
}

int ddb_ctx_destroy(ddb_ctx* ctx) {
    if (!ctx) return DDB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    comm_free(ctx);
    userfeat_free(ctx);
    sweep_free(ctx);
    dmd_free(ctx);
    ring_free(ctx);
    moment_free(ctx);
    trsm_free(ctx);
    ctx->d_factors.release();
    ctx->own_X.release();
    ctx->own_Y.release();
    ctx->d_packed.release();
    ctx->imp_packed.release();
    ctx->imp_idx.release();
    ctx->d_partials.release();
    ctx->d_plan.release();
    ctx->d_tree.release();
    ctx->d_pow_factors.release();
    ctx->d_chunk_packed.release();
    ctx->d_term_scale.release();
    ctx->d_term_shift.release();
    ctx->d_cand_shift.release();
    ctx->d_resid.release();
    ctx->d_stats.release();
    ctx->d_resid_terms.release();
    ctx->d_resid_partials.release();
    ctx->d_feat.release();
    ctx->raw_X.release();
    ctx->view_X.release();
    ctx->view_Y.release();
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_ar)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->chunk_ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->stage_ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return DDB_OK;
}

int ddb_set_library_monomial(ddb_ctx* ctx, int n, int k, const uint8_t* exps) {
    CHECK_CTX(ctx);
    if (n <= 0 || k <= 0 || !exps) {
        set_error("ddb_set_library_monomial: n and k must be positive and exps non-null");
        return DDB_INVALID_ARG;
    }
    int maxdeg = 0;
    for (int j = 0; j < k; ++j) {
        int d = 0;
        for (int i = 0; i < n; ++i) d += exps[(size_t)i + (size_t)n * j];
        if (d > maxdeg) maxdeg = d;
    }
    if (maxdeg > kMaxFactors) {
        set_error("ddb_set_library_monomial: total degree %d exceeds the supported maximum %d", maxdeg, kMaxFactors);
        return DDB_UNSUPPORTED;
    }
    ctx->n = n;
    ctx->k = k;
    ctx->max_degree = maxdeg;
    ctx->exps.assign(exps, exps + (size_t)n * k);
    ctx->lib_version++;
    ctx->factors_nt = -1;
    ctx->plan_m = -1;
    ctx->have_gram = false;
    ctx->imp_k = 0;
    ctx->term_scale_on = false;
    return DDB_OK;
}

// ---------------------------------------------------------------------------------------------
// feature map: phi_f(x) = g_f(c_f * x[src_f])
// ---------------------------------------------------------------------------------------------
struct FeatDesc {
    int32_t kind, src;
    double coef;
};

__global__ void feature_kernel(const double* __restrict__ raw, int n_raw, const FeatDesc* __restrict__ feat, int n,
                               int64_t m, double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * n) return;
    const int64_t s = t / n;
    const int f = (int)(t % n);
    const FeatDesc d = feat[f];
    const double x = raw[s * n_raw + d.src];
    double v;
    switch (d.kind) {
        case DDB_FEAT_SIN: v = sin(d.coef * x); break;
        case DDB_FEAT_COS: v = cos(d.coef * x); break;
        case DDB_FEAT_EXP: v = exp(d.coef * x); break;
        case DDB_FEAT_CHEBYSHEV: v = cos(d.coef * acos(x)); break;
        default: v = d.coef * x; break;
    }
    out[t] = v;
}

// expands raw states [s0, s0 + mc) (device pointer to the first of them) into ctx->own_X rows [s0, ...)
static int expand_features(ddb_ctx* ctx, const double* d_raw, int64_t s0, int64_t mc, cudaStream_t st) {
    if (ctx->userfeat) return userfeat_expand(ctx, d_raw, mc, ctx->own_X.as<double>() + s0 * ctx->n, st);
    const int64_t tot = mc * ctx->n;
    feature_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_raw, ctx->n_raw, ctx->d_feat.as<FeatDesc>(), ctx->n, mc,
                                                                  ctx->own_X.as<double>() + s0 * ctx->n);
    DDB_CUDA_TRY(cudaGetLastError());
    return DDB_OK;
}

int ddb_set_features(ddb_ctx* ctx, int n_raw, int n_feat, const int32_t* kind, const int32_t* source, const double* coef) {
    CHECK_CTX(ctx);
    userfeat_free(ctx);  // the fixed feature kinds replace a compiled feature kernel
    if (n_feat == 0) {
        ctx->n_raw = 0;
        ctx->have_gram = false;
        ctx->imp_k = 0;
        return DDB_OK;
    }
    if (n_raw <= 0 || n_feat < 0 || !kind || !source || !coef) {
        set_error("ddb_set_features: n_raw must be positive and kind, source, coef non-null");
        return DDB_INVALID_ARG;
    }
    std::vector<FeatDesc> tab(n_feat);
    for (int f = 0; f < n_feat; ++f) {
        if (kind[f] < DDB_FEAT_IDENTITY || kind[f] > DDB_FEAT_CHEBYSHEV || source[f] < 0 || source[f] >= n_raw ||
            !std::isfinite(coef[f])) {
            set_error("ddb_set_features: feature %d has an invalid kind / source / coefficient", f);
            return DDB_INVALID_ARG;
        }
        tab[f] = {kind[f], source[f], coef[f]};
    }
    DDB_TRY(ctx->d_feat.reserve(tab.size() * sizeof(FeatDesc)));
    DDB_CUDA_TRY(cudaMemcpyAsync(ctx->d_feat.p, tab.data(), tab.size() * sizeof(FeatDesc), cudaMemcpyHostToDevice, ctx->stream));
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->n_raw = n_raw;
    ctx->have_gram = false;
    ctx->imp_k = 0;
    ctx->term_scale_on = false;
    // the library must be (re)installed over n_feat states; data must be (re)bound
    ctx->dX = ctx->dY = nullptr;
    ctx->range_on = false;
    ctx->view_empty = false;
    ctx->m = 0;
    return DDB_OK;
}

static int check_data_args(ddb_ctx* ctx, const void* X, const void* Y, int n_t, int64_t m, const char* who) {
    if (ctx->n <= 0) {
        set_error("%s: install the library first (the number of states comes from it)", who);
        return DDB_BAD_STATE;
    }
    if (!X || !Y || n_t <= 0 || m <= 0) {
        set_error("%s: X, Y must be non-null and n_t, m positive", who);
        return DDB_INVALID_ARG;
    }
    if (ctx->n_raw > 0 && ctx->userfeat && userfeat_nfeat(ctx) != ctx->n) {
        set_error("%s: the installed library has %d states but the compiled feature kernel writes %d features", who, ctx->n,
                  userfeat_nfeat(ctx));
        return DDB_BAD_STATE;
    }
    if (ctx->n_raw > 0 && !ctx->userfeat && ctx->d_feat.bytes < (size_t)ctx->n * sizeof(FeatDesc)) {
        set_error("%s: the installed library has %d states but the feature map defines fewer features", who, ctx->n);
        return DDB_BAD_STATE;
    }
    return DDB_OK;
}

int ddb_upload(ddb_ctx* ctx, const double* X, const double* Y, int n_t, int64_t m) {
    CHECK_CTX(ctx);
    DDB_TRY(check_data_args(ctx, X, Y, n_t, m, "ddb_upload"));
    const size_t xb = (size_t)ctx->n * m * sizeof(double), yb = (size_t)n_t * m * sizeof(double);
    DDB_TRY(ctx->own_X.reserve(xb));
    DDB_TRY(ctx->own_Y.reserve(yb));
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (ctx->n_raw > 0) {  // raw states -> staging -> features
        const size_t rb = (size_t)ctx->n_raw * m * sizeof(double);
        DDB_TRY(ctx->raw_X.reserve(rb));
        DDB_CUDA_TRY(cudaMemcpyAsync(ctx->raw_X.p, X, rb, cudaMemcpyHostToDevice, ctx->stream));
        DDB_TRY(expand_features(ctx, ctx->raw_X.as<double>(), 0, m, ctx->stream));
    } else {
        DDB_CUDA_TRY(cudaMemcpyAsync(ctx->own_X.p, X, xb, cudaMemcpyHostToDevice, ctx->stream));
    }
    DDB_CUDA_TRY(cudaMemcpyAsync(ctx->own_Y.p, Y, yb, cudaMemcpyHostToDevice, ctx->stream));
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[3], ctx->stream));
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
    ctx->timings.h2d_ms = ms;
    ctx->dX = ctx->own_X.as<double>();
    ctx->dY = ctx->own_Y.as<double>();
    ctx->range_on = false;
    ctx->view_empty = false;
    ctx->n_t = n_t;
    ctx->m = m;
    ctx->have_gram = false;
    ctx->imp_k = 0;
    ctx->term_scale_on = false;
    return DDB_OK;
}

__global__ void packed_add_kernel(double* __restrict__ acc, const double* __restrict__ part, int64_t count, int first) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) acc[i] = first ? part[i] : acc[i] + part[i];
}

// Pageable host memory -> device through our own page-locked bounce buffers.  cudaMemcpyAsync on pageable memory
// stages through the driver's buffers with ONE host thread (~10 GB/s on these boxes: a 10 GB upload then outlasts the
// 0.7 s Gram stage it should hide behind -- e2e 10.1e6 samples/s against 13.9e6 from pinned memory); here a small
// thread pool fills a 32 MB piece of a 3-deep ring while the DMA engine drains the previous ones.
constexpr size_t kStageBytes = 32u << 20;
constexpr int kStagePieces = 3;
struct StagePool {
    int nthreads;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    int gen = 0, done = 0;
    bool stop = false;
    char* dst = nullptr;
    const char* src = nullptr;
    size_t len = 0;
    explicit StagePool(int n) : nthreads(n) {
        for (int t = 1; t < nthreads; ++t)
            workers.emplace_back([this, t] {
                int seen = 0;
                for (;;) {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_go.wait(lk, [&] { return stop || gen != seen; });
                    if (stop) return;
                    seen = gen;
                    char* d = dst;
                    const char* s = src;
                    const size_t n = len;
                    lk.unlock();
                    slice(d, s, n, t);
                    lk.lock();
                    ++done;
                    cv_done.notify_one();
                }
            });
    }
    void slice(char* d, const char* s, size_t n, int t) const {
        const size_t per = ((n + nthreads - 1) / nthreads + 63) & ~size_t(63);
        const size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
        if (hi > lo) memcpy(d + lo, s + lo, hi - lo);
    }
    void copy(char* d, const char* s, size_t n) {
        {
            std::lock_guard<std::mutex> lk(mu);
            dst = d, src = s, len = n, done = 0;
            ++gen;
        }
        cv_go.notify_all();
        slice(d, s, n, 0);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return done == nthreads - 1; });
    }
    ~StagePool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_go.notify_all();
        for (auto& w : workers) w.join();
    }
};
// Measured on a 16-core box (tools/pageable_probe.py, C3 end to end from pageable arrays, 718 ms from pinned ones):
// 1 thread 786 ms, 2: 747, 4: 738, 8: 727, 16: 723.  Half the cores divided by the ranks that share the node.
static int stage_threads(const ddb_ctx* ctx) {
    if (const char* e = getenv("DDB_STAGE_THREADS")) return std::max(1, std::min(16, atoi(e)));
    const unsigned hc = std::max(2u, std::thread::hardware_concurrency());
    const unsigned ranks = (unsigned)std::max(1, ctx->nranks);
    return (int)std::max(2u, std::min(8u, hc / (2 * ranks)));
}
// host (pageable) -> device, `bytes` bytes, on ctx->copy_stream through the bounce ring; *piece counts ring slots used
static cudaError_t staged_h2d(ddb_ctx* ctx, StagePool& pool, void* dst, const void* src, size_t bytes, int* piece) {
    for (size_t off = 0; off < bytes; off += kStageBytes) {
        const size_t len = std::min(kStageBytes, bytes - off);
        const int b = (*piece)++ % kStagePieces;
        cudaError_t e = cudaEventSynchronize(ctx->stage_ev[b]);  // the DMA that last read this slot has finished
        if (e != cudaSuccess) return e;
        char* slot = static_cast<char*>(ctx->stage_host) + (size_t)b * kStageBytes;
        pool.copy(slot, static_cast<const char*>(src) + off, len);
        e = cudaMemcpyAsync(static_cast<char*>(dst) + off, slot, len, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->stage_ev[b], ctx->copy_stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Host buffers -> device (kept resident, as ddb_upload) with the copies overlapped by the Gram stage: the
// samples are cut into chunks, chunk c + 1 travels on a copy stream while the Gram kernels run on chunk c;
// the per-chunk blocks are added in chunk order (deterministic).  With pinned host memory the transfer
// disappears behind the tensor work except for the first chunk.
int ddb_upload_gram(ddb_ctx* ctx, const double* X, const double* Y, int n_t, int64_t m, double* dPacked) {
    CHECK_CTX(ctx);
    DDB_TRY(check_data_args(ctx, X, Y, n_t, m, "ddb_upload_gram"));
    ctx->timings.allreduce_ms = 0.0;
    ctx->timings.allreduce_bytes = 0.0;
    ctx->timings.allreduce_calls = 0;
    if (ctx->k <= 0) {
        set_error("ddb_upload_gram: no library installed");
        return DDB_BAD_STATE;
    }
    const int n = ctx->n;
    const int n_in = ctx->n_raw > 0 ? ctx->n_raw : n;  // states per sample in the caller's buffer
    const size_t row_bytes = (size_t)(n_in + n_t) * sizeof(double);
    // At most 16 chunks of at least ~64 MB, chunk boundaries on multiples of 64 samples.  Only the first chunk's copy
    // is exposed, so large uploads start with two short chunks (1/128 and 1/32 of the samples: the copy of the second
    // takes about as long as the Gram kernels of the first, the host link being ~4 x faster than the Gram stage
    // consumes samples) and split the rest evenly; the plans of the four chunk lengths are cached.
    std::vector<int64_t> starts;  // chunk c = samples [starts[c], starts[c + 1])
    {
        const double total_mb = (double)m * row_bytes / (1024.0 * 1024.0);
        auto round64 = [](int64_t v) { return std::max<int64_t>(64, (v + 63) / 64 * 64); };
        starts.push_back(0);
        int64_t rest_begin = 0;
        int even = (int)std::min<int64_t>(16, std::max<int64_t>(1, (int64_t)(total_mb / 64.0)));
        const char* front_env = getenv("DDB_UPLOAD_FRONT_MB");  // uploads from this size on start with the short chunks
        if (total_mb >= (front_env ? atof(front_env) : 2048.0) && m >= 128 * 64) {
            starts.push_back(round64(m / 128));
            starts.push_back(starts.back() + round64(m / 32));
            rest_begin = starts.back();
            even = 14;
        }
        const int64_t each = round64((m - rest_begin + even - 1) / even);
        for (int64_t s0 = rest_begin + each; s0 < m; s0 += each) starts.push_back(s0);
        starts.push_back(m);
    }
    const int nchunks = (int)starts.size() - 1;
    if (nchunks <= 1) {
        DDB_TRY(ddb_upload(ctx, X, Y, n_t, m));
        return gram_run(ctx, dPacked);
    }
    const size_t xb = (size_t)n * m * sizeof(double), yb = (size_t)n_t * m * sizeof(double);
    DDB_TRY(ctx->own_X.reserve(xb));
    DDB_TRY(ctx->own_Y.reserve(yb));
    if (!ctx->copy_stream) DDB_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    while ((int)ctx->chunk_ev.size() < nchunks + 1) {
        cudaEvent_t e;
        DDB_CUDA_TRY(cudaEventCreate(&e));
        ctx->chunk_ev.push_back(e);
    }
    ctx->n_t = n_t;
    ctx->range_on = false;
    ctx->view_empty = false;
    ctx->have_gram = false;
    ctx->imp_k = 0;
    ctx->term_scale_on = false;
    const int64_t count = (int64_t)ctx->k * ctx->k + (int64_t)ctx->k * n_t + n_t;
    if (!dPacked) {
        DDB_TRY(ctx->d_packed.reserve((size_t)count * sizeof(double)));
        dPacked = ctx->d_packed.as<double>();
    }
    DDB_TRY(ctx->d_chunk_packed.reserve((size_t)count * sizeof(double)));
    double* dX = ctx->own_X.as<double>();
    double* dY = ctx->own_Y.as<double>();
    double* dIn = dX;  // where the copies land: the state buffer itself, or the raw staging of a feature map
    if (ctx->n_raw > 0) {
        DDB_TRY(ctx->raw_X.reserve((size_t)n_in * m * sizeof(double)));
        dIn = ctx->raw_X.as<double>();
    }
    // Page-locked host buffers: all copies are queued up front on the copy stream (event c fires when chunk c is
    // resident) and the DMA engine runs ahead of the kernels.  PAGEABLE buffers (a Julia `Matrix`, a NumPy array) go
    // through the library's own bounce ring (staged_h2d), filled by a helper thread and its pool -- the calling thread
    // keeps launching the Gram kernels of the chunks that have landed, and the transfer still hides behind the tensor
    // work.
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
    };
    const bool pinned = is_pinned(X) && is_pinned(Y);
    std::mutex mu;
    std::condition_variable cv;
    int landed = 0;                 // chunks whose copies have been issued and whose event is recorded
    cudaError_t copy_err = cudaSuccess;
    if (!pinned && !ctx->stage_host) {
        DDB_CUDA_TRY(cudaHostAlloc(&ctx->stage_host, kStagePieces * kStageBytes, cudaHostAllocDefault));
        for (int b = 0; b < kStagePieces; ++b) DDB_CUDA_TRY(cudaEventCreateWithFlags(&ctx->stage_ev[b], cudaEventDisableTiming));
    }
    auto issue_copies = [&](bool notify) {
        std::unique_ptr<StagePool> pool;
        int piece = 0;
        if (notify) pool.reset(new StagePool(stage_threads(ctx)));  // pageable caller: this is the helper thread
        for (int c = 0; c < nchunks; ++c) {
            const int64_t s0 = starts[c], mc = starts[c + 1] - s0;
            cudaError_t e;
            if (pool) {
                e = staged_h2d(ctx, *pool, dIn + s0 * n_in, X + s0 * n_in, (size_t)mc * n_in * sizeof(double), &piece);
                if (e == cudaSuccess) e = staged_h2d(ctx, *pool, dY + s0 * n_t, Y + s0 * n_t, (size_t)mc * n_t * sizeof(double), &piece);
            } else {
                e = cudaMemcpyAsync(dIn + s0 * n_in, X + s0 * n_in, (size_t)mc * n_in * sizeof(double), cudaMemcpyHostToDevice,
                                    ctx->copy_stream);
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(dY + s0 * n_t, Y + s0 * n_t, (size_t)mc * n_t * sizeof(double), cudaMemcpyHostToDevice,
                                        ctx->copy_stream);
            }
            if (e == cudaSuccess) e = cudaEventRecord(ctx->chunk_ev[c], ctx->copy_stream);
            if (notify) {
                std::lock_guard<std::mutex> lk(mu);
                if (e != cudaSuccess) copy_err = e;
                landed = (e == cudaSuccess) ? c + 1 : nchunks;  // on failure release the consumer; it checks copy_err
                cv.notify_all();
            } else if (e != cudaSuccess) {
                copy_err = e;
            }
            if (e != cudaSuccess) return;
        }
    };
    DDB_CUDA_TRY(cudaEventRecord(ctx->chunk_ev[nchunks], ctx->copy_stream));
    std::thread copier;
    if (pinned) {
        issue_copies(false);
    } else {
        const int device = ctx->device;
        copier = std::thread([&, device] {
            cudaSetDevice(device);
            issue_copies(true);
        });
    }
    double gram_ms = 0.0;
    int launches = 0, status = DDB_OK;
    cudaError_t cerr = pinned ? copy_err : cudaSuccess;
    for (int c = 0; c < nchunks && status == DDB_OK && cerr == cudaSuccess; ++c) {
        const int64_t s0 = starts[c], mc = starts[c + 1] - s0;
        if (!pinned) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return landed > c; });
            cerr = copy_err;
            if (cerr != cudaSuccess) break;
        }
        cerr = cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[c], 0);
        if (cerr != cudaSuccess) break;
        if (ctx->n_raw > 0) status = expand_features(ctx, dIn + s0 * n_in, s0, mc, ctx->stream);
        if (status != DDB_OK) break;
        ctx->dX = dX + s0 * n;
        ctx->dY = dY + s0 * n_t;
        ctx->m = mc;
        status = gram_run(ctx, ctx->d_chunk_packed.as<double>());
        if (status != DDB_OK) break;
        gram_ms += ctx->timings.gram_ms;
        launches += ctx->timings.gram_launches + 1;
        packed_add_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(dPacked, ctx->d_chunk_packed.as<double>(),
                                                                                    count, c == 0);
    }
    // every exit path: the helper thread is joined, both streams are drained (the caller may free X / Y as soon as
    // we return) and the context describes the whole data set again
    if (copier.joinable()) copier.join();
    ctx->dX = dX;
    ctx->dY = dY;
    ctx->m = m;
    const cudaError_t ce = cudaStreamSynchronize(ctx->copy_stream);
    const cudaError_t se = cudaStreamSynchronize(ctx->stream);
    if (status != DDB_OK) return status;
    DDB_CUDA_TRY(cerr);
    DDB_CUDA_TRY(copy_err);
    DDB_CUDA_TRY(ce);
    DDB_CUDA_TRY(se);
    DDB_CUDA_TRY(cudaGetLastError());
    float h2d = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&h2d, ctx->chunk_ev[nchunks], ctx->chunk_ev[nchunks - 1]));
    ctx->timings.h2d_ms = h2d;
    ctx->timings.gram_ms = gram_ms;
    ctx->timings.gram_launches = launches;
    ctx->have_gram = (dPacked == ctx->d_packed.as<double>());
    return DDB_OK;
}

// ---------------------------------------------------------------------------------------------
// sample views, permutation, term normalisation, residuals (pre-/post-processing options of the reference)
// ---------------------------------------------------------------------------------------------
int ddb_set_sample_range(ddb_ctx* ctx, int64_t first, int64_t last) {
    CHECK_CTX(ctx);
    if (!ctx->range_on) {
        ctx->full_X = ctx->dX;
        ctx->full_Y = ctx->dY;
        ctx->full_m = ctx->m;
    }
    if (!ctx->full_X || ctx->full_m <= 0) {
        set_error("ddb_set_sample_range: no data bound");
        return DDB_BAD_STATE;
    }
    if (last < 0) last = ctx->full_m;
    // first == last: an EMPTY view (a rank of a sharded run that holds no sample of a batch or of the test part): the
    // Gram / rss / residual / statistics passes then contribute zeros to the sums over the ranks
    if (first < 0 || first > last || last > ctx->full_m) {
        set_error("ddb_set_sample_range: need 0 <= first <= last <= m");
        return DDB_INVALID_ARG;
    }
    ctx->range_on = !(first == 0 && last == ctx->full_m);
    ctx->view_empty = (first == last);
    ctx->dX = ctx->full_X + first * ctx->n;
    ctx->dY = ctx->full_Y + first * ctx->n_t;
    ctx->m = last - first;
    // The Gram kernels fetch sample tiles with 16-byte bulk copies: a view that starts on an odd double (odd
    // first with an odd number of states or targets, e.g. Lorenz with a batch size of 125) is staged once into an
    // aligned buffer of its own -- the reference accepts any batch size and split.
    if (!ctx->view_empty && ((((uintptr_t)ctx->dX) & 15) || (((uintptr_t)ctx->dY) & 15))) {
        const size_t xb = (size_t)ctx->n * ctx->m * sizeof(double), yb = (size_t)ctx->n_t * ctx->m * sizeof(double);
        DDB_TRY(ctx->view_X.reserve(xb));
        DDB_TRY(ctx->view_Y.reserve(yb));
        DDB_CUDA_TRY(cudaMemcpyAsync(ctx->view_X.p, ctx->dX, xb, cudaMemcpyDeviceToDevice, ctx->stream));
        DDB_CUDA_TRY(cudaMemcpyAsync(ctx->view_Y.p, ctx->dY, yb, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->dX = ctx->view_X.as<double>();
        ctx->dY = ctx->view_Y.as<double>();
    }
    ctx->have_gram = false;
    ctx->imp_k = 0;
    ctx->term_scale_on = false;
    return DDB_OK;
}

__global__ void permute_kernel(const double* __restrict__ in, int w, const int64_t* __restrict__ perm, int64_t m,
                               double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * w) return;
    const int64_t s = t / w;
    out[t] = in[perm[s] * w + (t % w)];
}

int ddb_permute_samples(ddb_ctx* ctx, const int64_t* perm) {
    CHECK_CTX(ctx);
    if (ctx->range_on) {
        set_error("ddb_permute_samples: reset the sample range first");
        return DDB_BAD_STATE;
    }
    if (!ctx->dX || !ctx->dY || ctx->m <= 0 || !perm) {
        set_error("ddb_permute_samples: no data bound or null permutation");
        return DDB_BAD_STATE;
    }
    const int64_t m = ctx->m;
    std::vector<char> seen((size_t)m, 0);
    for (int64_t s = 0; s < m; ++s) {
        if (perm[s] < 0 || perm[s] >= m || seen[(size_t)perm[s]]) {
            set_error("ddb_permute_samples: perm is not a permutation of 0..m-1");
            return DDB_INVALID_ARG;
        }
        seen[(size_t)perm[s]] = 1;
    }
    const int n = ctx->n, n_t = ctx->n_t;
    ddb::DevBuf px, py, pp;
    int st = px.reserve((size_t)n * m * 8);
    if (st == DDB_OK) st = py.reserve((size_t)n_t * m * 8);
    if (st == DDB_OK) st = pp.reserve((size_t)m * 8);
    if (st != DDB_OK) {
        px.release(), py.release(), pp.release();
        return st;
    }
    cudaMemcpyAsync(pp.p, perm, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream);
    permute_kernel<<<(unsigned)((m * n + 255) / 256), 256, 0, ctx->stream>>>(ctx->dX, n, pp.as<int64_t>(), m, px.as<double>());
    permute_kernel<<<(unsigned)((m * n_t + 255) / 256), 256, 0, ctx->stream>>>(ctx->dY, n_t, pp.as<int64_t>(), m, py.as<double>());
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    pp.release();
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
        px.release(), py.release();
        set_error("ddb_permute_samples: %s", cudaGetErrorString(e));
        return DDB_CUDA_ERROR;
    }
    ctx->own_X.release();
    ctx->own_Y.release();
    ctx->own_X = px;
    ctx->own_Y = py;
    ctx->dX = ctx->own_X.as<double>();
    ctx->dY = ctx->own_Y.as<double>();
    ctx->have_gram = false;
    ctx->imp_k = 0;
    ctx->term_scale_on = false;
    return DDB_OK;
}

__global__ void scale_terms_kernel(double* __restrict__ packed, int k, int n_t, const double* __restrict__ scale) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t kk = (size_t)k * k, kn = (size_t)k * n_t;
    if (t < kk) {
        packed[t] /= scale[t % k] * scale[t / k];
    } else if (t < kk + kn) {
        packed[t] /= scale[(t - kk) % k];
    }
}

int ddb_gram_scale_terms(ddb_ctx* ctx, const double* scale) {
    CHECK_CTX(ctx);
    if (!ctx->have_gram || !ctx->d_packed.p) {
        set_error("ddb_gram_scale_terms: no normal-equation blocks in the context");
        return DDB_BAD_STATE;
    }
    if (ctx->term_scale_on) {
        set_error("ddb_gram_scale_terms: the blocks are already scaled");
        return DDB_BAD_STATE;
    }
    if (!scale) {
        set_error("ddb_gram_scale_terms: null scale");
        return DDB_INVALID_ARG;
    }
    for (int j = 0; j < ctx->k; ++j)
        if (!(scale[j] > 0.0) || !std::isfinite(scale[j])) {
            set_error("ddb_gram_scale_terms: scale[%d] must be positive and finite", j);
            return DDB_INVALID_ARG;
        }
    DDB_TRY(ctx->d_term_scale.reserve((size_t)ctx->k * 8));
    DDB_CUDA_TRY(cudaMemcpyAsync(ctx->d_term_scale.p, scale, (size_t)ctx->k * 8, cudaMemcpyHostToDevice, ctx->stream));
    const size_t tot = (size_t)ctx->k * ctx->k + (size_t)ctx->k * ctx->n_t;
    scale_terms_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_packed.as<double>(), ctx->k, ctx->n_t,
                                                                             ctx->d_term_scale.as<double>());
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->term_scale_on = true;
    ctx->term_shift_on = false;
    ctx->imp_k = 0;
    return DDB_OK;
}

int ddb_residual(ddb_ctx* ctx, const double* coef, double* rss) {
    CHECK_CTX(ctx);
    return residual_run(ctx, coef, nullptr, rss);
}

int ddb_residual_affine(ddb_ctx* ctx, const double* coef, const double* offset, double* rss) {
    CHECK_CTX(ctx);
    return residual_run(ctx, coef, offset, rss);
}

// ---- min / max of every library term over the current sample view (UnitRangeTransform) ----------------------
// monotone map double -> signed 64-bit key, so that integer atomicMin / atomicMax order doubles (no NaN in the data:
// check_domain rejects them)
__device__ __forceinline__ long long dkey(double v) {
    const long long b = __double_as_longlong(v);
    return b >= 0 ? b : (b ^ 0x7FFFFFFFFFFFFFFFLL);
}
static double dkey_inv(long long k) {
    const long long b = k >= 0 ? k : (k ^ 0x7FFFFFFFFFFFFFFFLL);
    double v;
    memcpy(&v, &b, 8);
    return v;
}
constexpr int kMMSamples = 32;   // samples per tile
constexpr int kMMTerms = 16;     // terms per thread (k <= 256 * 16)
__global__ void __launch_bounds__(256) term_minmax_kernel(const double* __restrict__ X, int n, int64_t m, const uint16_t* __restrict__ factors,
                                                          int k, long long* __restrict__ kmin, long long* __restrict__ kmax) {
    extern __shared__ double xs[];  // kMMSamples x (n + 1): the states of the tile, then 1.0 (missing factors)
    const int tid = threadIdx.x;
    const int pitch = n + 1;
    double mn[kMMTerms], mx[kMMTerms];
#pragma unroll
    for (int q = 0; q < kMMTerms; ++q) mn[q] = INFINITY, mx[q] = -INFINITY;
    const int64_t ntiles = (m + kMMSamples - 1) / kMMSamples;
    for (int64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const int64_t s0 = tl * kMMSamples;
        const int valid = (int)min((int64_t)kMMSamples, m - s0);
        __syncthreads();
        for (int e = tid; e < valid * n; e += 256) xs[(e / n) * pitch + e % n] = X[s0 * n + e];
        for (int s = tid; s < valid; s += 256) xs[s * pitch + n] = 1.0;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kMMTerms; ++q) {
            const int j = tid + 256 * q;
            if (j >= k) break;
            const uint4 w4 = *reinterpret_cast<const uint4*>(factors + (size_t)j * kMaxFactors);
            const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
            int f[kMaxFactors];
#pragma unroll
            for (int d = 0; d < kMaxFactors; ++d) {
                const uint32_t c = (w[d >> 1] >> (16 * (d & 1))) & 0xFFFFu;
                f[d] = (c == kNoFactor) ? n : (int)c;
            }
            for (int s = 0; s < valid; ++s) {
                const double* xr = xs + s * pitch;
                double p = xr[f[0]];
#pragma unroll
                for (int d = 1; d < kMaxFactors; ++d) p *= xr[f[d]];
                mn[q] = fmin(mn[q], p);
                mx[q] = fmax(mx[q], p);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kMMTerms; ++q) {
        const int j = tid + 256 * q;
        if (j < k && mn[q] <= mx[q]) {
            atomicMin(kmin + j, dkey(mn[q]));
            atomicMax(kmax + j, dkey(mx[q]));
        }
    }
}

int ddb_term_minmax(ddb_ctx* ctx, double* mn, double* mx) {
    CHECK_CTX(ctx);
    if (!mn || !mx || ctx->k <= 0 || !ctx->dX || ctx->m <= 0) {
        set_error("ddb_term_minmax: library and data must be set, outputs non-null");
        return DDB_INVALID_ARG;
    }
    const int k = ctx->k, n = ctx->n;
    if (k > 256 * kMMTerms) {
        set_error("ddb_term_minmax: libraries of more than %d terms are not supported", 256 * kMMTerms);
        return DDB_UNSUPPORTED;
    }
    if (ctx->factors_nt != ctx->n_t || !ctx->d_factors.p) DDB_TRY(gram_ensure_factors(ctx, ((k + ctx->n_t + 127) / 128) * 128));
    DDB_TRY(ctx->d_stats.reserve((size_t)2 * k * 8));
    long long* kmin = ctx->d_stats.as<long long>();
    long long* kmax = kmin + k;
    std::vector<long long> init(2 * (size_t)k);
    for (int j = 0; j < k; ++j) init[j] = 0x7FFFFFFFFFFFFFFFLL, init[k + j] = (long long)0x8000000000000000ULL;
    cudaStream_t st = ctx->stream;
    DDB_CUDA_TRY(cudaMemcpyAsync(kmin, init.data(), init.size() * 8, cudaMemcpyHostToDevice, st));
    const size_t smem = (size_t)kMMSamples * (n + 1) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("ddb_term_minmax: too many states for the shared-memory tile");
        return DDB_UNSUPPORTED;
    }
    DDB_CUDA_TRY(cudaFuncSetAttribute(term_minmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (ctx->m + kMMSamples - 1) / kMMSamples;
    const int nblk = (int)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * 4);
    term_minmax_kernel<<<nblk, 256, smem, st>>>(ctx->dX, n, ctx->m, ctx->d_factors.as<uint16_t>(), k, kmin, kmax);
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaMemcpyAsync(init.data(), kmin, init.size() * 8, cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    for (int j = 0; j < k; ++j) mn[j] = dkey_inv(init[j]), mx[j] = dkey_inv(init[k + j]);
    return DDB_OK;
}

// Centred sums of squares of every library term over the current view: css_j = sum_s (theta_j(s) - center_j)^2, the
// second pass of a two-pass variance (what `std` of the materialised row does in the reference; the one-pass form
// (sum theta^2 - (sum theta)^2 / m) cancels when a term's mean is much larger than its spread).  Per-CTA partial sums,
// added in CTA order on the host: bit-reproducible.
__global__ void __launch_bounds__(256) term_css_kernel(const double* __restrict__ X, int n, int64_t m, const uint16_t* __restrict__ factors,
                                                       int k, const double* __restrict__ center, double* __restrict__ partial) {
    extern __shared__ double xs[];  // kMMSamples x (n + 1): the states of the tile, then 1.0 (missing factors)
    const int tid = threadIdx.x;
    const int pitch = n + 1;
    double acc[kMMTerms];
#pragma unroll
    for (int q = 0; q < kMMTerms; ++q) acc[q] = 0.0;
    const int64_t ntiles = (m + kMMSamples - 1) / kMMSamples;
    for (int64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const int64_t s0 = tl * kMMSamples;
        const int valid = (int)min((int64_t)kMMSamples, m - s0);
        __syncthreads();
        for (int e = tid; e < valid * n; e += 256) xs[(e / n) * pitch + e % n] = X[s0 * n + e];
        for (int s = tid; s < valid; s += 256) xs[s * pitch + n] = 1.0;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kMMTerms; ++q) {
            const int j = tid + 256 * q;
            if (j >= k) break;
            const uint4 w4 = *reinterpret_cast<const uint4*>(factors + (size_t)j * kMaxFactors);
            const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
            int f[kMaxFactors];
#pragma unroll
            for (int d = 0; d < kMaxFactors; ++d) {
                const uint32_t c = (w[d >> 1] >> (16 * (d & 1))) & 0xFFFFu;
                f[d] = (c == kNoFactor) ? n : (int)c;
            }
            const double cj = center[j];
            for (int s = 0; s < valid; ++s) {
                const double* xr = xs + s * pitch;
                double p = xr[f[0]];
#pragma unroll
                for (int d = 1; d < kMaxFactors; ++d) p *= xr[f[d]];
                const double r = p - cj;
                acc[q] = fma(r, r, acc[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kMMTerms; ++q) {
        const int j = tid + 256 * q;
        if (j < k) partial[(size_t)blockIdx.x * k + j] = acc[q];
    }
}

int ddb_term_css(ddb_ctx* ctx, const double* center, double* css) {
    CHECK_CTX(ctx);
    if (!center || !css || ctx->k <= 0) {
        set_error("ddb_term_css: library must be set, center and css non-null");
        return DDB_INVALID_ARG;
    }
    const int k = ctx->k, n = ctx->n;
    if (ctx->view_empty) {  // no sample in this rank's view
        for (int j = 0; j < k; ++j) css[j] = 0.0;
        return DDB_OK;
    }
    if (!ctx->dX || ctx->m <= 0) {
        set_error("ddb_term_css: no data bound");
        return DDB_BAD_STATE;
    }
    if (k > 256 * kMMTerms) {
        set_error("ddb_term_css: libraries of more than %d terms are not supported", 256 * kMMTerms);
        return DDB_UNSUPPORTED;
    }
    if (ctx->factors_nt != ctx->n_t || !ctx->d_factors.p) DDB_TRY(gram_ensure_factors(ctx, ((k + ctx->n_t + 127) / 128) * 128));
    const size_t smem = (size_t)kMMSamples * (n + 1) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("ddb_term_css: too many states for the shared-memory tile");
        return DDB_UNSUPPORTED;
    }
    const int64_t ntiles = (ctx->m + kMMSamples - 1) / kMMSamples;
    const int nblk = (int)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * 4);
    DDB_TRY(ctx->d_stats.reserve(((size_t)nblk + 1) * k * 8));
    double* d_center = ctx->d_stats.as<double>();
    double* d_part = d_center + k;
    cudaStream_t st = ctx->stream;
    DDB_CUDA_TRY(cudaMemcpyAsync(d_center, center, (size_t)k * 8, cudaMemcpyHostToDevice, st));
    DDB_CUDA_TRY(cudaFuncSetAttribute(term_css_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    term_css_kernel<<<nblk, 256, smem, st>>>(ctx->dX, n, ctx->m, ctx->d_factors.as<uint16_t>(), k, d_center, d_part);
    DDB_CUDA_TRY(cudaGetLastError());
    std::vector<double> part((size_t)nblk * k);
    DDB_CUDA_TRY(cudaMemcpyAsync(part.data(), d_part, part.size() * 8, cudaMemcpyDeviceToHost, st));
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    for (int j = 0; j < k; ++j) {
        double s = 0.0;
        for (int b = 0; b < nblk; ++b) s += part[(size_t)b * k + j];  // fixed order
        css[j] = s;
    }
    return DDB_OK;
}

// G~ = S (G - mu t' - t mu' + m mu mu') S,  R~ = S (R - mu ysum'),  S = diag(1 / scale): the blocks of
// theta~_j = (theta_j - mu_j) / scale_j from those of theta  (t = Theta 1, ysum = Y 1)
__global__ void affine_terms_kernel(double* __restrict__ packed, int k, int n_t, const double* __restrict__ mu,
                                    const double* __restrict__ scale, const double* __restrict__ tsum,
                                    const double* __restrict__ ysum, double m) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t kk = (size_t)k * k, kn = (size_t)k * n_t;
    if (t < kk) {
        const int i = (int)(t % k), j = (int)(t / k);
        // symmetric in (i, j) term by term: the result stays exactly symmetric
        const double v = packed[t] - (mu[i] * tsum[j] + mu[j] * tsum[i]) + m * (mu[i] * mu[j]);
        packed[t] = v / (scale[i] * scale[j]);
    } else if (t < kk + kn) {
        const int i = (int)((t - kk) % k), e = (int)((t - kk) / k);
        packed[t] = (packed[t] - mu[i] * ysum[e]) / scale[i];
    }
}

int ddb_gram_affine_terms(ddb_ctx* ctx, const double* shift, const double* scale, const double* tsum, const double* ysum,
                          int64_t m_total) {
    CHECK_CTX(ctx);
    if (!ctx->have_gram || !ctx->d_packed.p) {
        set_error("ddb_gram_affine_terms: no normal-equation blocks in the context");
        return DDB_BAD_STATE;
    }
    if (ctx->term_scale_on) {
        set_error("ddb_gram_affine_terms: the blocks are already transformed");
        return DDB_BAD_STATE;
    }
    if (!shift || !scale || !tsum || !ysum || m_total <= 0) {
        set_error("ddb_gram_affine_terms: null argument");
        return DDB_INVALID_ARG;
    }
    const int k = ctx->k, n_t = ctx->n_t;
    for (int j = 0; j < k; ++j)
        if (!(scale[j] > 0.0) || !std::isfinite(scale[j]) || !std::isfinite(shift[j])) {
            set_error("ddb_gram_affine_terms: scale[%d] must be positive and finite, shift finite", j);
            return DDB_INVALID_ARG;
        }
    DDB_TRY(ctx->d_term_scale.reserve((size_t)k * 8));
    DDB_TRY(ctx->d_term_shift.reserve((size_t)k * 8));
    DDB_TRY(ctx->d_stats.reserve(((size_t)k + n_t) * 8));
    cudaStream_t st = ctx->stream;
    DDB_CUDA_TRY(cudaMemcpyAsync(ctx->d_term_scale.p, scale, (size_t)k * 8, cudaMemcpyHostToDevice, st));
    DDB_CUDA_TRY(cudaMemcpyAsync(ctx->d_term_shift.p, shift, (size_t)k * 8, cudaMemcpyHostToDevice, st));
    DDB_CUDA_TRY(cudaMemcpyAsync(ctx->d_stats.p, tsum, (size_t)k * 8, cudaMemcpyHostToDevice, st));
    DDB_CUDA_TRY(cudaMemcpyAsync(ctx->d_stats.as<double>() + k, ysum, (size_t)n_t * 8, cudaMemcpyHostToDevice, st));
    const size_t tot = (size_t)k * k + (size_t)k * n_t;
    affine_terms_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ctx->d_packed.as<double>(), k, n_t,
                                                                       ctx->d_term_shift.as<double>(), ctx->d_term_scale.as<double>(),
                                                                       ctx->d_stats.as<double>(), ctx->d_stats.as<double>() + k,
                                                                       (double)m_total);
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaStreamSynchronize(st));
    ctx->term_scale_on = true;
    ctx->term_shift_on = true;
    ctx->imp_k = 0;
    return DDB_OK;
}

int ddb_bind_device(ddb_ctx* ctx, const double* dX, const double* dY, int n_t, int64_t m) {
    CHECK_CTX(ctx);
    DDB_TRY(check_data_args(ctx, dX, dY, n_t, m, "ddb_bind_device"));
    if (ctx->n_raw > 0) {  // dX holds raw states: expand once into the context's own buffer
        DDB_TRY(ctx->own_X.reserve((size_t)ctx->n * m * sizeof(double)));
        DDB_TRY(expand_features(ctx, dX, 0, m, ctx->stream));
        DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        dX = ctx->own_X.as<double>();
    }
    ctx->dX = dX;
    ctx->dY = dY;
    ctx->range_on = false;
    ctx->view_empty = false;
    ctx->n_t = n_t;
    ctx->m = m;
    ctx->have_gram = false;
    ctx->imp_k = 0;
    ctx->term_scale_on = false;
    return DDB_OK;
}

int ddb_generate(ddb_ctx* ctx, int system_id, int n_states, int64_t m, int64_t first_sample, uint64_t seed,
                 double noise_std) {
    CHECK_CTX(ctx);
    return generate_run(ctx, system_id, n_states, m, first_sample, seed, noise_std);
}

int ddb_download_data(ddb_ctx* ctx, double* X, double* Y) {
    CHECK_CTX(ctx);
    if (!ctx->dX || !ctx->dY) {
        set_error("ddb_download_data: no data bound");
        return DDB_BAD_STATE;
    }
    if (X)
        DDB_CUDA_TRY(cudaMemcpyAsync(X, ctx->dX, (size_t)ctx->n * ctx->m * sizeof(double), cudaMemcpyDeviceToHost,
                                     ctx->stream));
    if (Y)
        DDB_CUDA_TRY(cudaMemcpyAsync(Y, ctx->dY, (size_t)ctx->n_t * ctx->m * sizeof(double), cudaMemcpyDeviceToHost,
                                     ctx->stream));
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DDB_OK;
}

int64_t ddb_gram_count(const ddb_ctx* ctx) {
    if (!ctx) return 0;
    return (int64_t)ctx->k * ctx->k + (int64_t)ctx->k * ctx->n_t + ctx->n_t;
}

static void reset_allreduce_timings(ddb_ctx* ctx) {  // a new Gram starts a new solve
    ctx->timings.allreduce_ms = 0.0;
    ctx->timings.allreduce_bytes = 0.0;
    ctx->timings.allreduce_calls = 0;
}

int ddb_gram(ddb_ctx* ctx, double* dPacked) {
    CHECK_CTX(ctx);
    reset_allreduce_timings(ctx);
    return gram_run(ctx, dPacked);
}

int ddb_gram_buffer(ddb_ctx* ctx, double** dPacked) {
    CHECK_CTX(ctx);
    if (!dPacked) {
        set_error("ddb_gram_buffer: null output");
        return DDB_INVALID_ARG;
    }
    if (ctx->k <= 0 || ctx->n_t <= 0) {
        set_error("ddb_gram_buffer: library and data must be set first");
        return DDB_BAD_STATE;
    }
    DDB_TRY(ctx->d_packed.reserve((size_t)ddb_gram_count(ctx) * sizeof(double)));
    *dPacked = ctx->d_packed.as<double>();
    return DDB_OK;
}

int ddb_gram_download(ddb_ctx* ctx, double* G, double* R, double* yy) {
    CHECK_CTX(ctx);
    if (!ctx->d_packed.p) {
        set_error("ddb_gram_download: no normal-equation blocks in the context");
        return DDB_BAD_STATE;
    }
    const size_t k = ctx->k, nt = ctx->n_t;
    const double* p = ctx->d_packed.as<double>();
    if (G) DDB_CUDA_TRY(cudaMemcpyAsync(G, p, k * k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (R) DDB_CUDA_TRY(cudaMemcpyAsync(R, p + k * k, k * nt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (yy)
        DDB_CUDA_TRY(cudaMemcpyAsync(yy, p + k * k + k * nt, nt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DDB_OK;
}

int ddb_gram_upload(ddb_ctx* ctx, const double* G, const double* R, const double* yy, int n_t) {
    CHECK_CTX(ctx);
    if (ctx->k <= 0 || !G || !R || !yy || n_t <= 0) {
        set_error("ddb_gram_upload: library must be installed and G, R, yy non-null");
        return DDB_INVALID_ARG;
    }
    if (ctx->n_t != n_t) {
        ctx->n_t = n_t;
        ctx->factors_nt = -1;
        ctx->plan_m = -1;
    }
    const size_t k = ctx->k, nt = n_t;
    DDB_TRY(ctx->d_packed.reserve((size_t)ddb_gram_count(ctx) * sizeof(double)));
    double* p = ctx->d_packed.as<double>();
    // cudaMemcpyDefault: the three sources may be host or device pointers (e.g. an all-reduced device buffer)
    DDB_CUDA_TRY(cudaMemcpyAsync(p, G, k * k * sizeof(double), cudaMemcpyDefault, ctx->stream));
    DDB_CUDA_TRY(cudaMemcpyAsync(p + k * k, R, k * nt * sizeof(double), cudaMemcpyDefault, ctx->stream));
    DDB_CUDA_TRY(cudaMemcpyAsync(p + k * k + k * nt, yy, nt * sizeof(double), cudaMemcpyDefault, ctx->stream));
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->have_gram = true;
    ctx->imp_k = 0;
    return DDB_OK;
}

int64_t ddb_gram_plan_describe(int kaug, int64_t m, int nctas, int64_t* segments, int64_t cap, int64_t* info) {
    if (kaug <= 0 || m <= 0 || nctas < 0) {
        set_error("ddb_gram_plan_describe: kaug, m must be positive");
        return DDB_INVALID_ARG;
    }
    const int bt = (kaug <= 64) ? 64 : 128;
    if (nctas == 0) nctas = (bt == 64) ? 148 * 3 : 148;
    GramPlan pl;
    build_gram_plan(pl, kaug, m, nctas, bt, 16);
    if (info) {
        info[0] = pl.bt, info[1] = pl.ks, info[2] = pl.nblk, info[3] = pl.npairs, info[4] = pl.nstages, info[5] = pl.nslots;
    }
    int64_t w = 0;
    for (int c = 0; c < pl.nctas; ++c)
        for (int sg = pl.cta_seg_begin[c]; sg < pl.cta_seg_begin[c + 1]; ++sg, ++w) {
            if (segments && w < cap) {
                const GramSegment& g = pl.segments[sg];
                int64_t* o = segments + 6 * w;
                o[0] = c, o[1] = g.bi, o[2] = g.bj, o[3] = g.slot, o[4] = g.stage_begin, o[5] = g.stage_end;
            }
        }
    return w;
}

int64_t ddb_moment_plan_describe(int n, int k, const uint8_t* exps, int n_t, int32_t* src, int64_t src_cap, uint16_t* factors,
                                 int64_t factors_cap, int32_t* pairs, int64_t pairs_cap, int64_t* info) {
    if (n <= 0 || k <= 0 || n_t <= 0 || !exps) {
        set_error("ddb_moment_plan_describe: n, k, n_t must be positive and exps non-null");
        return DDB_INVALID_ARG;
    }
    return moment_plan_describe(exps, n, k, n_t, src, src_cap, factors, factors_cap, pairs, pairs_cap, info);
}

int64_t ddb_moment_builder_describe(int n, int k, const uint8_t* exps, int n_t, uint32_t* optab, int64_t optab_cap,
                                    uint8_t* opcnt, int64_t opcnt_cap, int32_t* block_ops, int64_t block_cap) {
    if (n <= 0 || k <= 0 || n_t <= 0 || !exps) {
        set_error("ddb_moment_builder_describe: n, k, n_t must be positive and exps non-null");
        return DDB_INVALID_ARG;
    }
    return moment_builder_describe(exps, n, k, n_t, optab, optab_cap, opcnt, opcnt_cap, block_ops, block_cap);
}

int ddb_pow_factors_describe(int n, int k, const uint8_t* exps, int n_t, uint16_t* factors, int* degree) {
    if (n <= 0 || k <= 0 || n_t <= 0 || !exps) {
        set_error("ddb_pow_factors_describe: n, k, n_t must be positive and exps non-null");
        return DDB_INVALID_ARG;
    }
    int maxdeg = 0;
    for (int j = 0; j < k; ++j) {
        int d = 0;
        for (int i = 0; i < n; ++i) d += exps[(size_t)i + (size_t)n * j];
        maxdeg = std::max(maxdeg, d);
    }
    if (degree) *degree = maxdeg;
    std::vector<uint16_t> tab;
    const int nf = ddb::gram_pow_factors(exps, n, k, n_t, maxdeg, &tab);
    if (nf > 0 && factors) memcpy(factors, tab.data(), tab.size() * sizeof(uint16_t));
    return nf;
}

int64_t ddb_moment_wave_describe(int n, int k, const uint8_t* exps, int n_t, int64_t m, int nctas, int64_t* segments,
                                 int64_t cap, double* info) {
    if (n <= 0 || k <= 0 || n_t <= 0 || m <= 0 || nctas < 0 || !exps) {
        set_error("ddb_moment_wave_describe: n, k, n_t, m must be positive and exps non-null");
        return DDB_INVALID_ARG;
    }
    return moment_wave_describe(exps, n, k, n_t, m, nctas == 0 ? 148 : nctas, segments, cap, info);
}

int ddb_get_timings(const ddb_ctx* ctx, ddb_timings* out) {
    if (!ctx || !out) {
        set_error("ddb_get_timings: null argument");
        return DDB_INVALID_ARG;
    }
    *out = ctx->timings;
    return DDB_OK;
}

int ddb_get_stream(ddb_ctx* ctx, void** stream) {
    if (!ctx || !stream) {
        set_error("ddb_get_stream: null argument");
        return DDB_INVALID_ARG;
    }
    *stream = (void*)ctx->stream;
    return DDB_OK;
}

int ddb_synchronize(ddb_ctx* ctx) {
    CHECK_CTX(ctx);
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DDB_OK;
}

}  // extern "C"
