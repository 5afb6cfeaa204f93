// Arbitrary library terms through NVRTC: features written as CUDA C expressions of the raw variables.
//
// Replaces the code generation of the reference for terms the fixed feature kinds of ddb_set_features do not cover:
// `DataDrivenFunction` / `Symbolics.build_function` (src/basis/build_function.jl:6-38) turn every basis equation into
// a compiled Julia function of (implicits, states, parameters, independent variable, controls).  Here the host side
// hands the same expressions over as C source (Symbolics can emit it: `build_function(...; target = CTarget())`),
// NVRTC compiles ONE kernel for sm_100a that evaluates all features of a sample, and the installed monomial library
// then is a table of exponents over those features -- everything downstream (fused Theta + Gram, sweep, exact rss) is
// unchanged; the expanded features (n_feat x m doubles) are the only addition to HBM, Theta itself is still never
// built.  NVRTC and the driver API are bound at run time (libnvrtc.so.12, libcuda.so.1).
#include "common.cuh"
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace ddb {

struct UserFeat {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    DevBuf params;
    int n_params = 0;
    int n_feat = 0;
};

struct RtcApi {
    void* hrtc = nullptr;
    void* hcuda = nullptr;
    decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
    decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
    decltype(&nvrtcGetProgramLogSize) GetLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) GetLog = nullptr;
    decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
    decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
    decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
    decltype(&cuModuleLoadData) ModuleLoadData = nullptr;
    decltype(&cuModuleGetFunction) ModuleGetFunction = nullptr;
    decltype(&cuModuleUnload) ModuleUnload = nullptr;
    decltype(&cuLaunchKernel) LaunchKernel = nullptr;
};
static RtcApi g_rtc;
static std::mutex g_rtc_mutex;

static int rtc_load(bool need_driver) {
    std::lock_guard<std::mutex> lock(g_rtc_mutex);
    if (!g_rtc.hrtc) {
        const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
        void* h = nullptr;
        for (const char* nm : names)
            if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!h) {
            set_error("NVRTC is not available (dlopen libnvrtc.so.12: %s)", dlerror());
            return DDB_UNSUPPORTED;
        }
#define DDB_RTC_SYM(field, name)                                              \
    g_rtc.field = reinterpret_cast<decltype(g_rtc.field)>(dlsym(h, name));    \
    if (!g_rtc.field) {                                                       \
        set_error("NVRTC symbol %s is missing", name);                        \
        return DDB_UNSUPPORTED;                                               \
    }
        DDB_RTC_SYM(CreateProgram, "nvrtcCreateProgram")
        DDB_RTC_SYM(CompileProgram, "nvrtcCompileProgram")
        DDB_RTC_SYM(GetLogSize, "nvrtcGetProgramLogSize")
        DDB_RTC_SYM(GetLog, "nvrtcGetProgramLog")
        DDB_RTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        DDB_RTC_SYM(GetCUBIN, "nvrtcGetCUBIN")
        DDB_RTC_SYM(DestroyProgram, "nvrtcDestroyProgram")
        DDB_RTC_SYM(GetErrorString, "nvrtcGetErrorString")
        g_rtc.hrtc = h;
    }
    if (need_driver && !g_rtc.hcuda) {
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            set_error("the CUDA driver library is not available (dlopen libcuda.so.1: %s)", dlerror());
            return DDB_CUDA_ERROR;
        }
        DDB_RTC_SYM(ModuleLoadData, "cuModuleLoadData")
        DDB_RTC_SYM(ModuleGetFunction, "cuModuleGetFunction")
        DDB_RTC_SYM(ModuleUnload, "cuModuleUnload")
        DDB_RTC_SYM(LaunchKernel, "cuLaunchKernel")
#undef DDB_RTC_SYM
        g_rtc.hcuda = h;
    }
    return DDB_OK;
}

static std::string feature_source(const char* body) {
    std::string src =
        "extern \"C\" __global__ void ddb_user_features(const double* __restrict__ raw, int n_raw, long long m,\n"
        "                                             const double* __restrict__ p, double* __restrict__ out, int n_feat) {\n"
        "    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;\n"
        "    if (s >= m) return;\n"
        "    const double* x = raw + s * n_raw;   // the raw variables of sample s: [states | controls | t]\n"
        "    double* f = out + s * n_feat;        // its features\n"
        "    {\n";
    src += body;
    src += "\n    }\n}\n";
    return src;
}

// compiles the feature kernel; on success fills `cubin`.  The compiler log (errors, warnings) goes to `log`.
static int compile_features(const char* body, std::vector<char>* cubin, std::string* log) {
    DDB_TRY(rtc_load(false));
    const std::string src = feature_source(body);
    nvrtcProgram prog = nullptr;
    nvrtcResult r = g_rtc.CreateProgram(&prog, src.c_str(), "ddb_user_features.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) {
        set_error("nvrtcCreateProgram failed: %s", g_rtc.GetErrorString(r));
        return DDB_CUDA_ERROR;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false"};
    r = g_rtc.CompileProgram(prog, 3, opts);
    size_t ls = 0;
    g_rtc.GetLogSize(prog, &ls);
    if (log && ls > 1) {
        log->resize(ls);
        g_rtc.GetLog(prog, &(*log)[0]);
    }
    int status = DDB_OK;
    if (r != NVRTC_SUCCESS) {
        set_error("the feature expressions do not compile (%s): %.700s", g_rtc.GetErrorString(r), log ? log->c_str() : "");
        status = DDB_INVALID_ARG;
    } else if (cubin) {
        size_t cs = 0;
        g_rtc.GetCUBINSize(prog, &cs);
        cubin->resize(cs);
        g_rtc.GetCUBIN(prog, cubin->data());
    }
    g_rtc.DestroyProgram(&prog);
    return status;
}

int userfeat_nfeat(const ddb_ctx* ctx) { return ctx->userfeat ? ctx->userfeat->n_feat : 0; }

void userfeat_free(ddb_ctx* ctx) {
    if (ctx->userfeat) {
        if (ctx->userfeat->mod && g_rtc.ModuleUnload) g_rtc.ModuleUnload(ctx->userfeat->mod);
        ctx->userfeat->params.release();
        delete ctx->userfeat;
        ctx->userfeat = nullptr;
    }
}

// expands raw samples (device pointer d_raw, mc samples) into out (mc x n_feat)
int userfeat_expand(ddb_ctx* ctx, const double* d_raw, int64_t mc, double* out, cudaStream_t st) {
    UserFeat* U = ctx->userfeat;
    int n_raw = ctx->n_raw, n_feat = U->n_feat;
    long long m = mc;
    const double* p = U->params.as<double>();
    void* args[] = {(void*)&d_raw, (void*)&n_raw, (void*)&m, (void*)&p, (void*)&out, (void*)&n_feat};
    const unsigned grid = (unsigned)((mc + 127) / 128);
    const CUresult r = g_rtc.LaunchKernel(U->fn, grid, 1, 1, 128, 1, 1, 0, (CUstream)st, args, nullptr);
    if (r != CUDA_SUCCESS) {
        set_error("cuLaunchKernel(ddb_user_features) failed: CUresult %d", (int)r);
        return DDB_CUDA_ERROR;
    }
    return DDB_OK;
}

}  // namespace ddb

using namespace ddb;

extern "C" {

int ddb_features_source_check(const char* body, char* log, int log_cap) {
    if (!body) {
        set_error("ddb_features_source_check: null source");
        return DDB_INVALID_ARG;
    }
    std::string lg;
    const int status = compile_features(body, nullptr, &lg);
    if (log && log_cap > 0) {
        const size_t n = std::min<size_t>(lg.size(), (size_t)log_cap - 1);
        memcpy(log, lg.data(), n);
        log[n] = 0;
    }
    return status;
}

int ddb_set_features_source(ddb_ctx* ctx, int n_raw, int n_feat, const char* body, const double* params, int n_params) {
    if (!ctx || !body || n_raw <= 0 || n_feat <= 0 || n_params < 0 || (n_params > 0 && !params)) {
        set_error("ddb_set_features_source: need a context, the source, n_raw, n_feat > 0 and the parameter values");
        return DDB_INVALID_ARG;
    }
    DDB_CUDA_TRY(cudaSetDevice(ctx->device));
    DDB_CUDA_TRY(cudaFree(0));  // make sure the primary context is current for the driver calls below
    std::vector<char> cubin;
    std::string lg;
    DDB_TRY(compile_features(body, &cubin, &lg));
    DDB_TRY(rtc_load(true));
    userfeat_free(ctx);
    UserFeat* U = new UserFeat();
    CUresult r = g_rtc.ModuleLoadData(&U->mod, cubin.data());
    if (r == CUDA_SUCCESS) r = g_rtc.ModuleGetFunction(&U->fn, U->mod, "ddb_user_features");
    if (r != CUDA_SUCCESS) {
        set_error("loading the compiled feature kernel failed: CUresult %d", (int)r);
        delete U;
        return DDB_CUDA_ERROR;
    }
    U->n_feat = n_feat;
    U->n_params = n_params;
    int st = U->params.reserve((size_t)std::max(1, n_params) * 8);
    if (st == DDB_OK && n_params > 0) {
        if (cudaMemcpyAsync(U->params.p, params, (size_t)n_params * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            st = DDB_CUDA_ERROR;
    }
    if (st != DDB_OK) {
        U->params.release();
        delete U;
        return st;
    }
    ctx->userfeat = U;
    ctx->n_raw = n_raw;
    ctx->have_gram = false;
    ctx->imp_k = 0;
    ctx->term_scale_on = false;
    ctx->dX = ctx->dY = nullptr;  // the library must be (re)installed over n_feat "states"; data must be (re)bound
    ctx->range_on = false;
    ctx->m = 0;
    return DDB_OK;
}

}  // extern "C"
