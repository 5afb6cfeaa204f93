// Device-side synthetic trajectories so the benchmark inputs are born in HBM.
// One CTA per trajectory, one thread per state, fixed-step RK4; the targets are the exact ODE
// right-hand side at the saved states, which is what `DataDrivenProblem(sol::ODESolution)` stores
// in DX (reference src/problem/type.jl:577-586).
#include "common.cuh"

namespace ddb {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// counter-based standard normal: Box-Muller on two hashed uniforms
__device__ __forceinline__ double normal_from(uint64_t seed, uint64_t a, uint64_t b) {
    const uint64_t h1 = splitmix64(seed ^ splitmix64(a * 0xD1342543DE82EF95ull + b));
    const uint64_t h2 = splitmix64(h1 ^ 0xA5A5A5A5A5A5A5A5ull);
    const double u1 = ((h1 >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    const double u2 = (h2 >> 11) * (1.0 / 9007199254740992.0);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

struct GenArgs {
    double* X;
    double* Y;
    int system_id, n;
    int64_t m, first_sample;
    int samples_per_traj, stride, spinup;
    double h, noise_std;
    uint64_t seed;
};

template <int SYS>
__device__ __forceinline__ double rhs(const double* w, int i, int n) {
    if (SYS == 0) {  // Lorenz sigma=10, rho=28, beta=8/3
        const double x = w[0], y = w[1], z = w[2];
        return i == 0 ? 10.0 * (y - x) : (i == 1 ? x * (28.0 - z) - y : x * y - (8.0 / 3.0) * z);
    } else {  // Lorenz-96, F = 8, periodic
        const int ip = (i + 1 == n) ? 0 : i + 1;
        const int im1 = (i == 0) ? n - 1 : i - 1;
        const int im2 = (im1 == 0) ? n - 1 : im1 - 1;
        return (w[ip] - w[im2]) * w[im1] - w[i] + 8.0;
    }
}

template <int SYS>
__global__ void generate_kernel(const GenArgs a) {
    extern __shared__ double sm[];
    double* w0 = sm;
    double* w1 = sm + a.n;
    const int i = threadIdx.x;
    const bool act = i < a.n;
    const int64_t first_traj = a.first_sample / a.samples_per_traj;
    const int64_t traj = first_traj + blockIdx.x;
    double u = 0.0;
    if (act) {
        if (SYS == 0)
            u = (i == 0 ? 1.0 : 0.0) + normal_from(a.seed, (uint64_t)traj, (uint64_t)i);
        else
            u = 8.0 + 0.01 * normal_from(a.seed, (uint64_t)traj, (uint64_t)i);
    }
    const double h = a.h;
    const int64_t total_steps = (int64_t)a.spinup + (int64_t)a.samples_per_traj * a.stride;
    for (int64_t step = 0; step < total_steps; ++step) {
        // stage 1
        if (act) w0[i] = u;
        __syncthreads();
        double k = act ? rhs<SYS>(w0, i, a.n) : 0.0;
        if (step >= a.spinup && ((step - a.spinup) % a.stride) == 0) {
            const int64_t g = traj * a.samples_per_traj + (step - a.spinup) / a.stride - a.first_sample;
            if (act && g >= 0 && g < a.m) {
                a.X[g * a.n + i] = u;
                double y = k;
                if (a.noise_std > 0.0)
                    y += a.noise_std * normal_from(a.seed ^ 0x5EEDull, (uint64_t)(g + a.first_sample), (uint64_t)i);
                a.Y[g * a.n + i] = y;
            }
        }
        double acc = u + (h / 6.0) * k;
        if (act) w1[i] = u + 0.5 * h * k;
        __syncthreads();
        k = act ? rhs<SYS>(w1, i, a.n) : 0.0;  // stage 2
        acc += (h / 3.0) * k;
        if (act) w0[i] = u + 0.5 * h * k;
        __syncthreads();
        k = act ? rhs<SYS>(w0, i, a.n) : 0.0;  // stage 3
        acc += (h / 3.0) * k;
        if (act) w1[i] = u + h * k;
        __syncthreads();
        k = act ? rhs<SYS>(w1, i, a.n) : 0.0;  // stage 4
        u = acc + (h / 6.0) * k;
    }
}

int generate_run(ddb_ctx* ctx, int system_id, int n_states, int64_t m, int64_t first_sample, uint64_t seed,
                 double noise_std) {
    if (system_id != 0 && system_id != 1) {
        set_error("ddb_generate: unknown system %d (0 = Lorenz, 1 = Lorenz-96)", system_id);
        return DDB_INVALID_ARG;
    }
    const int n = system_id == 0 ? 3 : n_states;
    if (n < 3 || n > 1024 || m <= 0 || first_sample < 0) {
        set_error("ddb_generate: need 3 <= n <= 1024, m > 0, first_sample >= 0");
        return DDB_INVALID_ARG;
    }
    if (ctx->n_raw > 0) {
        set_error("ddb_generate: the synthetic systems are generated as states; remove the feature map first");
        return DDB_BAD_STATE;
    }
    if (ctx->n != 0 && ctx->n != n) {
        set_error("ddb_generate: system has %d states but the installed library has %d", n, ctx->n);
        return DDB_INVALID_ARG;
    }
    const size_t bytes = (size_t)n * m * sizeof(double);
    DDB_TRY(ctx->own_X.reserve(bytes));
    DDB_TRY(ctx->own_Y.reserve(bytes));
    GenArgs a;
    a.X = ctx->own_X.as<double>();
    a.Y = ctx->own_Y.as<double>();
    a.system_id = system_id;
    a.n = n;
    a.m = m;
    a.first_sample = first_sample;
    a.samples_per_traj = 1000;
    a.seed = seed;
    a.noise_std = noise_std;
    if (system_id == 0) {
        a.h = 0.01, a.stride = 1, a.spinup = 1000;
    } else {
        a.h = 0.01, a.stride = 5, a.spinup = 2000;
    }
    const int64_t t0 = first_sample / a.samples_per_traj;
    const int64_t t1 = (first_sample + m - 1) / a.samples_per_traj;
    const int ntraj = (int)(t1 - t0 + 1);
    const int threads = ((n + 31) / 32) * 32;
    const size_t smem = 2 * (size_t)n * sizeof(double);
    if (system_id == 0)
        generate_kernel<0><<<ntraj, threads, smem, ctx->stream>>>(a);
    else
        generate_kernel<1><<<ntraj, threads, smem, ctx->stream>>>(a);
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->range_on = false;
    ctx->view_empty = false;
    ctx->dX = a.X;
    ctx->dY = a.Y;
    ctx->n_t = n;
    ctx->m = m;
    ctx->have_gram = false;
    return DDB_OK;
}

}  // namespace ddb
