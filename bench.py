"""Benchmark of the sparse-identification hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c5]
                    [--alg STLSQ|ADMM|SR3] [--noise S] [--subs auto|all|none]

A "step" is one pass of the hot path over one batch of synthetic input: fused Theta+Gram over this rank's samples ->
sum all-reduce of the packed [G|R|yy] -> 41-threshold sweep -> exact-rss streaming pass -> all-reduce of the
candidate-rss table -> selector replay.  Both all-reduces run INSIDE the library (NCCL on the compute stream,
`ddb_allreduce_gram` / `ddb_allreduce_rss`); torch.distributed only carries the 128-byte NCCL id and the timing
reduction.  Headline workload c3 (the configuration the north-star target is quoted on; 10.24 GB, fits one GPU):
Lorenz-96, n = 64, degree-2 polynomial library (2145 terms), 10^7 samples in total, row-sharded over the ranks
(strong scaling), STLSQ.  One JSON line is printed by rank 0; the other BASELINE configs ride along under "sub":
C2 (Lorenz, 56 terms), C4 (ADMM and SR3 on the C3 data), C3 with noisy derivatives (the sweep's hard case: large
active sets) and C5 (exact-DMD Gram path), each with its own stage times and roofline fraction.
"""
import argparse
import json
import os
import sys
import threading
import time

if "reference" in sys.argv[1:] or os.environ.get("WORLD_SIZE", "1") == "1":  # (N = 1: the cpu_baseline leg of our arm)
    # The CPU arm uses every host core.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, and OpenBLAS
    # sizes its thread pool from the environment when NumPy / SciPy load it: raising the limit afterwards
    # (threadpoolctl) reports the full count but still ran the pivoted-QR sweep 5x slower (6.4 against 31 samples/s on
    # the same 8 cores), so the variables are set BEFORE NumPy is imported.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAMS = 10.0 ** np.arange(-5, -1 + 1e-9, 0.1)  # exp10.(-5:0.1:-1), 41 thresholds (README.md:55)
METRIC = "samples/sec through fused Theta+Gram; STLSQ lambda-sweep wall-time; % FP64 TC peak"
# FP64 tensor (DMMA) peak of this pool's B200s: not in MEASURED_PEAKS.json (bf16 only); measured with
# `make -C tools fp64_peak && tools/bin/fp64_peak` (profiles/fp64_peak_r1.json): DMMA issue-rate peak 37.0 TF =
# 148 SMs x 4 sub-partitions x 512 flop / 16 cycles x 1.965 GHz; cublasDgemm 8192^3 reaches 35.9 TF.
FP64_DMMA_PEAK_TFLOPS = 37.0
CUBLAS_DGEMM_TFLOPS = 35.9
# ncu dram__bytes_read.sum + dram__bytes_write.sum of the Gram kernels per sample at C3 (1e6-sample captures, per
# schedule; see profiles/): moment schedule gram_kernel; ring schedule gram_ring_kernel + diagonal-pair pass
DRAM_BYTES_PER_SAMPLE = {2: ((15057436160 + 71802880) / 1e6, "profiles/gram_traffic_r2k_c3_1e6.csv"),
                         1: ((1041148928 + 1886253824) / 1e6, "profiles/gram_traffic_r1g_c3_1e6.csv")}

WORKLOADS = {
    "c3": dict(system=1, n=64, degree=2, m=10_000_000, name="Lorenz-96 n=64, degree-2 library (2145 terms), 1e7 samples, 41-lambda STLSQ"),
    "c2": dict(system=0, n=3, degree=5, m=10_000_000, name="Lorenz, degree-5 library (56 terms), 1e7 samples, 41-lambda STLSQ"),
    "c5": dict(system=-1, n=4096, degree=1, m=1_000_000, name="exact-DMD Gram path: 4096 states x 1e6 snapshots, P = XX', Q = YX', K = Q P^-1"),
}


def hbm_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6548.2, "fallback (last MEASURED_PEAKS.json of this pool)"


def gram_flops_per_sample(k, n_t):
    """Algorithmic work of the Gram stage per sample: symmetric rank-1 update of G counted
    once (SYRK count) + Theta Y' + diag(Y Y'); monomial construction is not counted."""
    return k * (k + 1) + 2 * k * n_t + 2 * n_t


def host_threads():
    """All host cores for the CPU arm.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would
    silently run the NumPy/OpenBLAS reference single-threaded at N >= 2: the limit is lifted explicitly."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=n)
    except Exception:  # pragma: no cover
        pass
    used = n
    try:
        from threadpoolctl import threadpool_info

        blas = [p["num_threads"] for p in threadpool_info() if p.get("user_api") == "blas"]
        if blas:
            used = max(blas)
    except Exception:  # pragma: no cover
        pass
    return used


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.05)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (NumPy/SciPy restatement of the reference algorithm as written:
# materialise Theta, per-target pivoted-QR STLSQ, two rss passes per iteration) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(wl, seed=0):
    """One bounded sample of the workload on the CPU.  Returns (samples_per_s_equivalent, description)."""
    import oracle
    from ddb200 import systems

    n, deg = wl["n"], wl["degree"]
    ex = oracle.polynomial_exponents(n, deg)
    k = ex.shape[1]
    if wl["system"] == 1:
        m_s, nt_s = 4096, 2  # Theta 2145 x 4096; 2 of 64 targets (the reference loops targets sequentially)
        X, DX = systems.lorenz96_ensemble(m_s, n=n, seed=seed, samples_per_traj=512)
    else:
        m_s, nt_s = 200_000, 3
        X, DX = systems.lorenz_ensemble(m_s, seed=seed, samples_per_traj=1000)
    t0 = time.perf_counter()
    Th = oracle.evaluate_library(ex, X)
    oracle.sparse_regression(oracle.STLSQ(LAMS), Th, DX[:nt_s])
    dt = time.perf_counter() - t0
    n_t = DX.shape[0]
    scaled = dt * (n_t / nt_s)  # all targets
    desc = (f"oracle (NumPy/SciPy port of the reference algorithm: materialised Theta {k}x{m_s}, per-target pivoted-QR "
            f"STLSQ, 41 thresholds) on {m_s} samples and {nt_s} of {n_t} targets, time scaled x{n_t // nt_s} to all targets; "
            f"cost is linear in the sample count")
    return m_s / scaled, desc


def cpu_streaming_gram_sample(wl, seed=0):
    """Second CPU line: the best reasonable CPU algorithm for the same job -- streaming Gram
    (chunks of Theta through OpenBLAS dgemm) + one Cholesky-based sweep -- so that the GPU speed-up is not
    inflated by the reference's O(n_t m k^2) pivoted-QR choice.  Returns (samples/s, description)."""
    import oracle
    from ddb200 import systems

    n, deg = wl["n"], wl["degree"]
    ex = oracle.polynomial_exponents(n, deg)
    k = ex.shape[1]
    if wl["system"] == 1:
        m_s, chunk = 32768, 4096
        X, DX = systems.lorenz96_ensemble(m_s, n=n, seed=seed, samples_per_traj=512)
    else:
        m_s, chunk = 1_000_000, 100_000
        X, DX = systems.lorenz_ensemble(m_s, seed=seed, samples_per_traj=1000)
    t0 = time.perf_counter()
    G = np.zeros((k, k))
    R = np.zeros((k, DX.shape[0]))
    for s0 in range(0, m_s, chunk):
        Th = oracle.evaluate_library(ex, X[:, s0:s0 + chunk])
        G += Th @ Th.T
        R += Th @ DX[:, s0:s0 + chunk].T
    dt = time.perf_counter() - t0
    return m_s / dt, (f"NumPy/OpenBLAS streaming Gram (Theta chunks of {chunk} samples, dgemm) on {m_s} samples; the sweep on "
                      f"the {k}x{k} blocks is sample-count independent and not included")


def run_reference(args, wl, rank):
    if rank != 0:
        return
    cores = host_threads()
    for _ in range(args.warmup):
        cpu_reference_sample(wl)
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.steps):
        v, desc = cpu_reference_sample(wl)
        vals.append(v)
    dt = (time.perf_counter() - t0) / args.steps
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": wl["name"], "note": "reference CPU algorithm, bounded sample",
                                                         "host_cores": os.cpu_count(), "blas_threads": cores},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_reference_dmd(args, wl, rank):
    if rank != 0:
        return
    import oracle

    cores = host_threads()
    n, m_s = wl["n"], 20_000
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, m_s + 1))
    t0 = time.perf_counter()
    oracle.koopman_solve(oracle.DMDSVD(), x)
    dt = time.perf_counter() - t0
    v = m_s / dt
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
                      "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": wl["name"]},
                      "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                                       "sample": f"oracle DMDSVD (gesdd + geev + pivoted-QR output map) on {m_s} of 1e6 snapshots"},
                      "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Env:
    """Per-process plumbing: device, torch.distributed group, the ddb context with its in-library communicator."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import ddb200

        self.torch, self.dist, self.ddb = torch, dist, ddb200
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local_rank}"))
        self.group = dist.group.WORLD if self.world > 1 else None
        self.dev = torch.device(f"cuda:{self.local_rank}")
        torch.cuda.set_device(self.dev)
        self.ctx = ddb200.Context(self.local_rank)
        if self.world > 1:
            ddb200.dist.init_comm(self.ctx, self.group)  # ncclCommInitRank inside the library
        self.ext = torch.cuda.ExternalStream(self.ctx.stream(), device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def timed(env, fn, steps, warmup, sampler=None):
    """W untimed + K timed calls of fn, bracketed by barrier + synchronize; CUDA events on the context's stream;
    returns (ms_per_step max over ranks, per-step list of this rank)."""
    torch = env.torch
    with torch.cuda.stream(env.ext):
        for _ in range(warmup):
            fn()
        if sampler is not None:
            sampler.start()
        env.barrier()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        marks[0].record(env.ext)
        for i in range(steps):
            fn()
            marks[i + 1].record(env.ext)
        env.barrier()
        if sampler is not None:
            sampler.stop_flag = True
            sampler.join()
        total = marks[0].elapsed_time(marks[steps])
        per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
    return env.max_over_ranks(total) / steps, per_step


def run_sparse(env, wl, alg, noise, steps, warmup, want_e2e, want_clocks):
    """The hot path on one sparse-regression workload; returns the fields of a bench line."""
    ddb200, torch, ctx = env.ddb, env.torch, env.ctx
    basis = ddb200.polynomial_basis(wl["n"], wl["degree"])
    k, n = len(basis), wl["n"]
    n_t = 3 if wl["system"] == 0 else n
    m_total = wl["m"]
    lo, hi = ddb200.dist.shard_range(m_total, env.rank, env.world)
    m_local = hi - lo
    ctx.set_features(0, None)
    ctx.set_library(basis.exponents)
    ctx.generate(wl["system"], n, m_local, first_sample=lo, seed=0, noise_std=noise)  # synthetic trajectories born in HBM
    if alg == "STLSQ":
        prm = ctx.make_params("STLSQ")
        lams = LAMS
    elif alg == "ADMM":
        prm = ctx.make_params("ADMM", rho=1.0)
        lams = LAMS
    else:
        prm = ctx.make_params("SR3", rho=1.0, proximal="HardThreshold")
        lams = ddb200.SR3(LAMS, 1.0).thresholds  # SR3 + HardThreshold stores thr^2 / 2 (SR3.jl:63)
    live = {key: [] for key in ("gram_ms", "factor_ms", "sweep_ms", "rss_ms", "select_ms", "allreduce_ms", "launches", "big_steps",
                                "candidates")}
    info = {}
    last = {}

    def step():
        ctx.gram()                      # fused Theta + Gram over this rank's samples, into the context's own buffer
        t = ctx.timings()
        ctx.allreduce_gram()            # NCCL sum on the compute stream (no-op on one GPU)
        ctx.sweep(prm, lams, m_total)
        ctx.rss()
        ctx.allreduce_rss()
        last["out"] = ctx.select()
        t2 = ctx.timings()
        live["gram_ms"].append(t["gram_ms"])
        for key in ("factor_ms", "sweep_ms", "rss_ms", "select_ms", "big_steps", "candidates", "allreduce_ms"):
            live[key].append(t2[key])
        live["launches"].append(t["gram_launches"] + t2["sweep_launches"])
        info["schedule"], info["dmma_flops"] = t["gram_schedule"], t["gram_dmma_flops"]

    sampler = ClockSampler(env.local_rank) if want_clocks else None
    ms_step, per_step = timed(env, step, steps, warmup, sampler)
    out = last["out"]
    dof = out.dof.tolist()
    if alg == "STLSQ" and noise == 0.0:  # never report a number for wrong results
        expect = [4] * n_t if wl["system"] == 1 else [2, 3, 2]
        assert dof == expect, f"unexpected model: dof {dof}"
    stage = {key: float(np.mean(v[-steps:])) for key, v in live.items()}  # timed steps only (the first all-reduce sets NCCL up)
    F = gram_flops_per_sample(k, n_t)
    gram_s = stage["gram_ms"] * 1e-3
    syrk_equiv = F * m_local / gram_s * 1e-12
    executed = info["dmma_flops"] * m_local / gram_s * 1e-12
    sched = info["schedule"]
    kernels = {
        0: "Gram stage = ddb::gram_kernel (self-building: TMA-fed X/Y tiles, terms multiplied out in shared memory, FP64 DMMA.8x8x4) + reduce",
        1: "Gram stage = ddb::gram_ring_kernel (producer SMs + TMA-fed FP64 DMMA.8x8x4 consumers) + ddb::gram_kernel (diagonal pairs) + reduce",
        4: "Gram stage = ddb::gram_w4_kernel (single 64 x 64 diagonal pair, 36 live sub-tiles dealt 9 + 9 + 9 + 9 to four warps; "
           "TMA-fed X/Y tiles, terms multiplied out in shared memory, FP64 DMMA.8x8x4) + reduce",
        3: "Gram stage = ddb::gram_pow_kernel (single 64 x 64 diagonal pair, 36 live sub-tiles dealt 9 + 9 + 9 + 9 to four warps; builder reads a "
           "per-stage table of powers x_i^e: one factor per state instead of one per degree; TMA-fed X/Y tiles, FP64 DMMA.8x8x4) + reduce",
        2: "Gram stage = ddb::gram_kernel over the moment staircase (every distinct moment of the monomial library contracted once; "
           "TMA-fed X/Y tiles, virtual terms multiplied out in shared memory, FP64 DMMA.8x8x4) + ddb::moment_gather_kernel",
    }
    traffic = traffic_source = None
    if wl["system"] == 1 and wl["n"] == 64 and sched in DRAM_BYTES_PER_SAMPLE:
        per_sample, src = DRAM_BYTES_PER_SAMPLE[sched]
        traffic = per_sample * m_local
        traffic_source = (f"dram__bytes_read.sum + dram__bytes_write.sum of the Gram kernels at 1e6 samples, ncu ({src}), scaled by samples per launch; "
                          "the kernel is tensor bound: this is 3.4 % of the HBM rate, the X / Y tiles are re-read once per wave and cost class by design")
    hbm, hbm_src = hbm_peak_gbs()
    roofline = {
        "bound": "tensor", "kernel": kernels.get(sched, kernels[0]), "schedule": {0: "self-building", 1: "ring", 2: "moment", 3: "self-building (four balanced warps, power table)", 4: "self-building (four balanced warps)"}.get(sched, "?"),
        # achieved = FP64 tensor-core flops the kernel ISSUED (512 per DMMA.8x8x4, counted by the plan and confirmed by ncu's
        # DMMA instruction count) / its CUDA-event duration; frac = its share of the DMMA issue-rate peak
        "achieved": executed, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": executed / FP64_DMMA_PEAK_TFLOPS,
        "issued_flops_per_sample": info["dmma_flops"],
        # the SURVEY 8(d) algorithmic count k(k+1) + 2 k n_t + 2 n_t per sample: the moment schedule contracts every distinct
        # moment once and therefore issues FEWER flops than that -- the saving is reported on its own, not as a fraction of peak
        "algorithmic_flops_per_sample": F, "algorithmic_bytes_per_sample": 8 * (n + n_t),
        "syrk_equiv_tflops": syrk_equiv, "algorithmic_speedup": F / info["dmma_flops"] if info["dmma_flops"] else None,
        "traffic": traffic, "traffic_source": traffic_source,
        "hbm_gbs_x_stream": 8.0 * (n + n_t) * m_local / gram_s * 1e-9, "hbm_peak_gbs": hbm, "hbm_peak_source": hbm_src,
        "peak_source": "DMMA issue-rate peak measured on this pool with tools/fp64_peak.cu (profiles/fp64_peak_r1.json; reproduce: "
                       "`make -C tools fp64_peak && tools/bin/fp64_peak`); MEASURED_PEAKS.json has no FP64 figure",
        "frac_of_cublas_dgemm": executed / CUBLAS_DGEMM_TFLOPS,
    }
    res = {
        "value": m_total / (ms_step * 1e-3), "unit": "samples/s", "ms_per_step": ms_step, "ms_per_step_median": float(np.median(per_step)),
        "ms_per_step_best": float(np.min(per_step)),
        "config": {"workload": wl["name"] if alg == "STLSQ" and noise == 0.0 else f"{wl['name']} [{alg}, noise_std={noise:g}]",
                   "samples_total": m_total, "samples_per_gpu": m_local, "terms": k, "targets": n_t, "thresholds": int(LAMS.size),
                   "algorithm": alg, "noise_std": noise, "dof": dof, "parallelism": f"row-sharded x{env.world}",
                   "collectives": "ncclAllReduce inside libddb200 (compute stream)" if env.world > 1 else "none",
                   "l2": "inputs (8(n+n_t) B/sample x samples_per_gpu) are far larger than the 126 MB L2"},
        "stages_ms": {key: stage[key] for key in ("gram_ms", "allreduce_ms", "factor_ms", "sweep_ms", "rss_ms", "select_ms")},
        "gram_samples_per_s": m_total / gram_s,
        "sweep_wall_ms": stage["factor_ms"] + stage["sweep_ms"] + stage["rss_ms"] + stage["select_ms"],
        "big_steps": stage["big_steps"], "candidates": stage["candidates"],
        "fp64_tc_pct": 100.0 * executed / FP64_DMMA_PEAK_TFLOPS,
        "rss_hbm_frac": (8.0 * (n + n_t) * m_local / (stage["rss_ms"] * 1e-3) * 1e-9 / hbm) if stage["rss_ms"] > 0 else None,
        "roofline": roofline, "gpu_launches": int(np.sum(live["launches"])),
        "clocks": sampler.summary() if sampler is not None else None,
    }

    # ---- end to end through the public host-buffer call (ddb200's Context.solve_sparse = ddb_solve_sparse; on N GPUs
    # its sharded form ddb_solve_sparse_sharded): host -> H2D -> hot path -> D2H, all inside the timed region
    if want_e2e:
        res["e2e"] = run_e2e(env, prm, lams, n, n_t, k, m_local, m_total, steps)
    return res


def run_e2e(env, prm, lams, n, n_t, k, m_local, m_total, steps):
    ddb200, torch, ctx = env.ddb, env.torch, env.ctx
    h2d = (n + n_t) * m_local * 8
    d2h = (n_t * k + 5 * n_t) * 8
    n_e2e = max(1, min(steps, 3))
    out = {}
    for kind in ("pinned", "pageable"):
        if kind == "pinned":
            X_h = torch.empty((m_local, n), dtype=torch.float64).pin_memory()  # row-major (m, n) == column-major n x m
            Y_h = torch.empty((m_local, n_t), dtype=torch.float64).pin_memory()
            Xn, Yn = X_h.numpy().T, Y_h.numpy().T  # Fortran-order views: what a Julia Matrix looks like
        else:
            Xn = np.empty((n, m_local), order="F")
            Yn = np.empty((n_t, m_local), order="F")
        ddb200._lib.check(ctx._lib.ddb_download_data(ctx._ctx, Xn.ctypes.data, Yn.ctypes.data))

        def call():
            if env.world == 1:
                return ctx.solve_sparse(Xn, Yn, prm, lams)
            return ctx.solve_sparse_sharded(Xn, Yn, prm, lams, m_total)

        ms, _ = timed(env, call, n_e2e, 1)
        out[kind] = {"value": m_total / (ms * 1e-3), "ms_per_step": ms}
        del Xn, Yn
        if kind == "pinned":
            del X_h, Y_h
    return {"value": out["pinned"]["value"], "unit": "samples/s", "h2d_bytes_per_step": h2d * env.world, "d2h_bytes_per_step": d2h,
            "steps": n_e2e, "ms_per_step": out["pinned"]["ms_per_step"],
            "call": "ddb200.Context.solve_sparse -> ddb_solve_sparse" if env.world == 1
                    else "ddb200.Context.solve_sparse_sharded -> ddb_solve_sparse_sharded (one rank per GPU)",
            "host_memory": "page-locked",
            "pageable": {"value": out["pageable"]["value"], "ms_per_step": out["pageable"]["ms_per_step"],
                         "note": "same call from ordinary (pageable) host arrays, as a Julia Matrix is"}}


def run_dmd(env, wl, steps, warmup, want_clocks):
    """Config C5: P = X X', Q = Y X' (row-sharded, in-library all-reduce of [P | Q]) and K = Q P^-1."""
    ddb200, torch, ctx = env.ddb, env.torch, env.ctx
    n, m_total = wl["n"], wl["m"]
    lo, hi = ddb200.dist.shard_range(m_total, env.rank, env.world)
    m_local = hi - lo
    dev = env.dev
    xt = torch.empty((m_local + 1, n), dtype=torch.float64, device=dev)  # one-snapshot halo: Y is X shifted by one
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + env.rank)
    for s0 in range(0, m_local + 1, 65536):
        xt[s0:s0 + 65536].normal_(generator=g)
    PQ = torch.zeros((2, n, n), dtype=torch.float64, device=dev)
    K = torch.zeros((n, n), dtype=torch.float64, device=dev)
    gram_ms, op_ms = [], []

    def step():
        ctx.dmd_gram(xt.data_ptr(), xt.data_ptr() + 8 * n, n, m_local, PQ[0].data_ptr(), PQ[1].data_ptr())
        gram_ms.append(ctx.timings()["gram_ms"])
        ctx.allreduce_sum(PQ.data_ptr(), PQ.numel())
        ctx.dmd_operator(PQ[0].data_ptr(), PQ[1].data_ptr(), n, K.data_ptr())
        op_ms.append(ctx.timings()["factor_ms"])

    sampler = ClockSampler(env.local_rank) if want_clocks else None
    for _ in range(warmup):
        step()
    gram_ms.clear(), op_ms.clear()
    ms_step, per_step = timed(env, step, steps, 0, sampler)
    F = n * (n + 1) + 2 * n * n  # P symmetric (SYRK count) + Q full, per snapshot
    achieved = F * m_local / (float(np.mean(gram_ms)) * 1e-3) * 1e-12
    res = {
        "value": m_total / (ms_step * 1e-3), "unit": "samples/s", "ms_per_step": ms_step, "ms_per_step_median": float(np.median(per_step)),
        "config": {"workload": wl["name"], "snapshots_total": m_total, "snapshots_per_gpu": m_local, "states": n,
                   "parallelism": f"row-sharded x{env.world}", "l2": "inputs (32.8 GB / ranks) far exceed the 126 MB L2"},
        "stages_ms": {"gram_ms": float(np.mean(gram_ms)), "operator_ms": float(np.mean(op_ms))},
        "fp64_tc_pct": 100.0 * achieved / FP64_DMMA_PEAK_TFLOPS,
        "roofline": {"bound": "tensor", "kernel": "ddb::dmd_gram_kernel (FP64 DMMA.8x8x4, TMA-fed)", "achieved": achieved,
                     "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_DMMA_PEAK_TFLOPS, "traffic": None,
                     "algorithmic_flops_per_sample": F, "algorithmic_bytes_per_sample": 8 * n,
                     "note": "the identity library has no repeated moments: issued = algorithmic flops"},
        "gpu_launches": 2 * steps, "clocks": sampler.summary() if sampler is not None else None,
    }
    del xt, PQ, K
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--alg", default="STLSQ", choices=["STLSQ", "ADMM", "SR3"],
                    help="sweep algorithm; ADMM (rho=1) / SR3 (nu=1, HardThreshold) at workload c3 = config C4")
    ap.add_argument("--noise", type=float, default=0.0, help="std of N(0, s^2) added to the derivatives (SURVEY 8d stress variant: 1e-3)")
    ap.add_argument("--samples", type=int, default=0, help="override the total sample count (debugging)")
    ap.add_argument("--subs", default="auto", choices=["auto", "all", "none"],
                    help="other BASELINE configs under the 'sub' key: auto = C2, C4, noisy C3 on one GPU and C5 at every N")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.samples:
        wl["m"] = args.samples
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        (run_reference_dmd if args.workload == "c5" else run_reference)(args, wl, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    env = Env()
    if args.workload == "c5":
        main_res = run_dmd(env, wl, args.steps, args.warmup, True)
        main_res.update({"cpu_baseline": None, "e2e": None})
    else:
        main_res = run_sparse(env, wl, args.alg, args.noise, args.steps, args.warmup, not args.no_e2e, True)
    subs = {}
    headline = args.workload == "c3" and args.alg == "STLSQ" and args.noise == 0.0 and not args.samples
    if args.subs != "none" and headline:
        one = env.world == 1 or args.subs == "all"
        s_steps, s_warm = 2, 3
        if one:
            subs["c2"] = run_sparse(env, dict(WORKLOADS["c2"]), "STLSQ", 0.0, s_steps, s_warm, False, False)
            subs["c4_admm"] = run_sparse(env, dict(WORKLOADS["c3"]), "ADMM", 0.0, s_steps, s_warm, False, False)
            subs["c4_sr3"] = run_sparse(env, dict(WORKLOADS["c3"]), "SR3", 0.0, s_steps, s_warm, False, False)
            # SURVEY 8(d) stress variant (N(0, 1e-3^2) on the derivatives): with 1e7 samples the coefficient noise is
            # ~1e-7, far below the smallest threshold 1e-5, so the active sets still collapse at once (big_steps = 0);
            # c3_stress raises the noise to 1.0 so that most of the 2145 terms survive the first thresholds and the
            # sweep runs batched 64 x ~2145^2 masked factorisations (the blocked DMMA Cholesky)
            subs["c3_noisy"] = run_sparse(env, dict(WORKLOADS["c3"]), "STLSQ", 1e-3, s_steps, s_warm, False, False)
            stress = dict(WORKLOADS["c3"])
            stress["m"] = 1_000_000  # the exact-rss pass costs m x (sum of the candidates' term counts): dense candidates
            stress["name"] = "Lorenz-96 n=64, degree-2 library (2145 terms), 1e6 samples, derivative noise 1.0 (large active sets)"
            subs["c3_stress"] = run_sparse(env, stress, "STLSQ", 1.0, 1, 1, False, False)
        env.ctx.close()                       # release the sparse workloads' buffers before the 32.8 GB snapshot matrix
        env.ctx = env.ddb.Context(env.local_rank)
        if env.world > 1:
            env.ddb.dist.init_comm(env.ctx, env.group)
        env.ext = env.torch.cuda.ExternalStream(env.ctx.stream(), device=env.dev)
        subs["c5"] = run_dmd(env, dict(WORKLOADS["c5"]), s_steps, s_warm, False)

    if env.rank == 0:
        cpu = None
        if not args.no_cpu_baseline and env.world == 1 and args.workload != "c5":
            cores = host_threads()
            v, desc = cpu_reference_sample(wl)
            cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": desc}
            v2, desc2 = cpu_streaming_gram_sample(wl)
            cpu["best_cpu_algorithm"] = {"value": v2, "unit": "samples/s", "cores": cores, "sample": desc2}
        line = {"metric": METRIC, "value": main_res.pop("value"), "unit": main_res.pop("unit"), "n_gpus": env.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_res.pop("ms_per_step"), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        line.update(main_res)
        line.setdefault("cpu_baseline", cpu)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        line.setdefault("e2e", None)
        if subs:
            line["sub"] = subs
        print(json.dumps(line), flush=True)
    env.close()


if __name__ == "__main__":
    main()
