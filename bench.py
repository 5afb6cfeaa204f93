"""Benchmark of the sparse-identification hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5] [--alg STLSQ|ADMM|SR3] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input: fused Theta+Gram over this
rank's samples -> sum all-reduce of the packed [G|R|yy] -> 41-threshold STLSQ sweep -> exact-rss
streaming pass -> all-reduce of the candidate-rss table -> selector replay.  Workload c3 (default, the
configuration the north-star target is quoted on; it fits one GPU: 10.24 GB): Lorenz-96, n = 64, degree-2
polynomial library (2145 terms), 10^7 samples in total, row-sharded over the ranks (strong scaling).
Workload c2: Lorenz, degree 5 (56 terms), 10^7 samples.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAMS = 10.0 ** np.arange(-5, -1 + 1e-9, 0.1)  # exp10.(-5:0.1:-1), 41 thresholds (README.md:55)
METRIC = "samples/sec through fused Theta+Gram; STLSQ lambda-sweep wall-time; % FP64 TC peak"
# FP64 tensor (DMMA) peak of this pool's B200s: not in MEASURED_PEAKS.json (bf16 only), measured with
# tools/fp64_peak.cu (profiles/fp64_peak_r1.json): DMMA issue-rate peak 37.0 TF, cublasDgemm 8192^3 35.9 TF.
FP64_DMMA_PEAK_TFLOPS = 37.0
CUBLAS_DGEMM_TFLOPS = 35.9
# DRAM read traffic of the two Gram kernels per sample (ncu dram__bytes_read.sum at 1e6 samples,
# profiles/gram_traffic_r1g_c3_1e6.csv): gram_ring_kernel 1041 B (the algorithmic 1024 B + ring spill) and the
# diagonal-pair pass of gram_kernel 1886 B (its 9 sample ranges x 16 pairs share tiles through L2)
C3_RING_DRAM_BYTES_PER_SAMPLE = (1041148928 + 1886253824) / 1e6
# moment schedule (default at C3): gram_kernel over 78 block pairs in four lock-step waves (37 whole pairs x 4 sample
# ranges, 21 x 7, 2 x 74, then 2 whole x 15, 6 half x 8 and 10 quarter x 7 ranges: about six passes over the X / Y
# stream): dram__bytes_read.sum 6.9288 GB + dram__bytes_write.sum 0.0696 GB at 1e6 samples
# (profiles/gram_traffic_r1o_c3_1e6.csv; a two-wave plan reads 4.43 GB and is 2.2 % slower, profiles/gram_traffic_r1n_c3_1e6.csv)
C3_MOMENT_DRAM_BYTES_PER_SAMPLE = (6928822016 + 69605632) / 1e6

WORKLOADS = {
    "c3": dict(system=1, n=64, degree=2, m=10_000_000, name="Lorenz-96 n=64, degree-2 library (2145 terms), 1e7 samples, 41-lambda STLSQ"),
    "c2": dict(system=0, n=3, degree=5, m=10_000_000, name="Lorenz, degree-5 library (56 terms), 1e7 samples, 41-lambda STLSQ"),
    "c5": dict(system=-1, n=4096, degree=1, m=1_000_000, name="exact-DMD Gram path: 4096 states x 1e6 snapshots, P = XX', Q = YX', K = Q P^-1"),
}


def gram_flops_per_sample(k, n_t):
    """Algorithmic work of the Gram stage per sample: symmetric rank-1 update of G counted
    once (SYRK count) + Theta Y' + diag(Y Y'); monomial construction is not counted."""
    return k * (k + 1) + 2 * k * n_t + 2 * n_t


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
    cores = os.cpu_count() or 1
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
        "dtype": "f64", "data": "synthetic", "config": {"workload": wl["name"], "note": "reference CPU algorithm, bounded sample"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# config C5: DataDrivenDMD Gram path (not the default bench line; `--workload c5`)
# ------------------------------------------------------------------------------------------------
def run_dmd(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import ddb200

    n, m_total = wl["n"], wl["m"]
    if args.impl == "reference":
        if rank == 0:  # the reference path: thin SVD of X (gesdd) + eigen, on a bounded sample
            import oracle

            m_s = 20_000
            rng = np.random.default_rng(0)
            x = rng.standard_normal((n, m_s + 1))
            t0 = time.perf_counter()
            oracle.koopman_solve(oracle.DMDSVD(), x)
            dt = time.perf_counter() - t0
            v = m_s / dt
            print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
                              "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
                              "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": wl["name"]},
                              "cpu_baseline": {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                               "sample": f"oracle DMDSVD (gesdd + geev + pivoted-QR output map) on {m_s} of 1e6 snapshots"},
                              "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    lo, hi = ddb200.dist.shard_range(m_total, rank, world)
    m_local = hi - lo
    xt = torch.empty((m_local + 1, n), dtype=torch.float64, device=dev)  # one-snapshot halo: Y is X shifted by one
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    for s0 in range(0, m_local + 1, 65536):
        xt[s0:s0 + 65536].normal_(generator=g)
    PQ = torch.zeros((2, n, n), dtype=torch.float64, device=dev)
    K = torch.zeros((n, n), dtype=torch.float64, device=dev)
    ctx = ddb200.Context(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    gram_ms, op_ms = [], []

    def step():
        ctx.dmd_gram(xt.data_ptr(), xt.data_ptr() + 8 * n, n, m_local, PQ[0].data_ptr(), PQ[1].data_ptr())
        gram_ms.append(ctx.timings()["gram_ms"])
        if world > 1:
            ddb200.dist.allreduce_sum(PQ, dist.group.WORLD)
        ctx.dmd_operator(PQ[0].data_ptr(), PQ[1].data_ptr(), n, K.data_ptr())
        op_ms.append(ctx.timings()["factor_ms"])

    with torch.cuda.stream(ext):
        for _ in range(max(3, args.warmup)):
            step()
        gram_ms.clear(), op_ms.clear()
        sampler = ClockSampler(local_rank)
        sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(args.steps):
            step()
        e1.record(ext)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        sampler.stop_flag = True
        sampler.join()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    F = n * (n + 1) + 2 * n * n  # P symmetric (SYRK count) + Q full, per snapshot
    achieved = F * m_local / (float(np.mean(gram_ms)) * 1e-3) * 1e-12
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": m_total / (ms_step * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "snapshots_total": m_total, "snapshots_per_gpu": m_local, "states": n,
                       "l2": "inputs (32.8 GB / ranks) far exceed the 126 MB L2"},
            "stages_ms": {"gram_ms": float(np.mean(gram_ms)), "operator_ms": float(np.mean(op_ms))},
            "fp64_tc_pct": 100.0 * achieved / FP64_DMMA_PEAK_TFLOPS,
            "roofline": {"bound": "tensor", "kernel": "ddb::dmd_gram_kernel (FP64 DMMA.8x8x4, TMA-fed)", "achieved": achieved,
                         "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_DMMA_PEAK_TFLOPS, "traffic": None,
                         "algorithmic_flops_per_sample": F, "algorithmic_bytes_per_sample": 8 * n},
            "cpu_baseline": None, "e2e": None, "gpu_launches": 2 * args.steps, "clocks": sampler.summary()}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--alg", default="STLSQ", choices=["STLSQ", "ADMM", "SR3"],
                    help="sweep algorithm; ADMM (rho=1) / SR3 (nu=1, HardThreshold) at workload c3 = config C4")
    ap.add_argument("--samples", type=int, default=0, help="override the total sample count (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.samples:
        wl["m"] = args.samples
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "c5":
        run_dmd(args, wl, rank, world, local_rank)
        return
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist

    import ddb200

    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    group = dist.group.WORLD if world > 1 else None
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)

    basis = ddb200.polynomial_basis(wl["n"], wl["degree"])
    k, n, n_t = len(basis), wl["n"], wl["n"]
    m_total = wl["m"]
    lo, hi = ddb200.dist.shard_range(m_total, rank, world)
    m_local = hi - lo
    ctx = ddb200.Context(local_rank)
    ctx.set_library(basis.exponents)
    ctx.generate(wl["system"], n, m_local, first_sample=lo, seed=0)  # synthetic trajectories born in HBM
    if args.alg == "STLSQ":
        prm = ctx.make_params("STLSQ")
    elif args.alg == "ADMM":
        prm = ctx.make_params("ADMM", rho=1.0)
    else:
        prm = ctx.make_params("SR3", rho=1.0, proximal="HardThreshold")
    ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    packed = torch.empty(ctx.gram_count(), dtype=torch.float64, device=dev)
    stage_live = {"gram_ms": [], "factor_ms": [], "sweep_ms": [], "rss_ms": [], "select_ms": [], "launches": []}
    gram_info = {"schedule": 0, "dmma_flops": 0.0}

    def step(host=None):
        """One pass of the hot path; inputs device-resident, or (host = pinned X, Y) uploaded inside the step
        with the copies overlapped by the Gram kernels (ddb_upload_gram, the call ddb_solve_sparse makes)."""
        if host is None:
            ctx.gram(packed.data_ptr())
        else:
            ctx.upload_gram_raw(host[0].data_ptr(), host[1].data_ptr(), n_t, m_local, packed.data_ptr())
        t = ctx.timings()
        if world > 1:
            ddb200.dist.allreduce_sum(packed, group)
        base = packed.data_ptr()
        ctx.gram_upload_ptr(base, base + 8 * k * k, base + 8 * (k * k + k * n_t), n_t)
        ctx.sweep(prm, LAMS, m_total)
        rss = torch.empty(max(1, ctx.rss_count()), dtype=torch.float64, device=dev)
        ctx.rss(rss.data_ptr())
        if world > 1:
            ddb200.dist.allreduce_sum(rss, group)
        out = ctx.select()
        t2 = ctx.timings()
        stage_live["gram_ms"].append(t["gram_ms"])
        for key in ("factor_ms", "sweep_ms", "rss_ms", "select_ms"):
            stage_live[key].append(t2[key])
        stage_live["launches"].append(t["gram_launches"] + t2["sweep_launches"])
        gram_info["schedule"], gram_info["dmma_flops"] = t["gram_schedule"], t["gram_dmma_flops"]
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(ext):
        for _ in range(args.warmup):
            out = step()
        for v in stage_live.values():
            v.clear()
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]  # per-step times (median / best)
        e0.record(ext)
        marks[0].record(ext)
        for i in range(args.steps):
            out = step()
            marks[i + 1].record(ext)
        e1.record(ext)
        barrier()
        sampler.stop_flag = True
        sampler.join()
        ms_total = e0.elapsed_time(e1)
        per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = m_total / (ms_step * 1e-3)

    # ---- sanity: the timed work produced the right model (never report a number for wrong results)
    dof = out.dof.tolist()
    expect_dof = [4] * n_t if wl["system"] == 1 else [2, 3, 2]
    if args.alg == "STLSQ":
        assert dof == expect_dof, f"unexpected model: dof {dof}"

    stage = {kk: list(v) for kk, v in stage_live.items()}  # the timed steps only (the e2e steps below call step() too)

    # ---- end to end through the public host-buffer call: pinned host -> H2D -> hot path -> D2H
    e2e = None
    if not args.no_e2e:
        X_h = torch.empty((m_local, n), dtype=torch.float64).pin_memory()  # row-major (m, n) == column-major n x m
        Y_h = torch.empty((m_local, n_t), dtype=torch.float64).pin_memory()
        ddb200._lib.check(ctx._lib.ddb_download_data(ctx._ctx, X_h.data_ptr(), Y_h.data_ptr()))
        h2d = (X_h.numel() + Y_h.numel()) * 8
        d2h = (n_t * k + 4 * n_t) * 8

        def e2e_step():
            return step(host=(X_h, Y_h))

        with torch.cuda.stream(ext):
            e2e_step()
            barrier()
            n_e2e = max(1, min(args.steps, 3))
            e0.record(ext)
            for _ in range(n_e2e):
                e2e_step()
            e1.record(ext)
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": m_total / (float(t.item()) / n_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h, "steps": n_e2e}
        del X_h, Y_h

    # ---- roofline of the dominant kernel (gram_kernel): algorithmic flops / CUDA-event duration
    F = gram_flops_per_sample(k, n_t)
    gram_ms = float(np.mean(stage["gram_ms"]))
    achieved = F * m_local / (gram_ms * 1e-3) * 1e-12
    executed = gram_info["dmma_flops"] * m_local / (gram_ms * 1e-3) * 1e-12
    sched = gram_info["schedule"]
    kernels = {
        0: "Gram stage = ddb::gram_kernel (self-building: TMA-fed X/Y tiles, terms multiplied out in shared memory, FP64 DMMA.8x8x4) + reduce",
        1: "Gram stage = ddb::gram_ring_kernel (producer SMs + TMA-fed FP64 DMMA.8x8x4 consumers) + ddb::gram_kernel (diagonal pairs) + reduce",
        2: "Gram stage = ddb::gram_kernel over the moment staircase (every distinct moment of the monomial library contracted once; "
           "TMA-fed X/Y tiles, virtual terms multiplied out in shared memory, FP64 DMMA.8x8x4) + ddb::moment_gather_kernel",
    }
    traffic, traffic_source = None, None
    if args.workload == "c3" and sched == 2:
        traffic = C3_MOMENT_DRAM_BYTES_PER_SAMPLE * m_local
        traffic_source = ("dram__bytes_read.sum + dram__bytes_write.sum of gram_kernel at 1e6 samples, ncu "
                          "(profiles/gram_traffic_r1o_c3_1e6.csv), scaled by samples per launch")
    elif args.workload == "c3" and sched == 1:
        traffic = C3_RING_DRAM_BYTES_PER_SAMPLE * m_local
        traffic_source = ("dram__bytes_read.sum of gram_ring_kernel (1.041 GB) + gram_kernel diagonal-pair pass (1.886 GB) per 1e6 "
                          "samples, ncu (profiles/gram_traffic_r1g_c3_1e6.csv), scaled by samples per launch")
    roofline = {"bound": "tensor", "kernel": kernels.get(sched, kernels[0]),
                "schedule": {0: "self-building", 1: "ring", 2: "moment"}.get(sched, "?"),
                "achieved": achieved, "peak": FP64_DMMA_PEAK_TFLOPS,
                "unit": "TFLOP/s", "frac": achieved / FP64_DMMA_PEAK_TFLOPS,
                # `achieved` credits the ALGORITHMIC flops of SURVEY 8(d) (k(k+1) + 2 k n_t + 2 n_t per sample); the moment
                # schedule issues fewer (it skips repeated moments), so `frac` can exceed 1 -- `executed` is what the
                # tensor pipe really issued (512 flops per DMMA.8x8x4) and `executed_frac` its share of the DMMA peak
                "executed": executed, "executed_frac": executed / FP64_DMMA_PEAK_TFLOPS,
                "executed_flops_per_sample": gram_info["dmma_flops"],
                "traffic": traffic, "traffic_source": traffic_source,
                "peak_source": "measured on this pool with tools/fp64_peak.cu (DMMA issue-rate; profiles/fp64_peak_r1.json); "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "frac_of_cublas_dgemm": achieved / CUBLAS_DGEMM_TFLOPS, "algorithmic_flops_per_sample": F,
                "algorithmic_bytes_per_sample": 8 * (n + n_t)}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            v, desc = cpu_reference_sample(wl)
            cpu = {"value": v, "unit": "samples/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": desc}
            v2, desc2 = cpu_streaming_gram_sample(wl)
            cpu["best_cpu_algorithm"] = {"value": v2, "unit": "samples/s", "cores": os.cpu_count() or 1, "sample": desc2}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "ms_per_step_median": float(np.median(per_step)), "ms_per_step_best": float(np.min(per_step)),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl["name"], "samples_total": m_total, "samples_per_gpu": m_local, "terms": k, "targets": n_t,
                       "thresholds": int(LAMS.size), "algorithm": args.alg, "dof": dof,
                       "parallelism": f"row-sharded x{world}",
                       "l2": "inputs (8(n+n_t) B/sample x samples_per_gpu) are far larger than the 126 MB L2"},
            "stages_ms": {kk: float(np.mean(v)) for kk, v in stage.items() if kk != "launches"},
            "gram_samples_per_s": m_total / (gram_ms * 1e-3), "sweep_wall_ms": float(np.mean(stage["factor_ms"]) + np.mean(stage["sweep_ms"])
                                                                                  + np.mean(stage["rss_ms"]) + np.mean(stage["select_ms"])),
            # algorithmic (SYRK-count) flops over the DMMA peak; above 100 when the moment schedule skips repeated moments
            "fp64_tc_pct": 100.0 * achieved / FP64_DMMA_PEAK_TFLOPS,
            # DMMA flops the tensor pipe actually issued over the DMMA peak
            "fp64_tc_pct_executed": 100.0 * executed / FP64_DMMA_PEAK_TFLOPS,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(np.sum(stage["launches"])),
            "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
