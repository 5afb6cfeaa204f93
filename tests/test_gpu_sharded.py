"""Parity of the row-sharded path on hardware: N shards of the samples, the two sum exchanges,
the replicated sweep -- against the single-GPU result on the same data.

* ``test_two_shards_on_one_gpu``: two contexts on device 0, each with half of the samples; the two exchange steps
  are done by hand on device tensors.  Runs on ONE GPU and checks what the exchanges rely on: entry c of the
  candidate-rss table is the same model in both contexts (canonical candidate numbering, ``canon_base_kernel``).
* ``test_group_two_gpus`` / ``test_torchrun_two_ranks`` (skipped below 2 GPUs): the NCCL all-reduces INSIDE the
  library -- one process driving two GPUs (``ddb_group_solve_sparse``) and one process per GPU
  (``ddb_comm_init_rank`` + ``ddb_solve_sparse_sharded``)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import ddb200
from ddb200 import systems

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch

    return torch.cuda.device_count()


def _problem(n=12, m=6000):
    X, D = systems.lorenz96_ensemble(m, n=n, seed=11, samples_per_traj=500)
    D = D + 1e-2 * np.random.default_rng(2).standard_normal(D.shape)
    return X, D, ddb200.polynomial_basis(n, 2), 10.0 ** np.arange(-3, -0.4, 0.25)


def _params(ctx, alg):
    return {"STLSQ": lambda: ctx.make_params("STLSQ"), "ADMM": lambda: ctx.make_params("ADMM", rho=1.0),
            "SR3": lambda: ctx.make_params("SR3", rho=1.0, proximal="HardThreshold")}[alg]()


def _thresholds(alg, lams):
    return ddb200.SR3(lams, 1.0).thresholds if alg == "SR3" else lams


def _assert_same(out, ref, alg="STLSQ"):
    """Sharded against single-GPU: supports equal, coefficients <= 1e-12 relative.  SR3: iterates of neighbouring
    thresholds are numerically tied (their rss differs by ~1e-16 relative), so the last-bit
    differences of the summed blocks may move `lambda*` to an equivalent iterate -- coefficients <= 1e-7 there."""
    np.testing.assert_array_equal(out.coefficients != 0, ref.coefficients != 0)
    nz = ref.coefficients != 0
    rtol = 1e-7 if alg == "SR3" else 1e-12
    assert (np.abs(out.coefficients[nz] - ref.coefficients[nz]) / np.abs(ref.coefficients[nz])).max() <= rtol
    if alg != "SR3":
        np.testing.assert_allclose(out.lambda_opt, ref.lambda_opt, rtol=1e-15)
    np.testing.assert_array_equal(out.dof, ref.dof)
    np.testing.assert_array_equal(out.iterations, ref.iterations)
    np.testing.assert_allclose(out.rss, ref.rss, rtol=1e-9)


@pytest.mark.parametrize("alg", ["STLSQ", "ADMM", "SR3"])
def test_two_shards_on_one_gpu(alg):
    import torch

    X, D, basis, lams = _problem()
    m = X.shape[1]
    thr = _thresholds(alg, lams)
    dev = torch.device("cuda:0")
    with ddb200.Context(0) as one:
        one.set_library(basis.exponents)
        one.upload(X, D)
        one.gram()
        prm = _params(one, alg)
        one.sweep(prm, thr)
        one.rss()
        ref = one.select(paths=True)
    cuts = [ddb200.dist.shard_range(m, r, 2) for r in range(2)]
    ctxs = [ddb200.Context(0) for _ in cuts]
    try:
        packed = []
        for c, (lo, hi) in zip(ctxs, cuts):
            c.set_library(basis.exponents)
            c.upload(X[:, lo:hi], D[:, lo:hi])
            t = torch.empty(c.gram_count(), dtype=torch.float64, device=dev)
            c.gram(t.data_ptr())
            c.synchronize()
            packed.append(t)
        total = packed[0] + packed[1]  # exchange 1: sum of the packed [G | R | yy]
        torch.cuda.synchronize()
        k, n_t = ctxs[0].k, ctxs[0].n_t
        tables = []
        for c in ctxs:
            b = total.data_ptr()
            c.gram_upload_ptr(b, b + 8 * k * k, b + 8 * (k * k + k * n_t), n_t)
            c.sweep(prm, thr, m)
        assert ctxs[0].rss_count() == ctxs[1].rss_count()
        for c in ctxs:
            t = torch.empty(c.rss_count(), dtype=torch.float64, device=dev)
            c.rss(t.data_ptr())
            c.synchronize()
            tables.append(t)
        rss_total = tables[0] + tables[1]  # exchange 2: sum of the candidate-rss tables, entry by entry
        for t in tables:
            t.copy_(rss_total)  # in place, as an all-reduce leaves it: ddb_select reads the table ddb_rss was given
        torch.cuda.synchronize()
        outs = [c.select(paths=True) for c in ctxs]
        for out in outs:
            _assert_same(out, ref, alg)
            np.testing.assert_array_equal(out.support_path, ref.support_path)
            np.testing.assert_array_equal(out.iters_path, ref.iters_path)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("alg", ["STLSQ", "ADMM", "SR3"])
def test_group_two_gpus(alg):
    """One process, two GPUs, NCCL inside the library: ``ddb_group_solve_sparse`` against one GPU."""
    X, D, basis, lams = _problem()
    thr = _thresholds(alg, lams)
    with ddb200.Context(0) as one:
        one.set_library(basis.exponents)
        prm = _params(one, alg)
        ref = one.solve_sparse(X, D, prm, thr)
    with ddb200.Group(2) as grp:
        grp.set_library(basis.exponents)
        out = grp.solve_sparse(X, D, prm, thr)
        t = grp.timings()
        out2 = grp.solve_sparse(X, D, prm, thr)  # and again on the same group (plans / buffers reused)
    _assert_same(out, ref, alg)
    _assert_same(out2, ref, alg)
    np.testing.assert_array_equal(out.coefficients, out2.coefficients)  # bit-reproducible from call to call
    assert t["allreduce_calls"] == 2 and t["allreduce_bytes"] > 8 * ddb200.polynomial_basis(12, 2).exponents.shape[1] ** 2
    # through the host mirror: B200Sparse(inner; ngpus = 2)
    inner = {"STLSQ": lambda: ddb200.STLSQ(lams), "ADMM": lambda: ddb200.ADMM(lams, 1.0), "SR3": lambda: ddb200.SR3(lams, 1.0)}[alg]()
    sol = ddb200.solve(ddb200.ContinuousDataDrivenProblem(X, DX=D), basis, ddb200.B200Sparse(inner, ngpus=2))
    np.testing.assert_array_equal(sol.results[0].coefficients, out.coefficients)


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
def test_group_two_gpus_parked_targets():
    """Noisy data on a 190-term library: active sets beyond the shared-memory solver park their targets, and in a
    sharded run the parked targets are dealt to the ranks (each factorises its share, one all-reduce of the solutions;
    ``stlsq_big_gather_kernel``).  Two GPUs against one: same supports and iteration counts, coefficients to the
    accuracy the differently-summed blocks allow."""
    n, m = 18, 5000
    X, D = systems.lorenz96_ensemble(m, n=n, seed=5, samples_per_traj=500)
    D = D + 0.5 * np.random.default_rng(4).standard_normal(D.shape)
    basis = ddb200.polynomial_basis(n, 2)
    lams = 10.0 ** np.arange(-4, -0.4, 0.25)
    with ddb200.Context(0) as one:
        one.set_library(basis.exponents)
        prm = one.make_params("STLSQ")
        ref = one.solve_sparse(X, D, prm, lams)
        assert one.timings()["big_steps"] > 0
    with ddb200.Group(2) as grp:
        grp.set_library(basis.exponents)
        out = grp.solve_sparse(X, D, prm, lams)
        t = grp.timings()
        out2 = grp.solve_sparse(X, D, prm, lams)
    assert t["big_steps"] > 0 and t["allreduce_calls"] > 2  # Gram, rss and one exchange per big step
    np.testing.assert_array_equal(out.coefficients != 0, ref.coefficients != 0)
    np.testing.assert_array_equal(out.iterations, ref.iterations)
    np.testing.assert_array_equal(out.dof, ref.dof)
    np.testing.assert_allclose(out.coefficients, ref.coefficients, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(out.lambda_opt, ref.lambda_opt, rtol=1e-15)
    np.testing.assert_allclose(out.rss, ref.rss, rtol=1e-9)
    np.testing.assert_array_equal(out.coefficients, out2.coefficients)  # bit-reproducible from call to call


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
def test_group_two_gpus_rank_deficient():
    """231 terms over 100 samples: the initial least squares of both targets is a parked big step whose factorisation
    breaks down; in a sharded run each rank factorises one of them, the failure flags travel with the solutions through
    the all-reduce, and every rank then takes the same minimum-norm steps (``pinv_big_*``).  Two GPUs against one."""
    rng = np.random.default_rng(60)
    X = rng.standard_normal((20, 100))
    Y = np.stack([X[0] * X[1] - 0.5 * X[2], 2.0 * X[3] + X[4] * X[4]])
    basis = ddb200.polynomial_basis(20, 2)
    lams = np.array([0.05, 0.1, 0.3])
    with ddb200.Context(0) as one:
        one.set_library(basis.exponents)
        prm = one.make_params("STLSQ")
        ref = one.solve_sparse(X, Y, prm, lams)
        assert one.timings()["big_steps"] > 0
    with ddb200.Group(2) as grp:
        grp.set_library(basis.exponents)
        out = grp.solve_sparse(X, Y, prm, lams)
    assert (ref.coefficients != 0).any()
    np.testing.assert_array_equal(out.coefficients != 0, ref.coefficients != 0)
    np.testing.assert_array_equal(out.iterations, ref.iterations)
    np.testing.assert_allclose(out.coefficients, ref.coefficients, rtol=1e-7, atol=1e-10)
    assert not (out.flags & 1).any() and (out.flags & 2).all()


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("alg", ["STLSQ", "ADMM"])
def test_torchrun_two_ranks(alg, tmp_path):
    """One process per GPU (the way bench.py is launched): communicator from ``ddb_comm_init_rank``."""
    out = tmp_path / "res.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "_sharded_worker.py"), str(out), alg]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    z = np.load(out)
    np.testing.assert_array_equal(z["coef"] != 0, z["coef1"] != 0)
    nz = z["coef1"] != 0
    assert (np.abs(z["coef"][nz] - z["coef1"][nz]) / np.abs(z["coef1"][nz])).max() <= 1e-12
    np.testing.assert_allclose(z["lam"], z["lam1"], rtol=1e-15)
    np.testing.assert_array_equal(z["iters"], z["iters1"])
    np.testing.assert_allclose(z["train"], z["train1"], rtol=1e-9)
    assert int(z["allreduce_calls"]) == 2 and int(z["nobs"]) == 12 * 6000


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
def test_torchrun_two_ranks_dmd(tmp_path):
    """Sharded exact DMD: [P | Q] summed by the in-library all-reduce; rss and nobs are those of ALL snapshots."""
    out = tmp_path / "dmd.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29612", os.path.join(ROOT, "tests", "_sharded_worker.py"), str(out), "DMD"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    z = np.load(out)
    sc = np.sqrt(np.outer(np.diag(z["P1"]), np.diag(z["P1"])))
    assert np.max(np.abs(z["P"] - z["P1"]) / sc) < 1e-13
    assert np.max(np.abs(np.real(z["K"]) - np.real(z["K1"]))) <= 1e-9
    assert int(z["nobs"]) == int(z["nobs1"]) == 16 * 4000
    # the statistic covers ALL snapshots (it used to be one rank's part, about half)
    assert float(z["rss1"]) > 1e-3
    np.testing.assert_allclose(float(z["rss"]), float(z["rss1"]), rtol=1e-8)


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["PROCESSED", "PROCESSED_UNIT"])
def test_torchrun_two_ranks_processing_options(tmp_path, mode):
    """ZScore / UnitRange normalisation + train/test split + shuffled mini-batches on the sharded path
    (``src/commonsolve.jl:92-139`` options on row shards): same batches, same ranking by test error, same models as one
    GPU.  (UnitRange: the extrema of all samples are gathered through the sum all-reduce, one slot per rank.)"""
    out = tmp_path / "proc.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613" if mode == "PROCESSED" else "29614", os.path.join(ROOT, "tests", "_sharded_worker.py"), str(out), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    z = np.load(out)
    assert int(z["nres"]) == int(z["nres1"]) == 4  # 1521 training samples: 3 x 400 + 321
    np.testing.assert_array_equal(z["coef"] != 0, z["coef1"] != 0)
    # noisy degree-3 Lorenz library, ZScore scales: cond(G) ~ 1e10, so the last-bit differences of the blocks (summed in
    # another order over the shards) show up at ~1e-6 relative in the coefficients; supports and ranking are identical
    np.testing.assert_allclose(z["coef"], z["coef1"], rtol=5e-5, atol=1e-9)
    np.testing.assert_allclose(z["test"], z["test1"], rtol=1e-6)
    np.testing.assert_allclose(z["train"], z["train1"], rtol=1e-6)
    np.testing.assert_allclose(z["final"], z["final1"], rtol=5e-5, atol=2e-9)
    np.testing.assert_allclose(float(z["rss"]), float(z["rss1"]), rtol=1e-6)
    np.testing.assert_allclose(float(z["r2"]), float(z["r21"]), rtol=1e-9)
    assert int(z["nobs"]) == int(z["nobs1"]) == 3 * 1901
