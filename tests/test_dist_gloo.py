"""world_size-2 gloo test (CPU) of the row-sharded path's host logic: sharding, the two all-reduces and
the replicated sweep.  The GPU context is replaced by a test double that computes its partial blocks
and candidate scores with the oracle on CPU tensors; the control flow under test (``ddb200.dist``) is
the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ddb200
import oracle
from ddb200 import systems


class FakeContext:
    """Duck-typed ``Context`` working on host memory: the oracle stands in for the kernels."""

    device = 0

    def __init__(self, exps):
        self.exps = exps
        self.n, self.k = exps.shape
        self.keep = []

    def upload(self, X, Y):
        self.X, self.Y = X, Y
        self.n_t, self.m = Y.shape

    def gram_count(self):
        return self.k * self.k + self.k * self.n_t + self.n_t

    def upload_gram(self, X, Y, ptr):
        self.upload(X, Y)
        self.gram(ptr)

    def _view(self, ptr, count):
        import ctypes

        return np.ctypeslib.as_array((ctypes.c_double * count).from_address(ptr))

    def gram(self, ptr):
        Th = oracle.evaluate_library(self.exps, self.X)
        out = self._view(ptr, self.gram_count())
        k, nt = self.k, self.n_t
        out[: k * k] = (Th @ Th.T).ravel(order="F")
        out[k * k : k * k + k * nt] = (Th @ self.Y.T).ravel(order="F")
        out[k * k + k * nt :] = (self.Y * self.Y).sum(1)

    def gram_upload_ptr(self, pG, pR, pyy, n_t):
        k = self.k
        self.G = self._view(pG, k * k).reshape(k, k, order="F").copy()
        self.R = self._view(pR, k * n_t).reshape(k, n_t, order="F").copy()

    def sweep(self, prm, lambdas, m_total):
        # replicated sweep on the reduced blocks: equilibrated normal equations, STLSQ only
        self.m_total = m_total
        self.lambdas = np.sort(np.asarray(lambdas))
        d = 1.0 / np.sqrt(np.diag(self.G))
        As = self.G * np.outer(d, d)
        self.cands = []
        for e in range(self.n_t):
            rs = self.R[:, e] * d
            x = np.linalg.solve(As, rs) * d
            act = np.abs(x) > self.lambdas[0]
            seq = []
            for lam in self.lambdas:
                for _ in range(prm.maxiters):
                    xp = x.copy()
                    p = np.nonzero(act)[0]
                    x = np.zeros_like(x)
                    if p.size:
                        x[p] = np.linalg.solve(As[np.ix_(p, p)], rs[p]) * d[p]
                    act = np.abs(x) > lam
                    x[~act] = 0.0
                    seq.append((lam, x.copy()))
                    if not act.any() or np.linalg.norm(x - xp) < prm.abstol or np.linalg.norm(x - xp) / np.linalg.norm(x) < prm.reltol:
                        break
            self.cands.append(seq)

    def rss_count(self):
        return sum(len(s) for s in self.cands) + self.n_t

    def rss(self, ptr):
        out = self._view(ptr, self.rss_count())
        Th = oracle.evaluate_library(self.exps, self.X)
        i = 0
        for e in range(self.n_t):
            out[i] = np.sum(self.Y[e] ** 2)
            i += 1
            for _, x in self.cands[e]:
                out[i] = np.sum((x @ Th - self.Y[e]) ** 2)
                i += 1
        self._rss = out

    def select(self, paths=False):
        C = np.zeros((self.n_t, self.k))
        lam_opt = np.zeros(self.n_t)
        i = 0
        n = self.m_total
        for e in range(self.n_t):
            best = (np.inf, None, None)
            i += 1
            for j, (lam, x) in enumerate(self.cands[e]):
                score = n * np.log(self._rss[i] / n) + np.count_nonzero(x) * np.log(n)
                i += 1
                if lam == self.lambdas[0] or score <= best[0]:
                    best = (score, x, lam)
            C[e], lam_opt[e] = best[1], best[2]
        return ddb200.SweepOutput(C, lam_opt, None, None, None, None)


def _worker(rank, world, port, m, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, DX = systems.lorenz_ensemble(m, seed=1, samples_per_traj=200)
        DX = DX + 0.01 * np.random.default_rng(0).standard_normal(DX.shape)
        ex = oracle.polynomial_exponents(3, 2)
        lo, hi = ddb200.dist.shard_range(m, rank, world)
        ctx = FakeContext(ex)
        ctx.upload(X[:, lo:hi], DX[:, lo:hi])
        m_total = ddb200.dist.total_samples(hi - lo)
        prm = ddb200.Context.make_params("STLSQ")
        out = ddb200.dist.reduce_blocks_and_sweep(ctx, prm, np.array([0.05, 0.2, 0.5]), m_total, tensor_device="cpu")
        ret[rank] = (m_total, out.coefficients, out.lambda_opt, ctx.G)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_solve_equals_single_rank():
    m = 1000
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, m, ret), nprocs=2, join=True)
    assert ret[0][0] == ret[1][0] == m
    # both ranks hold the same reduced blocks and the same answer
    np.testing.assert_array_equal(ret[0][3], ret[1][3])
    np.testing.assert_array_equal(ret[0][1], ret[1][1])
    # and it is the single-process answer of the reference algorithm
    X, DX = systems.lorenz_ensemble(m, seed=1, samples_per_traj=200)
    DX = DX + 0.01 * np.random.default_rng(0).standard_normal(DX.shape)
    ex = oracle.polynomial_exponents(3, 2)
    Th = oracle.evaluate_library(ex, X)
    np.testing.assert_allclose(ret[0][3], Th @ Th.T, rtol=1e-13)
    ref = oracle.sparse_regression(oracle.STLSQ(np.array([0.05, 0.2, 0.5])), Th, DX)
    np.testing.assert_array_equal(ret[0][1] != 0, ref.coefficients != 0)
    np.testing.assert_allclose(ret[0][1], ref.coefficients, rtol=1e-9, atol=0)
    np.testing.assert_allclose(ret[0][2], ref.lambdas)
