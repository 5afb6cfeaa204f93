"""The exact-rss pass of DENSE candidates (csrc/rss_dense.cu: fit = coefficient block x Theta block on DMMA, residuals
squared in the epilogue) against the term-by-term pass of rss.cu on the same candidates (``DDB_RSS_DENSE=0``) and against
a NumPy evaluation of the selected model.  Reference: the rss every recorded model is scored with,
lib/DataDrivenSparse/src/solver.jl:72-118."""
import os

import numpy as np
import pytest

import ddb200
import oracle
from ddb200 import systems

pytestmark = pytest.mark.gpu


def _solve(X, D, basis, lams, dense):
    old = os.environ.get("DDB_RSS_DENSE")
    os.environ["DDB_RSS_DENSE"] = "1" if dense else "0"
    try:
        with ddb200.Context(0) as ctx:
            ctx.set_library(basis.exponents)
            prm = ctx.make_params("STLSQ")
            out = ctx.solve_sparse(X, D, prm, lams)
            ctx.sweep(prm, lams)  # the candidate table itself: every entry, not only the winners
            n_c = ctx.rss_count()
            import torch

            t = torch.empty(n_c, dtype=torch.float64, device="cuda:0")
            ctx.rss(t.data_ptr())
            ctx.synchronize()
            return out, t.cpu().numpy()
    finally:
        if old is None:
            os.environ.pop("DDB_RSS_DENSE", None)
        else:
            os.environ["DDB_RSS_DENSE"] = old


@pytest.mark.parametrize("n,deg,m,noise", [(18, 2, 5000, 0.5), (15, 2, 4001, 0.3), (8, 3, 3000, 0.5), (64, 2, 6113, 1.0)])
def test_dense_pass_matches_term_by_term(n, deg, m, noise):
    X, D = systems.lorenz96_ensemble(m, n=n, seed=7, samples_per_traj=500)
    D = D + noise * np.random.default_rng(3).standard_normal(D.shape)
    basis = ddb200.polynomial_basis(n, deg)
    assert basis.exponents.shape[1] >= 128
    lams = 10.0 ** np.arange(-4, -0.4, 0.25)
    out_d, tab_d = _solve(X, D, basis, lams, True)
    out_s, tab_s = _solve(X, D, basis, lams, False)
    # candidate by candidate (streamed part of the table; zero models and group members carry 0 there in both runs)
    assert tab_d.shape == tab_s.shape
    np.testing.assert_allclose(tab_d, tab_s, rtol=1e-11, atol=1e-11 * np.max(tab_s))
    assert np.count_nonzero(tab_d != tab_s) > 0  # the dense pass really ran (another summation order)
    np.testing.assert_array_equal(out_d.coefficients, out_s.coefficients)
    np.testing.assert_allclose(out_d.lambda_opt, out_s.lambda_opt, rtol=1e-15)
    np.testing.assert_allclose(out_d.rss, out_s.rss, rtol=1e-11)
    # and against a direct evaluation of the winners
    theta = oracle.evaluate_library(basis.exponents, X)
    res = out_d.coefficients @ theta - D
    np.testing.assert_allclose(out_d.rss, np.sum(res * res, axis=1), rtol=1e-10)


def test_dense_pass_with_term_normalisation():
    """ZScore normalisation: the candidates live in the scaled space, the dense pass divides the coefficients by the
    scales like the term-by-term pass does (``dense_fill_kernel``)."""
    n, m = 18, 4000
    X, D = systems.lorenz96_ensemble(m, n=n, seed=9, samples_per_traj=500)
    D = D + 0.5 * np.random.default_rng(5).standard_normal(D.shape)
    basis = ddb200.polynomial_basis(n, 2)
    lams = 10.0 ** np.arange(-4, -0.4, 0.25)
    opts = ddb200.DataDrivenCommonOptions(normalize="ZScoreTransform")
    prob = ddb200.ContinuousDataDrivenProblem(X, DX=D)
    sols = []
    for dense in ("1", "0"):
        old = os.environ.get("DDB_RSS_DENSE")
        os.environ["DDB_RSS_DENSE"] = dense
        try:
            sols.append(ddb200.solve(prob, basis, ddb200.B200Sparse(ddb200.STLSQ(lams)), opts))
        finally:
            if old is None:
                os.environ.pop("DDB_RSS_DENSE", None)
            else:
                os.environ["DDB_RSS_DENSE"] = old
    a, b = sols
    np.testing.assert_array_equal(a.coefficients, b.coefficients)
    np.testing.assert_allclose(a.results[0].trainerror, b.results[0].trainerror, rtol=1e-10)
    np.testing.assert_allclose(a.rss, b.rss, rtol=1e-10)
    assert a.timings["big_steps"] > 0  # candidates of more than 128 terms exist: the dense pass had work
