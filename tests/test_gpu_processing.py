"""GPU parity of the reference's pre-processing options: ZScore normalisation of the
library rows (``src/utils/data_processing.jl:64-126``), train/test split and shuffled mini-batches
(``:29-39``), best batch by test error (``lib/DataDrivenSparse/src/commonsolve.jl:10-21``), and the device
residual pass that replaces the CPU re-evaluation of the solution (``src/solution.jl:37-56``, row f2)."""
import numpy as np
import pytest

import ddb200
import oracle
from ddb200 import systems

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ddb200.Context(0)
    yield c
    c.close()


def _lorenz(m=2000, noise=0.05, seed=1):
    X, DX = systems.lorenz_ensemble(m, seed=seed, samples_per_traj=500)
    rng = np.random.default_rng(seed)
    return X, DX + noise * rng.standard_normal(DX.shape)


def _solve(ctx, basis, X, DX, inner, **opts):
    alg = ddb200.B200Sparse(inner)
    alg._ctx = ctx
    return ddb200.solve(ddb200.ContinuousDataDrivenProblem(X, DX=DX), basis, alg, ddb200.DataDrivenCommonOptions(**opts))


# thresholds large enough that the selected models are sparse: with a dense active set the normal-equation
# solve of this noisy degree-3 library is only good to ~1e-7 against the reference's QR
LAMS = 10.0 ** np.arange(-1, 1.01, 0.25)
ALGS = {"STLSQ": lambda m: m.STLSQ(LAMS), "STLSQ-ridge": lambda m: m.STLSQ(LAMS, 1e-3), "ADMM": lambda m: m.ADMM(LAMS[::2]),
        "SR3": lambda m: m.SR3(LAMS)}


@pytest.mark.parametrize("name", list(ALGS))
@pytest.mark.parametrize("with_const", [True, False])
def test_zscore_normalisation_matches_oracle(ctx, name, with_const):
    X, DX = _lorenz()
    basis = ddb200.polynomial_basis(3, 3)
    if not with_const:  # library without the constant term: the column sums need their own pass
        basis = ddb200.Basis(basis.exponents[:, 1:])
    Th = oracle.evaluate_library(basis.exponents, X)
    C_ref, res_ref, sigma = oracle.solve_processed(ALGS[name](oracle), Th, DX, normalize="ZScoreTransform")
    sol = _solve(ctx, basis, X, DX, ALGS[name](ddb200), normalize="ZScoreTransform")
    Cs = sol.results[0].coefficients  # scaled space: thresholds act here
    np.testing.assert_array_equal(Cs != 0.0, res_ref[0].coefficients != 0.0)
    # noisy data, cond(Theta) ~ 1e5: both sides solve normal equations here (ridge) or QR vs Gram (rho = 0) and
    # agree to ~1e-8; the 1e-10 criterion of the north star is checked on the well-conditioned configurations
    np.testing.assert_allclose(Cs, res_ref[0].coefficients, rtol=1e-7, atol=1e-11)
    np.testing.assert_allclose(sol.coefficients, oracle.construct_coefficients(C_ref), rtol=1e-7, atol=2e-10)
    np.testing.assert_allclose(sol.results[0].trainerror, res_ref[0].trainerror, rtol=1e-7)


@pytest.mark.parametrize("normalize", [None, "ZScoreTransform"])
def test_split_and_batches_match_oracle(ctx, normalize):
    X, DX = _lorenz(m=2000)
    X, DX = X[:, :1501], DX[:, :1501]
    basis = ddb200.polynomial_basis(3, 2)
    Th = oracle.evaluate_library(basis.exponents, X)
    m = X.shape[1]
    m_train = int(round(0.8 * m))
    perm = np.random.default_rng(3).permutation(m_train)
    kw = dict(split=0.8, batchsize=250, partial=True)
    C_ref, res_ref, _ = oracle.solve_processed(oracle.STLSQ(LAMS), Th, DX, normalize=normalize, perm=perm, **kw)
    sol = _solve(ctx, basis, X, DX, ddb200.STLSQ(LAMS), normalize=normalize, perm=perm, **kw)
    assert len(sol.results) == len(res_ref) == 5  # 1201 training samples: 4 x 250 + 201; the test part starts at an odd sample
    for got, ref in zip(sol.results, res_ref):  # same order after sort!(results, by = l2error)
        np.testing.assert_array_equal(got.coefficients != 0.0, ref.coefficients != 0.0)
        np.testing.assert_allclose(got.coefficients, ref.coefficients, rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(got.testerror, ref.testerror, rtol=1e-7)
        np.testing.assert_allclose(got.trainerror, ref.trainerror, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(sol.coefficients, oracle.construct_coefficients(C_ref), rtol=1e-7, atol=2e-10)
    # rss of the final (rounded) model on ALL samples, as DataDrivenSolution recomputes it
    assert sol.rss == pytest.approx(float(np.sum((DX - sol.coefficients @ Th) ** 2)), rel=1e-9)
    # partial = false drops the one-sample batch
    sol2 = _solve(ctx, basis, X, DX, ddb200.STLSQ(LAMS), normalize=normalize, perm=perm, split=0.8, batchsize=250, partial=False)
    assert len(sol2.results) == 4


def test_residual_permute_and_range(ctx):
    X, DX = _lorenz(m=1000)
    ex = oracle.polynomial_exponents(3, 2)
    Th = oracle.evaluate_library(ex, X)
    rng = np.random.default_rng(0)
    C = rng.standard_normal((3, ex.shape[1])) * (rng.random((3, ex.shape[1])) < 0.4)
    ctx.set_features(0, None)
    ctx.set_library(ex)
    ctx.upload(X, DX)
    np.testing.assert_allclose(ctx.residual(C), ((DX - C @ Th) ** 2).sum(1), rtol=1e-12)
    ctx.set_sample_range(200, 701)
    np.testing.assert_allclose(ctx.residual(C), ((DX[:, 200:701] - C @ Th[:, 200:701]) ** 2).sum(1), rtol=1e-12)
    ctx.gram()
    G, R, yy = ctx.gram_download()
    np.testing.assert_allclose(G, Th[:, 200:701] @ Th[:, 200:701].T, rtol=1e-12, atol=1e-9)
    ctx.set_sample_range(0, -1)
    perm = rng.permutation(1000)
    ctx.permute_samples(perm)
    Xd, Yd = ctx.download_data()
    np.testing.assert_array_equal(Xd, X[:, perm])
    np.testing.assert_array_equal(Yd, DX[:, perm])
    with pytest.raises(ddb200._lib.DDBError):
        ctx.permute_samples(np.zeros(1000, dtype=np.int64))
    with pytest.raises(ddb200._lib.DDBError):
        ctx.set_sample_range(10, 5)
    ctx.set_sample_range(201, 500)  # odd first sample: fine for the residual pass
    np.testing.assert_allclose(ctx.residual(C), ((DX[:, perm][:, 201:500] - C @ Th[:, perm][:, 201:500]) ** 2).sum(1), rtol=1e-12)


def test_odd_batch_size_on_a_three_state_problem(ctx):
    """Batches of 125 samples of a 3-state problem start on odd doubles (``first * n`` odd): the views are staged
    into aligned buffers for the Gram kernel's bulk copies; the reference accepts any batch size."""
    X, DX = _lorenz(m=1000)
    basis = ddb200.polynomial_basis(3, 2)
    Th = oracle.evaluate_library(basis.exponents, X)
    perm = np.random.default_rng(5).permutation(875)
    kw = dict(split=0.875, batchsize=125, partial=True)
    C_ref, res_ref, _ = oracle.solve_processed(oracle.STLSQ(LAMS), Th, DX, perm=perm, **kw)
    sol = _solve(ctx, basis, X, DX, ddb200.STLSQ(LAMS), perm=perm, **kw)
    assert len(sol.results) == len(res_ref) == 7
    for got, ref in zip(sol.results, res_ref):
        np.testing.assert_array_equal(got.coefficients != 0.0, ref.coefficients != 0.0)
        np.testing.assert_allclose(got.coefficients, ref.coefficients, rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(got.testerror, ref.testerror, rtol=1e-7)
    ctx.set_features(0, None)
    ctx.set_library(basis.exponents)
    ctx.upload(X, DX)
    ctx.set_sample_range(125, 250)  # 375 doubles into X: not 16-byte aligned
    ctx.gram()
    G = ctx.gram_download()[0]
    np.testing.assert_allclose(G, Th[:, 125:250] @ Th[:, 125:250].T, rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("n_t,m", [(3, 2000), (1, 77), (64, 5000), (300, 1001)])
def test_target_stats_kernel(ctx, n_t, m):
    """``ddb_target_stats``: per-target mean and centred sum of squares (two passes, deterministic), any target count."""
    rng = np.random.default_rng(n_t + m)
    X = rng.standard_normal((2, m))
    Y = 50.0 + rng.standard_normal((n_t, m)) * np.linspace(0.1, 3.0, n_t)[:, None]  # mean >> spread: one-pass sums would cancel
    ctx.set_features(0, None)
    ctx.set_library(oracle.polynomial_exponents(2, 1))
    ctx.upload(X, Y)
    mean, css = ctx.target_stats()
    np.testing.assert_allclose(mean, Y.mean(axis=1), rtol=1e-14)
    np.testing.assert_allclose(css, ((Y - Y.mean(axis=1, keepdims=True)) ** 2).sum(axis=1), rtol=1e-12)
    m2, c2 = ctx.target_stats()
    np.testing.assert_array_equal(css, c2)  # bit-reproducible
    ctx.set_sample_range(10, 60)
    mean, css = ctx.target_stats(50)
    np.testing.assert_allclose(css, ((Y[:, 10:60] - Y[:, 10:60].mean(axis=1, keepdims=True)) ** 2).sum(axis=1), rtol=1e-12)
    ctx.set_sample_range(0, -1)


def test_solution_statistics_match_the_reference_formulas(ctx):
    """StatsAPI statistics of ``DataDrivenSolution`` (``src/solution.jl:105-157``): rss, nobs, loglikelihood,
    nullloglikelihood, r2 (Cox-Snell), aic / aicc / bic -- from the device passes, against NumPy on the same data."""
    X, DX = _lorenz(m=1500, noise=0.1)
    basis = ddb200.polynomial_basis(3, 2)
    sol = _solve(ctx, basis, X, DX, ddb200.STLSQ(LAMS))
    Th = oracle.evaluate_library(basis.exponents, X)
    n = DX.size
    rss = float(np.sum((DX - sol.results[0].coefficients @ Th) ** 2))
    ll = -n / 2 * np.log(rss / n)
    nll = -n / 2 * np.log(np.sum((DX - DX.mean(axis=1, keepdims=True)) ** 2) / n)
    assert sol.nobs == n
    assert sol.rss == pytest.approx(rss, rel=1e-9)
    assert sol.loglikelihood() == pytest.approx(ll, rel=1e-9)
    assert sol.nullloglikelihood() == pytest.approx(nll, rel=1e-12)
    assert sol.r2() == pytest.approx(1 - np.exp(2 * (nll - ll) / n), rel=1e-9)
    k = sol.dof
    assert sol.aic() == pytest.approx(-2 * ll + 2 * k, rel=1e-9)
    assert sol.bic() == pytest.approx(-2 * ll + k * np.log(n), rel=1e-9)
    assert sol.aicc() == pytest.approx(-2 * ll + 2 * k + 2 * k * (k + 1) / (n - k - 1), rel=1e-9)
    assert 0.99 < sol.r2() <= 1.0  # the reference's own tests assert r2 ~ 1 on identifiable systems


@pytest.mark.parametrize("name", ["STLSQ", "STLSQ-ridge", "SR3"])
@pytest.mark.parametrize("with_const", [True, False])
def test_unit_range_normalisation_matches_oracle(ctx, name, with_const):
    """``DataNormalization(UnitRangeTransform)`` (``src/utils/data_processing.jl:68-73, 119-126``): every library row is
    shifted by its minimum and divided by its range (fitted on all samples), the regression runs on the transformed
    rows and -- as the reference is written -- the same transform is applied to the coefficient matrix afterwards
    (``lib/DataDrivenSparse/src/commonsolve.jl:19-21``).  On the device: term minima / maxima in one pass
    (``ddb_term_minmax``), the blocks of the transformed rows from the blocks of the raw ones
    (``ddb_gram_affine_terms``), candidates scored on the raw data with the equivalent constant."""
    X, DX = _lorenz(m=1500)
    basis = ddb200.polynomial_basis(3, 2)
    if not with_const:
        basis = ddb200.Basis(basis.exponents[:, 1:])
    lams = 10.0 ** np.arange(-0.5, 1.6, 0.25)  # coefficients of the unit-range rows are ~range x the raw ones
    algs = {"STLSQ": lambda mm: mm.STLSQ(lams), "STLSQ-ridge": lambda mm: mm.STLSQ(lams, 1e-3), "SR3": lambda mm: mm.SR3(lams)}
    Th = oracle.evaluate_library(basis.exponents, X)
    C_ref, res_ref, _ = oracle.solve_processed(algs[name](oracle), Th, DX, normalize="UnitRangeTransform", split=0.8)
    sol = _solve(ctx, basis, X, DX, algs[name](ddb200), normalize="UnitRangeTransform", split=0.8)
    Cs = sol.results[0].coefficients
    np.testing.assert_array_equal(Cs != 0.0, res_ref[0].coefficients != 0.0)
    np.testing.assert_allclose(Cs, res_ref[0].coefficients, rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(sol.results[0].trainerror, res_ref[0].trainerror, rtol=1e-6)
    np.testing.assert_allclose(sol.results[0].testerror, res_ref[0].testerror, rtol=1e-6)
    np.testing.assert_allclose(sol.coefficients, oracle.construct_coefficients(C_ref), rtol=2e-6, atol=2e-9)


def test_term_minmax(ctx):
    rng = np.random.default_rng(3)
    X = rng.standard_normal((5, 4001))
    ex = oracle.polynomial_exponents(5, 3)
    Th = oracle.evaluate_library(ex, X)
    ctx.set_features(0, None)
    ctx.set_library(ex)
    ctx.upload(X, X[:2])
    mn, mx = ctx.term_minmax()
    np.testing.assert_allclose(mn, Th.min(axis=1), rtol=1e-14)
    np.testing.assert_allclose(mx, Th.max(axis=1), rtol=1e-14)
    ctx.set_sample_range(100, 777)
    mn, mx = ctx.term_minmax()
    np.testing.assert_allclose(mn, Th[:, 100:777].min(axis=1), rtol=1e-14)
    ctx.set_sample_range(0, -1)


def test_zscore_scales_are_two_pass(ctx):
    """the one-pass variance ``(sum theta^2 - (sum theta)^2 / m) / (m - 1)`` cancels for terms whose mean
    dwarfs their spread (monomials of offset states).  The scales come from ``ddb_term_css`` -- centred sums of squares,
    the second pass of ``std`` in the reference (``data_processing.jl:75-80``) -- and only an exactly zero spread is a
    constant."""
    rng = np.random.default_rng(12)
    m = 5001
    X = np.array([[1000.0], [3.0], [0.0]]) + np.array([[1e-2], [1.0], [2.0]]) * rng.standard_normal((3, m))
    basis = ddb200.polynomial_basis(3, 3)
    Th = oracle.evaluate_library(basis.exponents, X)
    ctx.set_library(basis.exponents)
    ctx.upload(X, X)
    mean = Th.mean(axis=1)
    css = ctx.term_css(mean)
    ref = np.sum((Th - mean[:, None]) ** 2, axis=1)
    np.testing.assert_allclose(css, ref, rtol=1e-11, atol=0.0)
    assert css[0] == 0.0  # the constant term: exactly zero spread
    sigma = ddb200.api.zscore_scales(css, m)
    np.testing.assert_allclose(sigma, oracle.zscore_fit(Th), rtol=1e-11)
    # the one-pass form is visibly wrong here: x1^3 has mean 1e9 and spread 3e4
    onepass = np.sqrt(np.maximum((np.sum(Th * Th, axis=1) - Th.sum(axis=1) ** 2 / m) / (m - 1), 0.0))
    j = int(np.argmax(np.abs(mean)))
    assert abs(onepass[j] / sigma[j] - 1.0) > 1e-9 > abs(np.sqrt(ref[j] / (m - 1)) / sigma[j] - 1.0)
