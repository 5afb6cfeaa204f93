"""Pins the oracle against the reference's own known-answer tests.

Deterministic tests are restated exactly; tests that add StableRNG noise in the reference are restated
with NumPy Gaussian noise of the same amplitude and the reference's own tolerances.
"""
import numpy as np
import pytest

import oracle
from ddb200 import systems


def test_polynomial_basis_term_order():
    # test/Core/generators.jl:12-18: [1, u1, u1^2, u2, u1 u2, u2^2, u3, u1 u3, u2 u3, u3^2]
    e = oracle.polynomial_exponents(3, 2).T.tolist()
    assert e == [[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0], [1, 1, 0], [0, 2, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [0, 0, 2]]


def test_monomial_basis_terms():
    # test/Core/generators.jl:11: monomial_basis(u, 1) == [1; u .^ 1]
    assert oracle.monomial_exponents(3, 1).T.tolist() == [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]
    # src/utils/basis_generators.jl:134-149 : [1, x1, x1^2, x2, x2^2]
    assert oracle.monomial_exponents(2, 2).T.tolist() == [[0, 0], [1, 0], [2, 0], [0, 1], [0, 2]]


def test_library_size_is_binomial():
    from math import comb

    assert oracle.polynomial_exponents(3, 5).shape[1] == comb(8, 5) == 56
    assert oracle.polynomial_exponents(64, 2).shape[1] == comb(66, 2) == 2145


@pytest.mark.parametrize("prox", [oracle.SoftThreshold(), oracle.HardThreshold(), oracle.ClippedAbsoluteDeviation(1.0)])
def test_proximal_dispatch(prox):
    # lib/DataDrivenSparse/test/Core/sparse_linear_solve.jl:10-19
    c = np.array([2.0, 0.25])
    mask = oracle.apply_active_set(prox, c, 0.5)
    assert mask.tolist() == [True, False]
    assert c.tolist() == [2.0, 0.0]


def test_soft_threshold_values():
    # lib/DataDrivenSparse/test/Core/developer_api.jl:16-21
    assert oracle.SoftThreshold()(np.array([1.0, -2.0]), 0.5).tolist() == [0.5, -1.5]


def test_developer_api_shapes(golden_dir):
    # lib/DataDrivenSparse/test/Core/developer_api.jl:4-14
    g = np.load(f"{golden_dir}/ref_developer_api.npz")
    C, lams, iters = oracle.sparse.run_algorithm(oracle.STLSQ(float(g["threshold"])), g["Theta"], g["Y"])
    assert C.shape == (1, 2) and len(lams) == 1 and len(iters) == 1
    np.testing.assert_allclose(C @ g["Theta"], g["Y"], atol=1e-12)


def _fat_data(rng):
    # sparse_linear_solve.jl:27-45 ("Fat": 5 terms x 101 samples, noise 0.01)
    t = np.arange(0.0, 10.0 + 1e-9, 0.1)
    X = np.stack([np.sin(0.1 * t), np.cos(0.5 * t), np.sin(2.0 * t**2), np.cos(0.5 * t**2), np.exp(-t)])
    A = np.array([[0.68, 0.0, 0.0, 0.0, -1.2]])
    return X, A @ X + 0.01 * rng.standard_normal((1, t.size)), A


@pytest.mark.parametrize("alg", [oracle.STLSQ, oracle.ADMM, oracle.SR3])
def test_fat_statistics(alg):
    # sparse_linear_solve.jl:46-60: rss <= 1.2, aicc <= -400, dof == 2, r2 ~ 1 (atol 6e-2)
    X, Y, A = _fat_data(np.random.default_rng(42))
    lam = np.abs(A[A != 0])
    thresholds = np.linspace(0.5 * lam.min(), 1.5 * lam.max(), 20)
    solver = oracle.SparseLinearSolver(alg(thresholds), oracle.CommonOptions(maxiters=10_000))
    best, _, _ = solver.solve_target(X, Y[0])
    assert oracle.rss(best) <= 1.2
    assert oracle.aicc(best) <= -400.0
    assert oracle.dof(best) == 2
    assert abs(oracle.r2(best) - 1.0) <= 6e-2


@pytest.mark.parametrize("alg", [oracle.STLSQ, oracle.ADMM, oracle.SR3])
def test_skinny_statistics(alg):
    # sparse_linear_solve.jl:63-96: 6 terms x 5 samples, noise-free; rss <= 0.15, aicc <= -5, dof == 2
    rng = np.random.default_rng(52)
    t = np.arange(0.0, 2.0 + 1e-9, 0.5)
    X = np.stack([np.sin(0.5 * t), np.cos(0.5 * t), np.sin(2.0 * t**2), np.cos(0.5 * t**2), np.exp(-t), rng.standard_normal(t.size)])
    A = np.array([[0.68, 0.0, 0.0, 0.0, -1.2, 0.0]])
    Y = A @ X
    solver = oracle.SparseLinearSolver(alg(np.linspace(0.1, 1.6, 15)), oracle.CommonOptions(maxiters=10_000))
    best, _, _ = solver.solve_target(X, Y[0])
    assert oracle.rss(best) <= 1.5e-1
    assert oracle.aicc(best) <= -5.0
    assert oracle.dof(best) == 2
    assert abs(oracle.r2(best) - 1.0) <= 1e-1


def test_zero_target_does_not_throw():
    # sparse_linear_solve.jl:148-181: an all-zero target must run through without an exception
    rng = np.random.default_rng(0)
    Th = rng.standard_normal((4, 50))
    res = oracle.sparse_regression(oracle.STLSQ(np.array([0.1, 0.2])), Th, np.zeros((1, 50)))
    assert np.all(res.coefficients == 0.0)


def test_readme_lorenz_model():
    # README.md:40-85 (stale print format): x' = -10x + 10y ; y' = 28x - y - xz ; z' = xy - 8/3 z
    X, DX = systems.lorenz_readme()
    ex = oracle.polynomial_exponents(3, 5)
    lams = 10.0 ** np.arange(-5, -1 + 1e-9, 0.1)
    res = oracle.sparse_regression(oracle.STLSQ(lams), oracle.evaluate_library(ex, X), DX)
    C = oracle.construct_coefficients(res.coefficients)
    terms = {tuple(ex[:, j]): j for j in range(ex.shape[1])}
    expect = {0: {(1, 0, 0): -10.0, (0, 1, 0): 10.0}, 1: {(1, 0, 0): 28.0, (0, 1, 0): -1.0, (1, 0, 1): -1.0},
              2: {(1, 1, 0): 1.0, (0, 0, 1): -8.0 / 3.0}}
    for i, d in expect.items():
        assert set(np.nonzero(C[i])[0]) == {terms[t] for t in d}
        for t, v in d.items():
            assert abs(C[i, terms[t]] - v) <= 2e-9 * max(1.0, abs(v))


def test_rounding_is_toward_zero():
    # src/utils/build_basis.jl:154-157 with the defaults digits = 10, RoundToZero
    C = oracle.construct_coefficients(np.array([[2.66666666666666, -2.66666666666666, 1e-17, 0.99999999999]]))
    assert C.tolist() == [[2.6666666666, -2.6666666666, 0.0, 0.9999999999]]


@pytest.mark.parametrize("alg", [oracle.DMDPINV(), oracle.DMDSVD()])
def test_dmd_linear_discrete_groundtruth(alg):
    # lib/DataDrivenDMD/test/Core/linear_autonomous.jl:11-37
    A = np.array([[0.9, -0.2], [0.0, 0.2]])
    y = [np.array([10.0, -10.0])]
    for _ in range(10):
        y.append(A @ y[-1])
    x = np.stack(y, axis=1)
    r = oracle.koopman_solve(alg, x)
    np.testing.assert_allclose(r.K, A, atol=1e-10)
    np.testing.assert_allclose(r.C, np.eye(2), atol=1e-10)
    assert r.rss <= 1e-10
    # `dof(res) == 3` is asserted on the DataDrivenSolution, i.e. after convert_to_basis (solve.jl:86-100:
    # Theta = C * Matrix(K)) and the 10-digit RoundToZero post-processing of __construct_basis
    assert int(np.sum(oracle.construct_coefficients(r.C @ r.K) != 0)) == 3
    assert r.loglikelihood >= 400.0


def test_dmd_forced_unstable_operator():
    # lib/DataDrivenDMD/test/Core/linear_forced.jl:36-54 restated without the control channel
    A = np.array([[1.5, 0.0], [0.0, 0.1]])
    y = [np.array([1.0, 1.0])]
    for _ in range(8):
        y.append(A @ y[-1])
    r = oracle.koopman_solve(oracle.DMDSVD(), np.stack(y, axis=1))
    np.testing.assert_allclose(r.K, A, atol=1e-9)


def test_golden_fixtures_are_current(golden_dir):
    """The committed fixtures are what the oracle produces today (guards against silent drift)."""
    g = np.load(f"{golden_dir}/lorenz_noisy_deg3_stlsq.npz")
    Th = oracle.evaluate_library(g["exps"], g["X"])
    res = oracle.sparse_regression(oracle.STLSQ(g["thresholds"]), Th, g["Y"])
    np.testing.assert_array_equal(res.coefficients, g["coefficients"])
    np.testing.assert_allclose(Th @ Th.T, g["G"], rtol=0, atol=0)


def test_implicit_optimizer_reference_kat():
    """``lib/DataDrivenSparse/test/Core/sparse_linear_solve.jl:98-116`` (deterministic data, no RNG):
    ``vec(rescoeff) ~ [0.25, 0, -1, 0.11, 0, -0.5]`` atol 5e-2 for STLSQ(0.1, 1.0), ADMM(), SR3()."""
    t = np.arange(0.0, 10.0 + 1e-12, 0.1)
    X = np.vstack([np.sin(0.5 * t + 0.1), np.cos(0.5 * t), np.sin(2.0 * t**2), np.cos(0.5 * t**2 - 0.1), np.exp(-t)])
    Y = 0.5 * X[0] + 0.22 * X[3] - 2.0 * X[2]
    X = np.vstack([X, Y])
    for alg in (oracle.STLSQ(0.1, 1.0), oracle.ADMM(), oracle.SR3()):
        c, _lam, _it = oracle.ImplicitOptimizer(alg)(X, Y[None], options=oracle.CommonOptions(maxiters=2000))
        np.testing.assert_allclose(c.ravel(), [0.25, 0.0, -1.0, 0.11, 0.0, -0.5], atol=5e-2)


def test_implicit_candidate_matrix():
    """``commonsolve.jl:65-74``: a term is a candidate for an implicit variable when it depends on it or on no
    other implicit variable."""
    ii = np.array([[1, 0], [0, 1], [0, 0], [1, 1]], dtype=bool)
    cm = oracle.implicit_candidate_matrix(ii)
    np.testing.assert_array_equal(cm, [[True, False], [False, True], [True, True], [True, True]])


def test_generator_forms_match_the_reference():
    """``test/Core/generators.jl:7-11,20-23``: sin_basis(u,1) = sin.(1 .* u); cos_basis(u,2) = [cos.(u); cos.(2u)];
    chebyshev_basis(u,1) = cos.(acos.(u)); fourier_basis(u,1:2) = [sin.(u ./ 2); cos.(2u ./ 2)]."""
    rng = np.random.default_rng(0)
    U = rng.uniform(-0.9, 0.9, size=(3, 17))
    ev = lambda kind, c: oracle.evaluate_features(oracle.generator_features(kind, 3, c), U)
    np.testing.assert_array_equal(ev("sin", 1), np.sin(1 * U))
    np.testing.assert_array_equal(ev("cos", 2), np.vstack([np.cos(1 * U), np.cos(2 * U)]))
    np.testing.assert_array_equal(ev("chebyshev", 1), np.cos(1 * np.arccos(U)))
    np.testing.assert_array_equal(ev("fourier", 1), np.sin(1 * U / 2))
    np.testing.assert_array_equal(ev("sin", [1, 2]), np.vstack([np.sin(i * U) for i in (1, 2)]))
    np.testing.assert_array_equal(ev("cos", range(1, 6)), np.vstack([np.cos(i * U) for i in range(1, 6)]))
    np.testing.assert_array_equal(ev("chebyshev", [1, 2]), np.vstack([np.cos(i * np.arccos(U)) for i in (1, 2)]))
    np.testing.assert_array_equal(ev("fourier", [1, 2]), np.vstack([np.sin(1 * U / 2), np.cos(2 * U / 2)]))
    # the host mirror builds the same feature lists
    import ddb200

    for kind, fn in (("sin", ddb200.sin_basis), ("cos", ddb200.cos_basis), ("chebyshev", ddb200.chebyshev_basis),
                     ("fourier", ddb200.fourier_basis)):
        assert fn(3, 2).features == oracle.generator_features(kind, 3, 2)
    b = ddb200.concat_bases([ddb200.polynomial_basis(2, 5), ddb200.sin_basis(2, 1), ddb200.cos_basis(2, 1)])
    assert len(b) == 25 and b.n_states == 2 and b.term_strings()[-4:] == ["sin(x1)", "sin(x2)", "cos(x1)", "cos(x2)"]


def test_dmd_variants_reference_ground_truths():
    """``lib/DataDrivenDMD/test/Core/linear_autonomous.jl:11-37`` (every algorithm recovers
    ``[0.9 -0.2; 0.0 0.2]``) and ``linear_forced.jl:36-54`` (``K ~ [1.5 0; 0 0.1]`` with the known input matrix)."""
    A = np.array([[0.9, -0.2], [0.0, 0.2]])
    x = np.empty((2, 11))
    x[:, 0] = [10.0, -10.0]
    for i in range(10):
        x[:, i + 1] = A @ x[:, i]
    for alg in (oracle.TOTALDMD(0.0, oracle.DMDPINV()), oracle.TOTALDMD(0.0, oracle.DMDSVD()), oracle.FBDMD()):
        np.testing.assert_allclose(np.real(oracle.operator_matrix(*alg(x[:, :-1], x[:, 1:]))), A, atol=1e-8)
    X = np.array([[4, 2, 1, 0.5, 0.25], [7, 0.7, 0.07, 0.007, 0.0007]])
    U = np.array([[-4, -2, -1, -0.5, 0.0]])
    B = np.array([[1.0], [0.0]])
    Yeff = X[:, 1:] - B @ U[:, :-1]
    for alg in (oracle.DMDPINV(), oracle.DMDSVD(), oracle.TOTALDMD(2, oracle.DMDPINV())):
        np.testing.assert_allclose(np.real(oracle.operator_matrix(*alg(X[:, :-1], Yeff))), [[1.5, 0.0], [0.0, 0.1]], atol=1e-9)
