"""GPU parity of libraries with non-polynomial generators (``sin_basis``, ``cos_basis``,
``fourier_basis``, ``chebyshev_basis`` of ``src/utils/basis_generators.jl:30-82`` and their mixtures with
polynomials) through the feature map of the C ABI (``ddb_set_features``)."""
import numpy as np
import pytest

import ddb200
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ddb200.Context(0)
    yield c
    c.close()


def _theta(basis, X):
    F = oracle.evaluate_features(basis.features, X) if basis.features is not None else X
    return oracle.evaluate_library(basis.exponents, F)


def _pendulum_data(m=401, dt=0.05):
    """The reference's pendulum test system (``lib/DataDrivenSparse/test/Core/pendulum.jl:15-25``), RK4."""

    def f(u):
        return np.array([u[1], -9.81 * np.sin(u[0]) - 0.1 * u[1] ** 3 - 0.2 * np.cos(u[0])])

    u = np.array([0.99 * np.pi, -1.0])
    X = np.empty((2, m))
    h = dt / 20
    for s in range(m):
        X[:, s] = u
        for _ in range(20):
            k1 = f(u); k2 = f(u + 0.5 * h * k1); k3 = f(u + 0.5 * h * k2); k4 = f(u + h * k3)
            u = u + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    DX = np.stack([f(X[:, s]) for s in range(m)], axis=1)
    return X, DX


@pytest.mark.parametrize("make", [
    lambda: ddb200.sin_basis(3, 2), lambda: ddb200.cos_basis(3, [1, 3]), lambda: ddb200.fourier_basis(2, 4),
    lambda: ddb200.chebyshev_basis(2, 3),
    lambda: ddb200.concat_bases([ddb200.polynomial_basis(2, 5), ddb200.sin_basis(2, 1), ddb200.cos_basis(2, 1)]),
])
def test_gram_of_generator_libraries(ctx, make):
    basis = make()
    n, m = basis.n_states, 3001
    rng = np.random.default_rng(len(basis))
    X = rng.uniform(-0.95, 0.95, size=(n, m))
    Y = rng.standard_normal((n, m))
    ctx.set_features(basis.n_raw, basis.features)
    ctx.set_library(basis.exponents)
    ctx.upload(X, Y)
    ctx.gram()
    G, R, yy = ctx.gram_download()
    Th = _theta(basis, X)
    G0, R0 = Th @ Th.T, Th @ Y.T
    sc = np.sqrt(np.outer(np.diag(G0), np.diag(G0)))
    assert np.max(np.abs(G - G0) / sc) < 1e-13
    assert np.max(np.abs(R - R0)) < 1e-13 * np.sqrt(np.max(np.diag(G0)) * np.max((Y * Y).sum(1)))
    ctx.set_features(0, None)


@pytest.mark.parametrize("alg_name", ["STLSQ", "STLSQ-ridge", "ADMM", "SR3"])
def test_pendulum_library_solve_matches_oracle(ctx, alg_name):
    """``Basis([polynomial_basis(u, 5); sin.(u); cos.(u)], u)`` on the pendulum (pendulum.jl:13-46)."""
    X, DX = _pendulum_data()
    basis = ddb200.concat_bases([ddb200.polynomial_basis(2, 5), ddb200.sin_basis(2, 1), ddb200.cos_basis(2, 1)])
    # the reference's grid 1e-2:1e-2:1e-1 scaled by 0.97: the true coefficient -0.1 of x2^3 must not sit exactly
    # on a threshold (a 1-ulp knife edge decides the support there, in the reference as well)
    lams = 0.97 * np.arange(1e-2, 1e-1 + 1e-9, 1e-2)
    mk = {"STLSQ": lambda m: m.STLSQ(lams), "STLSQ-ridge": lambda m: m.STLSQ(lams, 1e-4),
          "ADMM": lambda m: m.ADMM(lams), "SR3": lambda m: m.SR3(lams)}[alg_name]
    alg = ddb200.B200Sparse(mk(ddb200))
    alg._ctx = ctx
    sol = ddb200.solve(ddb200.ContinuousDataDrivenProblem(X, DX=DX), basis, alg, paths=True)
    Th = _theta(basis, X)
    traces = []
    ref = oracle.sparse_regression(mk(oracle), Th, DX, traces=traces)
    C = sol.results[0].coefficients
    np.testing.assert_array_equal(C != 0.0, ref.coefficients != 0.0)
    tol = 1e-7 if alg_name in ("ADMM", "SR3") else 1e-8
    np.testing.assert_allclose(C, ref.coefficients, rtol=tol, atol=1e-11)
    for i, tr in enumerate(traces):  # identical support at every threshold
        np.testing.assert_array_equal(sol.paths.support_path[i], np.asarray(tr.support_path))
    if alg_name.startswith("STLSQ"):  # the reference's own assertions (pendulum.jl:36-44)
        assert 2 <= sol.results[0].dof <= 6
    ctx.set_features(0, None)


def test_feature_map_errors(ctx):
    with pytest.raises(ddb200._lib.DDBError):
        ctx.set_features(2, [(9, 0, 1.0)])  # unknown kind
    with pytest.raises(ddb200._lib.DDBError):
        ctx.set_features(2, [(1, 2, 1.0)])  # source out of range
    ctx.set_features(0, None)


def test_library_with_controls_and_time(ctx):
    """``Basis(eqs, x; controls = u, iv = t)``: to the kernels control inputs and the independent variable are
    more columns of the staged samples.  Forced linear system (the shape of
    ``lib/DataDrivenDMD/test/Core/linear_forced.jl:12-20``) with an explicitly time-dependent drift."""
    rng = np.random.default_rng(2)
    m = 1500
    t = np.linspace(0.0, 30.0, m)
    X = rng.uniform(-2, 2, size=(2, m))
    U = np.sin(t)[None, :]
    DX = np.vstack([-0.9 * X[0] + 0.1 * X[1], -0.8 * X[1] + 3.0 * U[0] + 0.05 * t])
    # variables: [x1, x2 | u1 | t]; library = all monomials of degree <= 2 in the four variables
    ex = oracle.polynomial_exponents(4, 2)
    basis = ddb200.Basis(ex, controls=1, use_time=True)
    assert basis.n_states == 2 and "u1" in basis.term_strings() and "t" in basis.term_strings()
    lams = 10.0 ** np.arange(-3, -0.9, 0.5)
    alg = ddb200.B200Sparse(ddb200.STLSQ(lams))
    alg._ctx = ctx
    sol = ddb200.solve(ddb200.ContinuousDataDrivenProblem(X, t=t, DX=DX, U=U), basis, alg)
    Th = oracle.evaluate_library(ex, np.vstack([X, U, t[None, :]]))
    ref = oracle.sparse_regression(oracle.STLSQ(lams), Th, DX)
    C = sol.results[0].coefficients
    np.testing.assert_array_equal(C != 0.0, ref.coefficients != 0.0)
    np.testing.assert_allclose(C, ref.coefficients, rtol=1e-9, atol=1e-12)
    names = basis.term_strings()
    got = {names[j]: C[1, j] for j in np.nonzero(C[1])[0]}
    assert set(got) == {"x2", "u1", "t"}
    assert got["u1"] == pytest.approx(3.0, rel=1e-9) and got["t"] == pytest.approx(0.05, rel=1e-9)
    with pytest.raises(ValueError):
        ddb200.solve(ddb200.ContinuousDataDrivenProblem(X, t=t, DX=DX), basis, alg)  # no control data


def test_expression_library_through_nvrtc():
    """Arbitrary terms and symbolic parameters (``src/basis/build_function.jl:6-38``): expressions in the raw variables
    and parameter values, compiled by NVRTC into one feature kernel; the Gram blocks and a full ``solve`` against the
    oracle on the same (host-evaluated) library."""
    rng = np.random.default_rng(21)
    m = 3000
    X = rng.uniform(0.2, 2.0, (2, m))
    p = np.array([0.7, 1.3])
    exprs = ["1.0", "x[0]", "x[1]", "x[0] * x[1]", "sin(p[0] * x[0]) * x[1]", "x[0] / (p[1] + x[1])", "exp(-x[0] * x[0])"]
    Theta = np.vstack([np.ones(m), X[0], X[1], X[0] * X[1], np.sin(p[0] * X[0]) * X[1], X[0] / (p[1] + X[1]), np.exp(-X[0] ** 2)])
    Ctrue = np.zeros((2, 7))
    Ctrue[0, [1, 4]] = [-0.8, 1.5]
    Ctrue[1, [0, 5, 3]] = [0.4, 2.0, -0.6]
    Y = Ctrue @ Theta + 1e-3 * rng.standard_normal((2, m))
    basis = ddb200.expression_basis(2, exprs, params=p)
    assert len(basis) == 7 and basis.n_states == 2
    with ddb200.Context(0) as ctx:
        ctx.set_features(2, basis.features)
        ctx.set_library(basis.exponents)
        ctx.upload(X, Y)
        ctx.gram()
        G, R, yy = ctx.gram_download()
    G0, R0 = Theta @ Theta.T, Theta @ Y.T
    sc = np.sqrt(np.outer(np.diag(G0), np.diag(G0)))
    assert np.max(np.abs(G - G0) / sc) < 1e-12  # device sin / exp / division against NumPy's: a few ulp per feature
    assert np.max(np.abs(R - R0)) < 1e-12 * np.sqrt(np.max(np.diag(G0)) * np.max((Y * Y).sum(1)))
    lams = 10.0 ** np.arange(-2, -0.2, 0.25)
    sol = ddb200.solve(ddb200.DirectDataDrivenProblem(X, Y), basis, ddb200.B200Sparse(ddb200.STLSQ(lams)))
    ref = oracle.sparse_regression(oracle.STLSQ(lams), Theta, Y)
    np.testing.assert_array_equal(sol.results[0].coefficients != 0, ref.coefficients != 0)
    np.testing.assert_array_equal(ref.coefficients != 0, Ctrue != 0)
    nz = ref.coefficients != 0
    assert (np.abs(sol.results[0].coefficients[nz] - ref.coefficients[nz]) / np.abs(ref.coefficients[nz])).max() <= 1e-9
    # a source that does not compile is an error with the compiler's message, never a silent fallback
    bad = ddb200.expression_basis(2, ["x[0] +* 2"])
    with ddb200.Context(0) as ctx:
        with pytest.raises(ddb200._lib.DDBError) as e:
            ctx.set_features(2, bad.features)
        assert e.value.status == -1 and "error" in str(e.value)
