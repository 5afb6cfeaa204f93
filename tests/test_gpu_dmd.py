"""GPU parity of the DataDrivenDMD Gram path (P = X X', Q = Y X', K = Q P^-1) against the oracle."""
import numpy as np
import pytest

import ddb200
import oracle
from ddb200 import systems

pytestmark = pytest.mark.gpu


def _snapshots(A, x0, steps):
    y = [np.asarray(x0, dtype=float)]
    for _ in range(steps):
        y.append(A @ y[-1])
    return np.stack(y, axis=1)


@pytest.mark.parametrize("inner", [ddb200.DMDPINV(), ddb200.DMDSVD()])
def test_linear_discrete_groundtruth(inner):
    # lib/DataDrivenDMD/test/Core/linear_autonomous.jl:11-37
    A = np.array([[0.9, -0.2], [0.0, 0.2]])
    x = _snapshots(A, [10.0, -10.0], 10)
    res = ddb200.solve_dmd(ddb200.DiscreteDataDrivenProblem(x), ddb200.B200DMD(inner))
    ref = oracle.koopman_solve(oracle.DMDPINV() if inner.name == "DMDPINV" else oracle.DMDSVD(), x)
    np.testing.assert_allclose(res.P, ref.P, rtol=1e-13)
    np.testing.assert_allclose(res.Q, ref.Q, rtol=1e-13)
    np.testing.assert_allclose(np.real(res.K), A, atol=1e-9)
    np.testing.assert_allclose(np.real(res.K), ref.K, atol=1e-9)
    np.testing.assert_allclose(np.sort(np.real(res.eigenvalues)), [0.2, 0.9], atol=1e-9)
    assert res.rss <= 1e-10 and res.dof == 3


@pytest.mark.parametrize("n,m", [(6, 1000), (130, 4001)])
def test_single_trajectory_operator(n, m):
    A, X, Y = systems.linear_snapshots(n, m, seed=n, restart=10**9)  # one trajectory: Y is X shifted by one
    x = np.concatenate([X, Y[:, -1:]], axis=1)
    res = ddb200.solve_dmd(ddb200.DiscreteDataDrivenProblem(x), ddb200.B200DMD(ddb200.DMDPINV()))
    P0, Q0 = X @ X.T, Y @ X.T
    sc = np.sqrt(np.outer(np.diag(P0), np.diag(P0)))
    assert np.max(np.abs(res.P - P0) / sc) < 1e-13
    assert np.max(np.abs(res.Q - Q0) / sc) < 1e-12
    np.testing.assert_array_equal(res.P, res.P.T)
    # K = Y / X  (pivoted QR in the reference) vs K = Q P^-1 (equilibrated Cholesky here)
    Kref = np.linalg.lstsq(X.T, Y.T, rcond=None)[0].T
    assert np.max(np.abs(np.real(res.K) - Kref)) <= 1e-8 * max(1.0, np.max(np.abs(Kref)))
    np.testing.assert_allclose(np.real(res.K), A, atol=1e-6)


@pytest.mark.parametrize("n,m", [(2, 16), (256, 3000), (300, 777), (514, 2049)])
def test_gram_blocks_and_operator_general(n, m):
    """Separate X and Y (restarted trajectories keep the snapshot matrix well conditioned): the C-ABI
    entry points directly, including ragged block rows (n not a multiple of 128) and a ragged last stage."""
    import torch

    A, X, Y = systems.linear_snapshots(n, m, seed=n, restart=64)
    dev = torch.device("cuda:0")
    Xt = torch.from_numpy(np.ascontiguousarray(X.T)).to(dev)  # (m, n) row-major == n x m column-major
    Yt = torch.from_numpy(np.ascontiguousarray(Y.T)).to(dev)
    PQ = torch.zeros((2, n, n), dtype=torch.float64, device=dev)
    K = torch.zeros((n, n), dtype=torch.float64, device=dev)
    with ddb200.Context(0) as ctx:
        ctx.dmd_gram(Xt.data_ptr(), Yt.data_ptr(), n, m, PQ[0].data_ptr(), PQ[1].data_ptr())
        ctx.dmd_operator(PQ[0].data_ptr(), PQ[1].data_ptr(), n, K.data_ptr())
    P, Q, Kh = PQ[0].T.cpu().numpy(), PQ[1].T.cpu().numpy(), K.T.cpu().numpy()
    P0, Q0 = X @ X.T, Y @ X.T
    sc = np.sqrt(np.outer(np.diag(P0), np.diag(P0)))
    assert np.max(np.abs(P - P0) / sc) < 1e-13
    assert np.max(np.abs(Q - Q0) / sc) < 1e-12
    np.testing.assert_array_equal(P, P.T)
    if m > 4 * n:  # K only when X X' is comfortably full rank
        Kref = np.linalg.lstsq(X.T, Y.T, rcond=None)[0].T
        assert np.max(np.abs(Kh - Kref)) <= 1e-9 * max(1.0, np.max(np.abs(Kref)))
        np.testing.assert_allclose(Kh, A, atol=1e-8)


# ------------------------------------------------------------------------------------------------
# variants on the Gram blocks: TOTALDMD, FBDMD, controlled DMD
# ------------------------------------------------------------------------------------------------
def _variant(name):
    return {
        "TOTALDMD-PINV": (ddb200.TOTALDMD(0.0, ddb200.DMDPINV()), oracle.TOTALDMD(0.0, oracle.DMDPINV())),
        "TOTALDMD-SVD": (ddb200.TOTALDMD(0.0, ddb200.DMDSVD()), oracle.TOTALDMD(0.0, oracle.DMDSVD())),
        "FBDMD": (ddb200.FBDMD(), oracle.FBDMD()),
    }[name]


@pytest.mark.parametrize("name", ["TOTALDMD-PINV", "TOTALDMD-SVD", "FBDMD"])
def test_variants_linear_discrete_groundtruth(name):
    # lib/DataDrivenDMD/test/Core/linear_autonomous.jl:11-37 runs the same ground truth for every algorithm
    A = np.array([[0.9, -0.2], [0.0, 0.2]])
    x = _snapshots(A, [10.0, -10.0], 10)
    mine, ref = _variant(name)
    res = ddb200.solve_dmd(ddb200.DiscreteDataDrivenProblem(x), ddb200.B200DMD(mine))
    Kref = np.real(oracle.operator_matrix(*ref(x[:, :-1], x[:, 1:])))
    np.testing.assert_allclose(np.real(res.K), A, atol=1e-8)
    np.testing.assert_allclose(np.real(res.K), Kref, atol=1e-8)
    assert res.rss <= 1e-10 and res.dof == 3


@pytest.mark.parametrize("name", ["TOTALDMD-PINV", "TOTALDMD-SVD", "FBDMD"])
def test_variants_match_oracle_on_noisy_snapshots(name):
    n, m = 12, 3000
    A, X, Y = systems.linear_snapshots(n, m, seed=5, restart=10**9)
    x = np.concatenate([X, Y[:, -1:]], axis=1)
    x = x + 1e-3 * np.random.default_rng(0).standard_normal(x.shape)  # total-least-squares setting: noise in X and Y
    mine, ref = _variant(name)
    res = ddb200.solve_dmd(ddb200.DiscreteDataDrivenProblem(x), ddb200.B200DMD(mine))
    lam, phi = ref(x[:, :-1], x[:, 1:])
    Kref = np.real(oracle.operator_matrix(lam, phi))
    assert np.max(np.abs(np.real(res.K) - Kref)) <= 1e-6 * max(1.0, np.max(np.abs(Kref)))
    np.testing.assert_allclose(np.sort_complex(res.eigenvalues), np.sort_complex(lam), atol=1e-6)


def test_totaldmd_integer_truncation():
    n, m = 8, 2000
    A, X, Y = systems.linear_snapshots(n, m, seed=2, restart=10**9)
    x = np.concatenate([X, Y[:, -1:]], axis=1)
    res = ddb200.solve_dmd(ddb200.DiscreteDataDrivenProblem(x), ddb200.B200DMD(ddb200.TOTALDMD(6, ddb200.DMDSVD())))
    lam, phi = oracle.TOTALDMD(6, oracle.DMDSVD())(x[:, :-1], x[:, 1:])
    np.testing.assert_allclose(np.sort_complex(res.eigenvalues), np.sort_complex(lam), atol=1e-7)


@pytest.mark.parametrize("inner", ["DMDPINV", "DMDSVD", "TOTALDMD"])
def test_forced_unstable_system_with_known_input_matrix(inner):
    # lib/DataDrivenDMD/test/Core/linear_forced.jl:36-54
    X = np.array([[4, 2, 1, 0.5, 0.25], [7, 0.7, 0.07, 0.007, 0.0007]])
    U = np.array([[-4, -2, -1, -0.5, 0.0]])
    B = np.array([[1.0], [0.0]])
    alg = {"DMDPINV": ddb200.DMDPINV(), "DMDSVD": ddb200.DMDSVD(), "TOTALDMD": ddb200.TOTALDMD(2, ddb200.DMDPINV())}[inner]
    res = ddb200.solve_dmd(ddb200.DiscreteDataDrivenProblem(X, U=U), ddb200.B200DMD(alg), control_input=B)
    np.testing.assert_allclose(np.real(res.K), [[1.5, 0.0], [0.0, 0.1]], atol=1e-9)
    np.testing.assert_array_equal(res.B, B)
    assert res.dof == 3 and res.rss <= 1e-20


@pytest.mark.parametrize("inner", ["DMDPINV", "DMDSVD"])
def test_controlled_dmd_matches_oracle(inner):
    rng = np.random.default_rng(11)
    n, nu, m = 6, 2, 800
    A = 0.9 * np.linalg.qr(rng.standard_normal((n, n)))[0]
    Bt = rng.standard_normal((n, nu))
    U = rng.standard_normal((nu, m + 1))
    x = np.empty((n, m + 1))
    x[:, 0] = rng.standard_normal(n)
    for j in range(m):
        x[:, j + 1] = A @ x[:, j] + Bt @ U[:, j]
    alg = ddb200.DMDPINV() if inner == "DMDPINV" else ddb200.DMDSVD()
    res = ddb200.solve_dmd(ddb200.DiscreteDataDrivenProblem(x, U=U), ddb200.B200DMD(alg))
    if inner == "DMDPINV":
        lam, phi, Bref = oracle.dmdpinv_controlled(x[:, :-1], x[:, 1:], U[:, :-1])
    else:
        lam, phi, Bref = oracle.dmdsvd_controlled(x[:, :-1], x[:, 1:], U[:, :-1])
    Kref = np.real(oracle.operator_matrix(lam, phi))
    np.testing.assert_allclose(np.real(res.K), Kref, atol=1e-8)
    np.testing.assert_allclose(res.B, Bref, atol=1e-8)
    np.testing.assert_allclose(np.real(res.K), A, atol=1e-8)
    np.testing.assert_allclose(res.B, Bt, atol=1e-8)
    assert res.rss <= 1e-16 * m


def test_dmd_solve_host_entry_point():
    """``ddb_dmd_solve``: the host-buffer entry point the Julia glue binds."""
    A = np.array([[0.9, -0.2], [0.0, 0.2]])
    x = _snapshots(A, [10.0, -10.0], 10)
    ctx = ddb200.Context(0)
    P, Q, K = ctx.dmd_solve(x)
    np.testing.assert_allclose(P, x[:, :-1] @ x[:, :-1].T, rtol=1e-13)
    np.testing.assert_allclose(Q, x[:, 1:] @ x[:, :-1].T, rtol=1e-13)
    np.testing.assert_allclose(K, A, atol=1e-9)
    n, m = 130, 2001
    _, X, Y = systems.linear_snapshots(n, m, seed=1, restart=10**9)
    xs = np.concatenate([X, Y[:, -1:]], axis=1)
    P, Q, K = ctx.dmd_solve(xs)
    Kref = np.linalg.lstsq(X.T, Y.T, rcond=None)[0].T
    assert np.max(np.abs(K - Kref)) <= 1e-8 * max(1.0, np.max(np.abs(Kref)))
    ctx.close()


def test_c5_state_count_blocks_and_operator():
    """BASELINE config 5 at its own state count: n = 4096 states, 2 x 10^4 snapshots (the Gram kernel is
    sample-count independent per stage; the 4096-wide `trsm.cu` substitution is what no smaller test reaches).
    P, Q <= 1e-13 scaled against NumPy; K = Q P^-1 against the least-squares solution of the reference
    (`K = Y / X`, algorithms.jl:80-83; LAPACK gelsy = pivoted QR, as Julia's `/`) <= 1e-8."""
    import scipy.linalg as sla
    import torch

    n, m = 4096, 20_000
    A, X, Y = systems.linear_snapshots_blocked(n, m, seed=5, restart=64)
    dev = torch.device("cuda:0")
    Xt = torch.from_numpy(np.ascontiguousarray(X.T)).to(dev)
    Yt = torch.from_numpy(np.ascontiguousarray(Y.T)).to(dev)
    PQ = torch.zeros((2, n, n), dtype=torch.float64, device=dev)
    K = torch.zeros((n, n), dtype=torch.float64, device=dev)
    with ddb200.Context(0) as ctx:
        ctx.dmd_gram(Xt.data_ptr(), Yt.data_ptr(), n, m, PQ[0].data_ptr(), PQ[1].data_ptr())
        ctx.dmd_operator(PQ[0].data_ptr(), PQ[1].data_ptr(), n, K.data_ptr())
    P, Q, Kh = PQ[0].T.cpu().numpy(), PQ[1].T.cpu().numpy(), K.T.cpu().numpy()
    P0, Q0 = X @ X.T, Y @ X.T
    sc = np.sqrt(np.outer(np.diag(P0), np.diag(P0)))
    assert np.max(np.abs(P - P0) / sc) < 1e-13
    assert np.max(np.abs(Q - Q0) / sc) < 1e-13
    np.testing.assert_array_equal(P, P.T)
    Kref = sla.lstsq(X.T, Y.T, lapack_driver="gelsy", check_finite=False)[0].T
    assert np.max(np.abs(Kh - Kref)) <= 1e-8 * max(1.0, np.max(np.abs(Kref)))
    # size-independent property: K P = Q (the normal equations the operator solves)
    assert np.max(np.abs(Kh @ P0 - Q0)) <= 1e-9 * np.max(np.abs(Q0))


# ------------------------------------------------------------------------------------------------
# lifted bases and continuous problems (lib/DataDrivenDMD/src/solve.jl:3-65), output map C = Z / Ytilde (:130),
# eigen-decompositions and residual behind the C ABI
# ------------------------------------------------------------------------------------------------
def _vdp_like(m, seed=0):
    """A discrete nonlinear map whose degree-2 lifting is (nearly) closed: x1' = a x1, x2' = b x2 + c x1^2."""
    rng = np.random.default_rng(seed)
    a, b, c = 0.9, 0.5, 0.3
    cols = []
    for _ in range(m // 20 + 1):
        x = rng.uniform(-1, 1, 2)
        for _ in range(21):
            cols.append(x.copy())
            x = np.array([a * x[0], b * x[1] + c * x[0] ** 2])
    return np.stack(cols[: m + 1], axis=1)


@pytest.mark.parametrize("inner_name", ["DMDPINV", "DMDSVD"])
def test_lifted_discrete_problem(inner_name):
    x = _vdp_like(800)  # restarts every 21 snapshots: the few pairs across a restart are outliers both sides see alike
    basis = ddb200.polynomial_basis(2, 2)
    ex = basis.exponents
    inner = getattr(ddb200, inner_name)()
    res = ddb200.solve_dmd(ddb200.DiscreteDataDrivenProblem(x), ddb200.B200DMD(inner), basis=basis)
    Th = oracle.evaluate_library(ex, x[:, :-1])
    Yt = oracle.evaluate_library(ex, x[:, 1:])
    Z = x[:, 1:]
    ref = oracle.koopman_solve_lifted(getattr(oracle, inner_name)(), Th, Yt, Z)
    sc = np.sqrt(np.outer(np.diag(ref.P), np.diag(ref.P)))
    assert np.max(np.abs(res.P - ref.P) / sc) < 1e-12
    assert np.max(np.abs(res.Q - ref.Q) / sc) < 1e-12
    assert np.max(np.abs(np.real(res.K) - np.real(ref.K))) <= 1e-7 * max(1.0, np.max(np.abs(ref.K)))
    assert np.max(np.abs(res.C - ref.C)) <= 1e-7 * max(1.0, np.max(np.abs(ref.C)))
    assert res.rss == pytest.approx(ref.rss, rel=1e-6, abs=1e-12)
    assert res.nobs == Z.size
    np.testing.assert_allclose(np.sort_complex(res.eigenvalues), np.sort_complex(ref.eigenvalues), atol=1e-7)


def test_lifted_continuous_problem():
    """Continuous problem with a degree-2 lifting: Ytilde = J_basis(X) DX (solve.jl:3-34)."""
    rng = np.random.default_rng(3)
    m = 1500
    X = rng.uniform(-1, 1, (2, m))
    A = np.array([[-0.5, 1.0], [-1.0, -0.3]])
    DX = A @ X + 0.2 * np.vstack([X[0] * X[1], X[0] ** 2])
    basis = ddb200.polynomial_basis(2, 2)
    ex = basis.exponents
    res = ddb200.solve_dmd(ddb200.ContinuousDataDrivenProblem(X, DX=DX), ddb200.B200DMD(ddb200.DMDPINV()), basis=basis)
    Th = oracle.evaluate_library(ex, X)
    Yt = oracle.monomial_jacobian_times(ex, X, DX)
    ref = oracle.koopman_solve_lifted(oracle.DMDPINV(), Th, Yt, DX)
    sc = np.sqrt(np.outer(np.diag(ref.P), np.diag(ref.P)))
    assert np.max(np.abs(res.P - ref.P) / sc) < 1e-12
    assert np.max(np.abs(res.Q - ref.Q)) < 1e-11 * np.max(np.abs(ref.Q))
    assert np.max(np.abs(np.real(res.K) - np.real(ref.K))) <= 1e-7 * max(1.0, np.max(np.abs(ref.K)))
    assert np.max(np.abs(res.C - ref.C)) <= 1e-6 * max(1.0, np.max(np.abs(ref.C)))
    assert res.rss == pytest.approx(ref.rss, rel=1e-5, abs=1e-10)


def test_continuous_identity_lifting():
    """``solve(ContinuousDataDrivenProblem, DMDPINV())``: Theta = X, Ytilde = Z = DX -> the generator A."""
    rng = np.random.default_rng(4)
    A = np.array([[-0.1, 2.0, 0.0, 0.0], [-2.0, -0.1, 0.0, 0.0], [0.0, 0.0, -0.3, 0.5], [0.0, 0.0, -0.5, -0.3]])
    X = rng.standard_normal((4, 900))
    DX = A @ X
    res = ddb200.solve_dmd(ddb200.ContinuousDataDrivenProblem(X, DX=DX), ddb200.B200DMD(ddb200.DMDPINV()))
    np.testing.assert_allclose(np.real(res.K), A, atol=1e-9)
    assert res.rss <= 1e-18 * DX.size * 100 and res.nobs == DX.size
    np.testing.assert_allclose(np.sort_complex(res.eigenvalues), np.sort_complex(np.linalg.eigvals(A)), atol=1e-9)


def test_eig_and_residual_entry_points():
    """``ddb_dmd_eig_sym`` / ``ddb_dmd_eig`` / ``ddb_dmd_residual`` / ``ddb_gemm_nt`` directly (what a Julia caller binds)."""
    import torch

    rng = np.random.default_rng(8)
    dev = torch.device("cuda:0")
    n, m = 130, 700
    B = rng.standard_normal((n, n))
    Psym = B @ B.T
    K = rng.standard_normal((n, n)) / np.sqrt(n)
    X = rng.standard_normal((n, m))
    Zm = K @ X + 1e-3 * rng.standard_normal((n, m))
    with ddb200.Context(0) as ctx:
        A = torch.from_numpy(Psym.copy()).to(dev)
        w = torch.empty(n, dtype=torch.float64, device=dev)
        ctx.dmd_eig_sym(A.data_ptr(), n, w.data_ptr())
        U = A.T.cpu().numpy()
        np.testing.assert_allclose(w.cpu().numpy(), np.linalg.eigvalsh(Psym), rtol=1e-10)
        np.testing.assert_allclose((U * w.cpu().numpy()) @ U.T, Psym, rtol=0, atol=1e-9 * np.abs(Psym).max())
        Kc = torch.from_numpy(np.ascontiguousarray(K.T)).to(dev)  # column-major K
        wc = torch.empty(2 * n, dtype=torch.float64, device=dev)
        vr = torch.empty((n, n), dtype=torch.float64, device=dev)
        Kwork = Kc.clone()  # the input is destroyed
        ctx.dmd_eig(Kwork.data_ptr(), n, wc.data_ptr(), vr.data_ptr())
        lam = wc.cpu().numpy()[0::2] + 1j * wc.cpu().numpy()[1::2]
        np.testing.assert_allclose(np.sort_complex(lam), np.sort_complex(np.linalg.eigvals(K)), atol=1e-9)
        V = vr.T.cpu().numpy()
        j = int(np.argmax(lam.imag > 0))  # a complex pair: columns Re, Im
        v = V[:, j] + 1j * V[:, j + 1]
        assert np.linalg.norm(K @ v - lam[j] * v) <= 1e-9 * np.linalg.norm(v)
        Xd = torch.from_numpy(np.ascontiguousarray(X.T)).to(dev)
        Zd = torch.from_numpy(np.ascontiguousarray(Zm.T)).to(dev)
        rss = ctx.dmd_residual(Xd.data_ptr(), Zd.data_ptr(), Kc.data_ptr(), n, n, m)
        assert rss == pytest.approx(float(np.sum((Zm - K @ X) ** 2)), rel=1e-10)
        D = torch.empty((m, n), dtype=torch.float64, device=dev)
        Kr = torch.from_numpy(np.ascontiguousarray(K)).to(dev)
        ctx.gemm_nt(D.data_ptr(), n, Xd.data_ptr(), n, Kr.data_ptr(), n, m, n, n)  # D = X' K'  (row-major)
        np.testing.assert_allclose(D.cpu().numpy(), (K @ X).T, rtol=0, atol=1e-12 * np.abs(K @ X).max() + 1e-13)
