"""Oracle: DataDrivenDMD operators for a discrete problem (test infrastructure only).

Restates ``lib/DataDrivenDMD/src/algorithms.jl:2-20`` (``truncated_svd``), ``:80-83`` (``DMDPINV``),
``:147-158`` (``DMDSVD``), the driver ``solve.jl:102-133`` (``Q = Y X'``, ``P = X X'``,
``C = Z / Y``) and ``result.jl:46-59`` (``KoopmanResult`` statistics), for the identity lifting
used by ``solve(prob, alg)`` (``src/commonsolve.jl:84-90`` -> ``unit_basis``) on a
``DiscreteDataDrivenProblem`` (``solve.jl:36-65``: X = x[:, 1:end-1], Y = Z = x[:, 2:end]).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.linalg as sla

from .sparse import _ldiv


def truncated_svd(A: np.ndarray, truncation):
    """``algorithms.jl:2-20``: real truncation keeps ``S > trunc * max(S)``; integer truncation keeps
    the first ``trunc`` strictly positive singular values."""
    U, S, Vt = sla.svd(A, full_matrices=False, lapack_driver="gesdd", check_finite=False)
    if isinstance(truncation, (int, np.integer)) and not isinstance(truncation, bool):
        r = np.array([(i < truncation and S[i] > 0.0) for i in range(len(S))])
    else:
        t = min(float(truncation), 1.0)
        r = S > t * np.max(S)
    return U[:, r], S[r], Vt[r, :].T


class DMDPINV:
    """``K = Y / X`` then ``eigen(K)`` (``algorithms.jl:77-83``)"""

    def __call__(self, X: np.ndarray, Y: np.ndarray):
        K = _ldiv(np.ascontiguousarray(X.T), np.ascontiguousarray(Y.T)).T
        lam, phi = sla.eig(K)
        return lam, phi


class DMDSVD:
    """Exact DMD (``algorithms.jl:140-158``)"""

    def __init__(self, truncation=0.0):
        self.truncation = truncation

    def __call__(self, X: np.ndarray, Y: np.ndarray):
        U, S, V = truncated_svd(X, self.truncation)
        B = (Y @ V) * (1.0 / S)[None, :]
        At = U.T @ B
        lam, om = sla.eig(At)
        phi = B @ om
        return lam, phi


def operator_matrix(lam: np.ndarray, phi: np.ndarray) -> np.ndarray:
    """``Matrix(::Eigen) = vectors * Diagonal(values) * inv(vectors)``; real part when the
    imaginary residue is rounding noise (the reference keeps the complex matrix only when the
    eigen-decomposition is genuinely complex)."""
    if phi.shape[0] == phi.shape[1]:
        K = (phi * lam[None, :]) @ np.linalg.inv(phi)
    else:  # rank-truncated modes: pseudo-inverse
        K = (phi * lam[None, :]) @ np.linalg.pinv(phi)
    return K


@dataclass
class KoopmanResult:
    """``result.jl:20-60``"""

    eigenvalues: np.ndarray
    modes: np.ndarray
    K: np.ndarray
    C: np.ndarray
    Q: np.ndarray
    P: np.ndarray
    rss: float
    dof: int
    nobs: int
    loglikelihood: float


def koopman_solve(alg, x: np.ndarray) -> KoopmanResult:
    """Driver of ``solve.jl:102-133`` for snapshot data ``x`` (n x (m+1)), identity lifting, no
    controls, single batch."""
    X, Y = x[:, :-1], x[:, 1:]
    Z = Y
    lam, phi = alg(X, Y)
    K = operator_matrix(lam, phi)
    Q = Y @ X.T  # :128
    P = X @ X.T  # :129
    C = _ldiv(np.ascontiguousarray(Y.T), np.ascontiguousarray(Z.T)).T  # :130  C = Z / Y_
    Kr = np.real(K) if np.max(np.abs(np.imag(K))) <= 1e-12 * max(1.0, np.max(np.abs(K))) else K
    R = Z - C @ Kr @ X
    rss = float(np.sum(np.abs(R) ** 2))  # result.jl:51
    d = int(np.sum(Kr != 0))  # :52
    n = int(Z.size)
    with np.errstate(divide="ignore"):
        ll = float(-n / 2 * np.log(np.float64(rss) / n))
    return KoopmanResult(lam, phi, Kr, C, Q, P, rss, d, n, ll)


def koopman_solve_lifted(alg, Theta: np.ndarray, Ytilde: np.ndarray, Z: np.ndarray) -> KoopmanResult:
    """The same driver (``solve.jl:102-133``) on lifted data as ``get_fit_targets`` hands it over
    (``solve.jl:3-65``): ``Theta = basis(X)``, ``Ytilde = basis(X shifted)`` (discrete) or ``J_basis(X) DX``
    (continuous), ``Z`` = the raw targets."""
    lam, phi = alg(Theta, Ytilde)
    K = operator_matrix(lam, phi)
    Q = Ytilde @ Theta.T
    P = Theta @ Theta.T
    C = _ldiv(np.ascontiguousarray(Ytilde.T), np.ascontiguousarray(Z.T)).T  # :130  C = Z / Y_
    Kr = np.real(K) if np.max(np.abs(np.imag(K))) <= 1e-9 * max(1.0, np.max(np.abs(K))) else K
    R = Z - C @ Kr @ Theta
    rss = float(np.sum(np.abs(R) ** 2))
    n = int(Z.size)
    with np.errstate(divide="ignore"):
        ll = float(-n / 2 * np.log(np.float64(rss) / n))
    return KoopmanResult(lam, phi, Kr, C, Q, P, rss, int(np.sum(Kr != 0)), n, ll)


def monomial_jacobian_times(exps: np.ndarray, X: np.ndarray, DX: np.ndarray) -> np.ndarray:
    """``jacobian(basis)(x) * dx`` for a monomial library (``solve.jl:17-32``), column by column."""
    n, k = exps.shape
    out = np.zeros((k, X.shape[1]))
    for j in range(k):
        for i in range(n):
            if exps[i, j] == 0:
                continue
            e = exps[:, j].astype(np.int64).copy()
            e[i] -= 1
            out[j] += exps[i, j] * np.prod(X ** e[:, None], axis=0) * DX[i]
    return out


# ----------------------------------------------------------------------------------------------
# variants (lib/DataDrivenDMD/src/algorithms.jl:84-93, 160-176, 215-235, 276-292)
# ----------------------------------------------------------------------------------------------
def dmdpinv_controlled(X, Y, U):
    """``(x::DMDPINV)(X, Y, U)`` (``algorithms.jl:84-93``): ``Ktilde = Y / [X; U]``; returns (lam, phi, B)."""
    nx = X.shape[0]
    XU = np.vstack([X, U])
    Kt = _ldiv(np.ascontiguousarray(XU.T), np.ascontiguousarray(Y.T)).T
    lam, phi = sla.eig(Kt[:, :nx])
    return lam, phi, Kt[:, nx:]


def dmdsvd_controlled(X, Y, U, truncation=0.0):
    """``(x::DMDSVD)(X, Y, U)`` (``algorithms.jl:160-176``, DMDc)."""
    nx = X.shape[0]
    Ut, St, Vt = truncated_svd(np.vstack([X, U]), truncation)
    Uh = sla.svd(Y, full_matrices=False, lapack_driver="gesdd")[0]
    U1, U2 = Ut[:nx], Ut[nx:]
    C = (Y @ Vt) * (1.0 / St)[None, :]
    At = Uh.T @ C @ U1.T @ Uh
    Bt = C @ U2.T
    lam, om = sla.eig(At)
    phi = C @ U1.T @ Uh @ om
    return lam, phi, Bt


class TOTALDMD:
    """``TOTALDMD(truncation, alg)`` (``algorithms.jl:215-235``): project on the right singular subspace of
    ``[X; Y]``, then run ``alg``."""

    def __init__(self, truncation=0.0, alg=None):
        self.truncation, self.alg = truncation, alg if alg is not None else DMDPINV()

    def __call__(self, X, Y):
        _, _, Q = truncated_svd(np.vstack([X, Y]), self.truncation)
        return self.alg(X @ Q, Y @ Q)


class FBDMD:
    """``FBDMD(truncation)`` (``algorithms.jl:276-292``): ``sqrt(A1 * inv(A2))`` of the forward and backward
    exact-DMD operators, signs taken from ``A1``."""

    def __init__(self, truncation=0.0):
        self.alg = DMDSVD(truncation)

    def __call__(self, X, Y):
        A1 = operator_matrix(*self.alg(X, Y))
        A2 = operator_matrix(*self.alg(Y, X))
        A1 = np.real_if_close(A1, tol=1e6)
        At = sla.sqrtm(A1 @ np.linalg.inv(A2))
        At = np.abs(At) * np.sign(np.real(A1))
        return sla.eig(At)
