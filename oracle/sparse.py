"""Oracle: DataDrivenSparse's thresholded sweep (test infrastructure only).

Restates, line by line, ``lib/DataDrivenSparse/src``:
``solver.jl:59-118`` (``SparseLinearSolver``), ``algorithms/STLSQ.jl:77-140``,
``algorithms/ADMM.jl:64-139``, ``algorithms/SR3.jl:45-141``, ``algorithms/proximals.jl:1-195``,
``DataDrivenSparse.jl:144-201`` (cache bookkeeping + StatsAPI statistics),
``DataDrivenSparse.jl:258-277`` (algorithm callable), ``commonsolve.jl:26-56``
(``__sparse_regression``) and ``src/utils/build_basis.jl:148-157`` (coefficient rounding).

Row-vector convention of the reference: a cache holds ``X`` as a 1 x k row; here 1-D arrays of
length k.  ``A`` is the library matrix Theta (k x m), ``b`` one target row (m).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np
import scipy.linalg as sla

EPS = float(np.finfo(np.float64).eps)


# ----------------------------------------------------------------------------------------------
# Julia's `B / A` (LinearAlgebra stdlib, un-vendored; see oracle/__init__.py)
# ----------------------------------------------------------------------------------------------
def _ldiv(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """Julia ``A \\ B`` for a dense matrix: triangular solve / LU when square, otherwise
    column-pivoted QR with ``rcond = min(size(A)) * eps`` and the minimum-norm solution
    (``LinearAlgebra.\\`` -> ``qr(A, ColumnNorm()) \\ B``; LAPACK ``geqp3`` + incremental condition
    estimation, i.e. ``gelsy``)."""
    rows, cols = A.shape
    if rows == 0 or cols == 0:
        return np.zeros((cols,) + B.shape[1:], dtype=np.float64)
    if rows == cols:
        if np.array_equal(A, np.tril(A)):
            return sla.solve_triangular(A, B, lower=True, check_finite=False)
        if np.array_equal(A, np.triu(A)):
            return sla.solve_triangular(A, B, lower=False, check_finite=False)
        lu, piv = sla.lu_factor(A, check_finite=False)
        return sla.lu_solve((lu, piv), B, check_finite=False)
    x, _, _, _ = sla.lstsq(A, B, cond=min(rows, cols) * EPS, lapack_driver="gelsy", check_finite=False)
    return x


def rdiv(B: np.ndarray, A: np.ndarray) -> np.ndarray:
    """Julia ``B / A`` with ``B`` a 1 x q row (given 1-D) and ``A`` p x q: returns the 1 x p row
    ``x`` with ``x A = B`` in the least-squares / minimum-norm sense (``/(B, A) = (A' \\ B')'``)."""
    return _ldiv(np.ascontiguousarray(A.T), np.asarray(B, dtype=np.float64))


# ----------------------------------------------------------------------------------------------
# Proximal operators (algorithms/proximals.jl)
# ----------------------------------------------------------------------------------------------
class SoftThreshold:
    """``proximals.jl:36-65``"""

    def active_set(self, x: np.ndarray, lam: float) -> np.ndarray:
        return np.abs(x) > lam  # :38-47 (strict)

    def __call__(self, x: np.ndarray, lam: float) -> np.ndarray:
        return np.sign(x) * np.maximum(np.abs(x) - lam, 0.0)  # :49-65


class HardThreshold:
    """``proximals.jl:88-117``; cut at ``sqrt(2*lambda)``"""

    def active_set(self, x: np.ndarray, lam: float) -> np.ndarray:
        return np.abs(x) > math.sqrt(2 * lam)  # :90-99

    def __call__(self, x: np.ndarray, lam: float) -> np.ndarray:
        return np.where(np.abs(x) > math.sqrt(2 * lam), x, 0.0)  # :101-117


class ClippedAbsoluteDeviation:
    """``proximals.jl:156-195``; default upper cut ``rho = 5*lambda`` (``:169,178``)"""

    def __init__(self, rho: float = float("nan")):
        self.rho = rho

    def _rho(self, lam: float) -> float:
        return 5.0 * lam if math.isnan(self.rho) else float(self.rho)

    def active_set(self, x: np.ndarray, lam: float) -> np.ndarray:
        return np.abs(x) > self._rho(lam)

    def __call__(self, x: np.ndarray, lam: float) -> np.ndarray:
        soft = np.sign(x) * np.maximum(np.abs(x) - lam, 0.0)
        return np.where(np.abs(x) > self._rho(lam), x, soft)


def apply_active_set(prox, x: np.ndarray, lam: float) -> np.ndarray:
    """``_apply_active_set!`` (``proximals.jl:1-13``): recompute the mask, zero everything outside
    it.  Mutates ``x``; returns the mask."""
    mask = prox.active_set(x, lam)
    x[~mask] = 0.0
    return mask


# ----------------------------------------------------------------------------------------------
# Algorithms and their caches
# ----------------------------------------------------------------------------------------------
class _Cache:
    """Common fields ``X, X_prev, active_set`` + original data ``Atilde, Btilde``."""

    X: np.ndarray
    X_prev: np.ndarray
    active_set: np.ndarray
    Atilde: np.ndarray
    Btilde: np.ndarray

    def set_from(self, other: "_Cache") -> None:  # _set!  DataDrivenSparse.jl:144-153
        self.X[:] = other.X
        self.X_prev[:] = other.X_prev
        self.active_set[:] = other.active_set

    def zero(self) -> None:  # _zero!  DataDrivenSparse.jl:155-158 (only X)
        self.X[:] = 0.0

    def is_converged(self, abstol: float, reltol: float) -> bool:  # DataDrivenSparse.jl:160-168
        if not self.active_set.any():
            return True
        delta = float(np.linalg.norm(self.X - self.X_prev))
        if delta < abstol:
            return True
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = delta / float(np.linalg.norm(self.X))
        return bool(rel < reltol)


def _thresholds(t) -> np.ndarray:
    return np.atleast_1d(np.asarray(t, dtype=np.float64)).copy()


class STLSQ:
    """``algorithms/STLSQ.jl:48-60``"""

    name = "STLSQ"

    def __init__(self, thresholds=1e-1, rho: float = 0.0):
        self.thresholds = _thresholds(thresholds)
        assert np.all(self.thresholds > 0), "Threshold must be positive definite"
        assert rho >= 0, "Ridge regression parameter must be positive definite!"
        self.rho = float(rho)
        self.proximal = SoftThreshold()  # get_proximal default, DataDrivenSparse.jl:232

    def init_cache(self, A: np.ndarray, b: np.ndarray) -> "_STLSQCache":
        return _STLSQCache(self, A, b)


class _STLSQCache(_Cache):
    def __init__(self, alg: STLSQ, A: np.ndarray, b: np.ndarray):  # STLSQ.jl:81-115
        k, m = A.shape
        lam = float(np.min(alg.thresholds))
        self.proximal = alg.proximal
        if k <= m and alg.rho != 0.0:  # :89-92
            self.A = A @ A.T + alg.rho * np.eye(k)
            self.B = b @ A.T
            self.usenormal = True
        else:  # :93-97
            self.A = A
            self.B = b
            self.usenormal = False
        self.X = rdiv(self.B, self.A)  # :99
        self.X_prev = np.zeros_like(self.X)  # :101
        self.active_set = self.proximal.active_set(self.X, lam)  # :103-105
        self.Atilde, self.Btilde = A, b

    def step(self, lam: float) -> None:  # STLSQ.jl:117-140
        self.X_prev[:] = self.X
        p = self.active_set
        if self.usenormal:
            self.X[p] = rdiv(self.B[p], self.A[np.ix_(p, p)])  # :128-133
        else:
            self.X[p] = rdiv(self.B, self.A[p, :])  # :135-140
        self.active_set[:] = apply_active_set(self.proximal, self.X, lam)  # :124


class ADMM:
    """``algorithms/ADMM.jl:32-43``"""

    name = "ADMM"

    def __init__(self, thresholds=1e-1, rho: float = 1.0):
        self.thresholds = _thresholds(thresholds)
        assert np.all(self.thresholds > 0)
        assert rho > 0
        self.rho = float(rho)
        self.proximal = SoftThreshold()

    def init_cache(self, A, b):
        return _ADMMCache(self, A, b)


class _ADMMCache(_Cache):
    def __init__(self, alg: ADMM, A: np.ndarray, b: np.ndarray):  # ADMM.jl:68-104
        k, m = A.shape
        lam = float(np.min(alg.thresholds))
        rho = alg.rho
        self.rho = rho
        self.B = b @ A.T  # :77
        if k <= m:  # :78-81
            self.F = sla.cho_factor(A @ A.T + rho * np.eye(k), lower=True, check_finite=False)
            self.fat = False
            self.X = sla.cho_solve(self.F, self.B, check_finite=False)
        else:  # :82-86
            self.F = sla.cho_factor(A.T @ A / rho + np.eye(m), lower=True, check_finite=False)
            self.fat = True
            self.X = rdiv(b, A)
        self.proximal = SoftThreshold()
        self.active_set = self.proximal.active_set(self.X, lam / rho)  # :92
        self.X_prev = np.zeros_like(self.X)
        self.alpha = np.zeros_like(self.X)
        self.w = np.zeros_like(self.X)
        self.Atilde, self.Btilde = A, b

    def step(self, lam: float) -> None:
        rho = self.rho
        self.X_prev[:] = self.X
        if not self.fat:  # ADMM.jl:107-121
            self.X[:] = sla.cho_solve(self.F, self.B + rho * (self.alpha - self.w), check_finite=False)
        else:  # ADMM.jl:123-139
            q = self.B + rho * (self.alpha - self.w)
            bb = (sla.cho_solve(self.F, q @ self.Atilde, check_finite=False) @ self.Atilde.T) / rho**2
            self.X[:] = q / rho - bb
        self.alpha[:] = self.proximal(self.X + self.w, lam / rho)
        self.w += self.X - self.alpha
        self.active_set[:] = apply_active_set(self.proximal, self.X, lam / rho)


class SR3:
    """``algorithms/SR3.jl:45-73``.  With ``HardThreshold`` the stored thresholds are
    ``threshold^2 / 2`` (``:63``)."""

    name = "SR3"

    def __init__(self, thresholds=1e-1, nu: float = 1.0, proximal=None):
        proximal = HardThreshold() if proximal is None else proximal
        t = _thresholds(thresholds)
        assert np.all(t > 0)
        assert nu > 0
        self.thresholds = t**2 / 2 if isinstance(proximal, HardThreshold) else t
        self.nu = float(nu)
        self.proximal = proximal

    def init_cache(self, A, b):
        return _SR3Cache(self, A, b)


class _SR3Cache(_Cache):
    def __init__(self, alg: SR3, A: np.ndarray, b: np.ndarray):  # SR3.jl:98-127
        k, m = A.shape
        lam = float(np.min(alg.thresholds))
        self.nu = alg.nu
        self.proximal = alg.proximal
        self.F = sla.cho_factor(A @ A.T + np.eye(k) * alg.nu, lower=True, check_finite=False)  # :108
        self.B = b @ A.T  # :109
        self.X = sla.cho_solve(self.F, self.B, check_finite=False)  # :111
        self.active_set = self.proximal.active_set(self.X, lam)  # :115
        self.X_prev = self.X.copy()  # :122
        self.W = np.zeros_like(self.X)
        self.Atilde, self.Btilde = A, b

    def step(self, lam: float) -> None:  # SR3.jl:129-141
        self.X_prev[:] = self.X
        self.W[:] = sla.cho_solve(self.F, self.B + self.X * self.nu, check_finite=False)
        self.X[:] = self.proximal(self.W, lam)
        self.active_set[:] = self.proximal.active_set(self.X, lam)


# ----------------------------------------------------------------------------------------------
# StatsAPI statistics on a cache (DataDrivenSparse.jl:171-201) and the information criteria
# ----------------------------------------------------------------------------------------------
def rss(c: _Cache) -> float:  # :173-176
    r = c.X @ c.Atilde - c.Btilde
    return float(np.sum(r * r))


def dof(c: _Cache) -> int:  # :178-181
    return int(np.sum(c.active_set))


def nobs(c: _Cache) -> int:  # :183-186
    return int(c.Btilde.size)


def loglikelihood(c: _Cache) -> float:  # :188-192
    n = nobs(c)
    with np.errstate(divide="ignore"):
        return float(-n / 2 * np.log(np.float64(rss(c)) / n))


def nullloglikelihood(c: _Cache) -> float:  # :194-199
    n = nobs(c)
    with np.errstate(divide="ignore"):
        return float(-n / 2 * np.log(np.mean((c.Btilde - np.mean(c.Btilde)) ** 2)))


def r2(c: _Cache) -> float:  # :201 -> StatsAPI r2(:CoxSnell) = 1 - exp(2(nll - ll)/n)
    return 1.0 - math.exp(2.0 * (nullloglikelihood(c) - loglikelihood(c)) / nobs(c))


def aic(c: _Cache) -> float:  # StatsAPI: -2ll + 2k
    return -2.0 * loglikelihood(c) + 2.0 * dof(c)


def aicc(c: _Cache) -> float:  # StatsAPI: aic + 2k(k+1)/(n-k-1)
    k, n = dof(c), nobs(c)
    with np.errstate(divide="ignore"):
        return float(-2.0 * loglikelihood(c) + 2.0 * k + np.float64(2.0 * k * (k + 1)) / np.float64(n - k - 1))


def bic(c: _Cache) -> float:  # StatsAPI: -2ll + k log n
    return -2.0 * loglikelihood(c) + dof(c) * math.log(nobs(c))


@dataclass
class CommonOptions:
    """The fields of ``DataDrivenCommonOptions`` the path reads
    (``src/utils/common_options.jl:22-54``)."""

    maxiters: int = 100
    abstol: float = math.sqrt(EPS)
    reltol: float = math.sqrt(EPS)
    digits: int = 10
    selector: Callable[[_Cache], float] = bic


@dataclass
class SweepTrace:
    """Everything the parity tests compare, recorded while the reference loop runs."""

    # per threshold: support and coefficients of the cache when the inner loop stops
    support_path: list = field(default_factory=list)
    coef_path: list = field(default_factory=list)
    iters_path: list = field(default_factory=list)
    # per step: (j, iter, dof, rss, selected)
    steps: list = field(default_factory=list)


class SparseLinearSolver:
    """``solver.jl:26-118``"""

    def __init__(self, algorithm, options: Optional[CommonOptions] = None):
        self.algorithm = algorithm
        self.options = options or CommonOptions()

    def solve_target(self, A: np.ndarray, b: np.ndarray, trace: Optional[SweepTrace] = None):
        """``(alg::SparseLinearSolver)(X, Y::AbstractVector)`` (``solver.jl:72-118``).
        Returns ``(best_cache, optimal_threshold, iteration_counter)``."""
        alg, opt = self.algorithm, self.options
        thresholds = alg.thresholds
        if np.any(np.diff(thresholds) < 0):  # :77-79
            thresholds.sort()
        cache = alg.init_cache(A, b)  # :81
        best = alg.init_cache(A, b)  # :82
        best.zero()  # :83
        optimal_threshold = float(np.min(thresholds))
        iteration_counter = 0
        for j, lam in enumerate(thresholds):  # :94
            it = 0
            for it in range(1, opt.maxiters + 1):  # :95
                iteration_counter += 1
                cache.step(float(lam))  # :98
                selected = (opt.selector(cache) <= opt.selector(best)) or j == 0  # :100
                if selected:
                    best.set_from(cache)
                    optimal_threshold = float(lam)
                if trace is not None:
                    trace.steps.append((j, it, dof(cache), rss(cache), bool(selected)))
                if cache.is_converged(opt.abstol, opt.reltol):  # :113
                    break
            if trace is not None:
                trace.support_path.append(cache.active_set.copy())
                trace.coef_path.append(cache.X.copy())
                trace.iters_path.append(it)
        return best, optimal_threshold, iteration_counter

    def __call__(self, A: np.ndarray, Y: np.ndarray, traces: Optional[list] = None):
        """``(alg::SparseLinearSolver)(X, Y::AbstractMatrix)`` (``solver.jl:59-70``): sequential
        map over target rows."""
        Y = np.atleast_2d(Y)
        out = []
        for i in range(Y.shape[0]):
            tr = SweepTrace() if traces is not None else None
            out.append(self.solve_target(A, Y[i], tr))
            if traces is not None:
                traces.append(tr)
        return out


def run_algorithm(alg, A: np.ndarray, Y: np.ndarray, options: Optional[CommonOptions] = None, traces=None):
    """The algorithm callable (``DataDrivenSparse.jl:258-277``): returns
    ``(coeff_matrix n_t x k, optimal_thresholds, optimal_iterations)``."""
    Y = np.atleast_2d(Y)
    results = SparseLinearSolver(alg, options)(A, Y, traces)
    C = np.zeros((Y.shape[0], A.shape[0]))
    lams, iters = [], []
    for i, (c, lam, it) in enumerate(results):
        C[i] = c.X
        lams.append(lam)
        iters.append(it)
    return C, lams, iters


@dataclass
class SparseRegressionResult:
    """``result.jl:1-17``"""

    coefficients: np.ndarray
    dof: int
    lambdas: list
    iterations: list
    testerror: Optional[float]
    trainerror: float
    retcode: int = 1


def sparse_regression(alg, Theta: np.ndarray, Y: np.ndarray, options: Optional[CommonOptions] = None, traces=None):
    """``__sparse_regression`` (``lib/DataDrivenSparse/src/commonsolve.jl:26-56``) for the default
    single batch with no test split."""
    Y = np.atleast_2d(Y)
    C, lams, iters = run_algorithm(alg, Theta, Y, options, traces)
    trainerror = float(np.sum((Y - C @ Theta) ** 2))  # :37
    d = int(np.sum(np.abs(C) > 0.0))  # :47
    return SparseRegressionResult(C, d, lams, iters, None, trainerror, 1)


def construct_coefficients(C: np.ndarray, digits: int = 10) -> np.ndarray:
    """Coefficient post-processing of ``__construct_basis`` (``src/utils/build_basis.jl:154-157``):
    ``round.(X, RoundToZero, digits = digits)`` then zero everything with ``abs <= eps``."""
    og = 10.0**digits
    out = np.trunc(np.asarray(C, dtype=np.float64) * og) / og
    out[np.abs(out) <= EPS] = 0.0
    return out


# ----------------------------------------------------------------------------------------------
# ImplicitOptimizer (lib/DataDrivenSparse/src/algorithms/Implicit.jl:37-108) and its driver
# (lib/DataDrivenSparse/src/commonsolve.jl:58-114)
# ----------------------------------------------------------------------------------------------
class ImplicitOptimizer:
    """``ImplicitOptimizer(opt)`` (``Implicit.jl:37-51``): wraps an explicit optimizer; every library row
    in turn is the left-hand side, regressed on the remaining rows."""

    name = "ImplicitOptimizer"

    def __init__(self, optimizer=None, threshold=1e-1):
        self.optimizer = optimizer if optimizer is not None else STLSQ(threshold)

    @property
    def thresholds(self):
        return self.optimizer.thresholds

    def __call__(self, X: np.ndarray, Y=None, options: Optional[CommonOptions] = None, necessary_idx=None, results_out=None):
        """``(x::ImplicitOptimizer)(X, Y; options, necessary_idx)`` (``Implicit.jl:57-108``).  ``Y`` is not
        used by the reference either.  Returns ``(x_opt 1 x n, optimal_threshold, optimal_iterations)``."""
        X = np.asarray(X, dtype=np.float64)
        n = X.shape[0]
        necessary_idx = np.ones(n, dtype=bool) if necessary_idx is None else np.asarray(necessary_idx, dtype=bool)
        solver = SparseLinearSolver(self.optimizer, options)
        results = []
        for i in range(n):  # :74-84
            inds = np.ones(n, dtype=bool)
            inds[i] = False
            results.append(solver.solve_target(X[inds], X[i]))
        best_id = -1
        for i, res in enumerate(results):  # :87-98
            inds = np.ones(n, dtype=bool)
            inds[i] = False
            c = res[0]
            if dof(c) >= 2 and np.any(c.X[necessary_idx[inds]] != 0.0):
                if best_id < 0 or aicc(c) < aicc(results[best_id][0]):
                    best_id = i
        if best_id < 0:
            raise IndexError("ImplicitOptimizer: no candidate with dof >= 2 touches a necessary term (Implicit.jl:100-102)")
        inds = np.ones(n, dtype=bool)
        inds[best_id] = False
        best_cache, lam, iters = results[best_id]
        x_opt = np.zeros((1, n))
        x_opt[0, best_id] = -1.0  # :105
        x_opt[0, inds] = best_cache.X  # :106
        if results_out is not None:
            results_out.extend(results)
        return x_opt, lam, iters


def implicit_candidate_matrix(implicit_idx: np.ndarray) -> np.ndarray:
    """``candidate_matrix`` of ``commonsolve.jl:65-74``: row i is a candidate for implicit variable j when it
    depends on variable j or on no OTHER implicit variable."""
    implicit_idx = np.asarray(implicit_idx, dtype=bool)
    k, ni = implicit_idx.shape
    cm = np.zeros((k, ni), dtype=bool)
    for i in range(k):
        for j in range(ni):
            others = np.ones(ni, dtype=bool)
            others[j] = False
            cm[i, j] = implicit_idx[i, j] or implicit_idx[i, others].sum() == 0
    return cm


def implicit_sparse_regression(alg: ImplicitOptimizer, Theta: np.ndarray, implicit_idx: np.ndarray,
                               options: Optional[CommonOptions] = None):
    """``__sparse_regression(ps::InternalDataDrivenProblem{<:ImplicitOptimizer}, X, Y)``
    (``commonsolve.jl:58-114``) for the default single batch without a test split."""
    Theta = np.asarray(Theta, dtype=np.float64)
    implicit_idx = np.asarray(implicit_idx, dtype=bool)
    cm = implicit_candidate_matrix(implicit_idx)
    C = np.zeros((cm.shape[1], cm.shape[0]))
    lams, iters = [], []
    for i in range(cm.shape[1]):  # :83-93
        idx = cm[:, i]
        coeff, lam, it = alg(Theta[idx], None, options=options, necessary_idx=implicit_idx[idx, i])
        C[i, idx] = coeff[0]
        lams.append(lam)
        iters.append(it)
    trainerror = float(np.sum((C @ Theta) ** 2))  # :95
    d = int(np.sum(np.abs(C) > 0.0))  # :107
    return SparseRegressionResult(C, d, lams, iters, None, trainerror, 1)


# ----------------------------------------------------------------------------------------------
# init + solve! with the pre-processing options (src/commonsolve.jl:92-139, src/utils/data_processing.jl:29-126,
# lib/DataDrivenSparse/src/commonsolve.jl:1-24)
# ----------------------------------------------------------------------------------------------
def zscore_fit(Theta: np.ndarray) -> np.ndarray:
    """``fit(DataNormalization{ZScoreTransform}, data)`` (``data_processing.jl:75-80``): per-row standard
    deviation (``center = false`` only suppresses the subtraction of the mean when the transform is APPLIED),
    zero scales replaced by one (constants)."""
    s = np.std(Theta, axis=1, ddof=1)
    s[s == 0.0] = 1.0
    return s


def solve_processed(alg, Theta: np.ndarray, Y: np.ndarray, options: Optional[CommonOptions] = None, normalize=None,
                    split: float = 1.0, batchsize: int = 0, partial: bool = True, perm=None):
    """Returns ``(coefficients after apply_transform, results sorted by l2error, sigma)``.  ``perm``: order of
    the training samples (the DataLoader's shuffle; ``None`` keeps the order)."""
    Theta = np.array(Theta, dtype=np.float64)
    Y = np.atleast_2d(np.asarray(Y, dtype=np.float64))
    m = Theta.shape[1]
    sigma = None
    if normalize == "ZScoreTransform":  # commonsolve.jl:128-130: fitted and applied on ALL samples before the split
        sigma = zscore_fit(Theta)
        Theta = Theta / sigma[:, None]
    elif normalize == "UnitRangeTransform":
        # fit(UnitRangeTransform, Theta, dims = 2) (data_processing.jl:68-73): per-row minimum and 1 / (max - min), an
        # infinite scale (constant row) replaced by 1; apply_transform! (:119-126): (Theta - min) .* scale
        tmin = Theta.min(axis=1)
        rng = Theta.max(axis=1) - tmin
        with np.errstate(divide="ignore"):
            tscale = 1.0 / rng
        tscale[np.isinf(tscale)] = 1.0
        Theta = (Theta - tmin[:, None]) * tscale[:, None]
        sigma = (tmin, tscale)
    elif normalize is not None:
        raise ValueError(normalize)
    split = min(max(float(split), 0.0), 1.0)  # data_processing.jl:32
    m_train = int(round(split * m))  # MLUtils.splitobs(at = split, shuffle = false)
    Tt, Yt = Theta[:, m_train:], Y[:, m_train:]
    order = np.arange(m_train) if perm is None else np.asarray(perm)
    bs = m_train if batchsize <= 0 else batchsize  # :36
    starts = list(range(0, m_train, bs))
    if not partial and m_train % bs:
        starts = starts[:-1]
    results = []
    for b0 in starts:
        idx = order[b0:b0 + bs]
        Tb, Yb = Theta[:, idx], Y[:, idx]
        C, lams, iters = run_algorithm(alg, Tb, Yb, options)
        trainerror = float(np.sum((Yb - C @ Tb) ** 2))  # lib commonsolve.jl:37
        testerror = float(np.sum((Yt - C @ Tt) ** 2)) if Tt.shape[1] else None  # :41-45
        results.append(SparseRegressionResult(C, int(np.sum(np.abs(C) > 0.0)), lams, iters, testerror, trainerror, 1))
    results.sort(key=lambda r: r.testerror if r.testerror is not None else r.trainerror)  # :12 sort!(results, by = l2error)
    best = results[0]
    C = best.coefficients.copy()
    if isinstance(sigma, tuple):
        # lib commonsolve.jl:19-21: the SAME apply_transform! is run on the transposed coefficient matrix -- the rows
        # of the coefficients are shifted by the library minima and multiplied by the scales (the reference as written)
        C = (C - sigma[0][None, :]) * sigma[1][None, :]
    elif sigma is not None:
        C = C / sigma[None, :]  # :19-21 apply_transform(transform, coefficients')' with an empty mean
    return C, results, sigma
