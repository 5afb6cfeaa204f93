"""Host-side mirror of the reference's plug-in interface for the sparse-identification path.

The reference is Julia; its toolchain is absent from the build image and the GPU box, so the host
side above the C ABI is written in Python with the reference's names, argument meaning and error
behaviour, and the Julia glue (``julia/DataDrivenB200.jl``) mirrors this file one to one:

=============================  ================================================================
here                           reference
=============================  ================================================================
``Basis``, ``polynomial_basis``  ``src/basis/type.jl:50-89``, ``src/utils/basis_generators.jl:84-149``
``ContinuousDataDrivenProblem``  ``src/problem/type.jl:92-113`` (+ ``get_implicit_data`` ``:433-435``)
``DataDrivenCommonOptions``      ``src/utils/common_options.jl:22-54``
``STLSQ`` / ``ADMM`` / ``SR3``   ``lib/DataDrivenSparse/src/algorithms/*.jl``
``B200Sparse``                   the new ``AbstractSparseRegressionAlgorithm`` subtype
``get_fit_targets``              ``src/commonsolve.jl:74-81`` (returns raw (X, Y): Theta is never built)
``solve``                        ``solve!`` ``lib/DataDrivenSparse/src/commonsolve.jl:1-56``
``SparseRegressionResult``       ``lib/DataDrivenSparse/src/result.jl:1-26``
``construct_coefficients``       ``src/utils/build_basis.jl:154-157``
=============================  ================================================================

Options the backend cannot honour raise ``ValueError`` (the reference's ``ArgumentError``); nothing
falls back to a CPU path.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .backend import Context, SweepOutput

EPS = float(np.finfo(np.float64).eps)


# ----------------------------------------------------------------------------------------------
# Basis (monomial libraries only; SURVEY.md 8f/f3 lists general terms as "next")
# ----------------------------------------------------------------------------------------------
def _exponent_table(n: int, degree: int) -> np.ndarray:
    """Exponent vectors of total degree <= ``degree`` with the LAST variable as the most significant
    key and the first variable varying fastest, constant term first: the term order pinned by
    ``test/Core/generators.jl:12-18``."""
    vecs = []

    def rec(idx: int, remaining: int, cur: list):
        if idx < 0:
            vecs.append(list(cur))
            return
        for p in range(remaining + 1):
            cur[idx] = p
            rec(idx - 1, remaining - p, cur)

    rec(n - 1, degree, [0] * n)
    return np.ascontiguousarray(np.asarray(vecs, dtype=np.uint8).T)


class FeatureSource:
    """Features written as CUDA C expressions of the raw variables ``x[i]`` (``[states | controls | t]``) and the
    parameter values ``p[j]`` -- compiled by NVRTC into one feature kernel (``ddb_set_features_source``).  Replaces the
    per-term code generation of ``DataDrivenFunction`` (``src/basis/build_function.jl:6-38``) for arbitrary symbolic
    terms and symbolic parameters."""

    def __init__(self, expressions: Sequence[str], params: Optional[Sequence[float]] = None):
        self.expressions = [str(e) for e in expressions]
        self.params = np.ascontiguousarray([] if params is None else params, dtype=np.float64)

    def __len__(self):
        return len(self.expressions)

    @property
    def body(self) -> str:
        return "\n".join(f"f[{i}] = ({e});" for i, e in enumerate(self.expressions))


@dataclass
class Basis:
    """A monomial candidate library over ``n`` states: ``exponents[i, j]`` = power of state i in
    term j.  Duplicate and zero terms are removed without reordering, as ``__preprocess_basis``
    does (``src/basis/type.jl:91-135``)."""

    exponents: np.ndarray
    names: Optional[List[str]] = None
    implicits: int = 0  # trailing rows of `exponents` that are implicit variables (``Basis(...; implicits = du)``,
    #                     src/basis/type.jl:15-17); they are evaluated on the problem's DX
    # Feature map: when set, row i of `exponents` is the power of FEATURE i = g(coef * x[source]) of the raw
    # states, a sequence of (kind, source, coef) with kind in FEAT_* (``ddb_set_features``); covers the
    # generators sin_basis / cos_basis / fourier_basis / chebyshev_basis and their products with monomials
    features: Optional[List[tuple]] = None
    n_raw: int = 0  # number of raw states when `features` is set
    # Variable order of the rows of `exponents` (the argument order of the generated function,
    # src/basis/build_function.jl:6-38): [states | controls | independent variable | implicits]
    controls: int = 0  # rows that are control inputs (``Basis(...; controls = u)``), evaluated on the problem's U
    use_time: bool = False  # one row for the independent variable (``iv = t``), evaluated on the problem's t

    def __post_init__(self):
        e = np.asarray(self.exponents)
        if e.ndim != 2:
            raise ValueError("exponents must be an (n_states x n_terms) integer table")
        if np.any(e < 0) or np.any(e != np.floor(e)):
            raise ValueError("only monomial terms with non-negative integer powers are supported by the B200 backend")
        e = e.astype(np.uint8)
        keep, seen = [], set()
        for j in range(e.shape[1]):
            key = e[:, j].tobytes()
            if key not in seen:
                seen.add(key)
                keep.append(j)
        self.exponents = np.ascontiguousarray(e[:, keep])

    @property
    def n_states(self) -> int:
        """States the PROBLEM must provide."""
        if self.features is not None:
            return int(self.n_raw)
        return self.exponents.shape[0] - int(self.implicits) - int(self.controls) - int(bool(self.use_time))

    @property
    def n_variables(self) -> int:
        return self.exponents.shape[0]

    @property
    def implicit_idx(self) -> np.ndarray:
        """``implicit_idx[i, j]``: term i depends on implicit variable j (``is_dependent`` table the reference
        builds for an implicit ``Basis``)."""
        lo = self.exponents.shape[0] - int(self.implicits)
        return (self.exponents[lo:, :] > 0).T.copy()

    def __len__(self) -> int:
        return self.exponents.shape[1]

    def term_strings(self) -> List[str]:
        out = []
        for j in range(len(self)):
            parts = []
            for i in range(self.n_variables):
                p = int(self.exponents[i, j])
                if isinstance(self.features, FeatureSource):
                    nm = f"({self.features.expressions[i]})"
                elif self.features is not None:
                    kind, src, coef = self.features[i]
                    fn = {FEAT_IDENTITY: "", FEAT_SIN: "sin", FEAT_COS: "cos", FEAT_EXP: "exp", FEAT_CHEBYSHEV: "cheb"}[kind]
                    arg = f"x{src + 1}" if coef == 1.0 else f"{coef:g}*x{src + 1}"
                    nm = f"{fn}({arg})" if fn else arg
                else:
                    ns, nc, nt = self.n_states, int(self.controls), int(bool(self.use_time))
                    nm = (f"x{i + 1}" if i < ns else f"u{i - ns + 1}" if i < ns + nc else "t" if i < ns + nc + nt
                          else f"dx{i - ns - nc - nt + 1}")
                if p == 1:
                    parts.append(nm)
                elif p > 1:
                    parts.append(f"{nm}^{p}")
            out.append("*".join(parts) if parts else "1")
        return out


FEAT_IDENTITY, FEAT_SIN, FEAT_COS, FEAT_EXP, FEAT_CHEBYSHEV = 0, 1, 2, 3, 4


def _generator_basis(kind: str, n_states: int, coefficients) -> Basis:
    """One library term per (coefficient, state), coefficient-major with the state varying fastest
    (``_generateBasis!``, ``src/utils/basis_generators.jl:15-22``)."""
    cs = list(range(1, int(coefficients) + 1)) if np.isscalar(coefficients) else list(coefficients)
    feats = []
    for t in cs:
        for i in range(int(n_states)):
            if kind == "sin":
                feats.append((FEAT_SIN, i, float(t)))
            elif kind == "cos":
                feats.append((FEAT_COS, i, float(t)))
            elif kind == "chebyshev":
                feats.append((FEAT_CHEBYSHEV, i, float(t)))
            else:  # fourier: iseven(t) ? cos(t x / 2) : sin(t x / 2)
                if int(t) != t:
                    raise ValueError("fourier_basis takes integer coefficients")
                feats.append((FEAT_COS if int(t) % 2 == 0 else FEAT_SIN, i, float(t) / 2.0))
    return Basis(np.eye(len(feats), dtype=np.uint8), features=feats, n_raw=int(n_states))


def sin_basis(n_states: int, coefficients=1) -> Basis:
    """``sin_basis(x, c)`` = ``sin.(t .* x)`` (``src/utils/basis_generators.jl:46-55``)."""
    return _generator_basis("sin", n_states, coefficients)


def cos_basis(n_states: int, coefficients=1) -> Basis:
    """``cos_basis(x, c)`` = ``cos.(t .* x)`` (``basis_generators.jl:61-70``)."""
    return _generator_basis("cos", n_states, coefficients)


def chebyshev_basis(n_states: int, coefficients=1) -> Basis:
    """``chebyshev_basis(x, c)`` = ``cos.(t .* acos.(x))`` (``basis_generators.jl:30-40``)."""
    return _generator_basis("chebyshev", n_states, coefficients)


def fourier_basis(n_states: int, coefficients=1) -> Basis:
    """``fourier_basis(x, c)`` (``basis_generators.jl:76-82``)."""
    return _generator_basis("fourier", n_states, coefficients)


def expression_basis(n_raw: int, expressions: Sequence[str], params: Optional[Sequence[float]] = None,
                     exponents: Optional[np.ndarray] = None) -> Basis:
    """A library of ARBITRARY terms: every expression is CUDA C in the raw variables ``x[0] .. x[n_raw - 1]`` (states,
    then controls, then the independent variable, as ``get_fit_targets`` stacks them) and the parameters ``p[j]``
    -- what ``Basis(eqs, u; parameters = p, iv = t, controls = c)`` is in the reference once Symbolics has generated
    code for it (``src/basis/build_function.jl:6-38``).  ``exponents`` (n_feat x k) makes the terms monomials OVER the
    expressions (default: identity, one term per expression)."""
    fs = FeatureSource(expressions, params)
    e = np.eye(len(fs), dtype=np.uint8) if exponents is None else np.asarray(exponents)
    if e.shape[0] != len(fs):
        raise ValueError("exponents must have one row per expression")
    return Basis(e, features=fs, n_raw=int(n_raw))


def concat_bases(bases: Sequence[Basis]) -> Basis:
    """``Basis([b1; b2; ...], u)``: the union of term lists over the same raw states, duplicates removed
    without reordering (``src/basis/type.jl:91-135``).  Monomial parts are expressed over identity features."""
    n_raw = bases[0].n_states
    feats: List[tuple] = [(FEAT_IDENTITY, i, 1.0) for i in range(n_raw)]
    index = {f: i for i, f in enumerate(feats)}
    cols = []
    for b in bases:
        if b.implicits or b.controls or b.use_time:
            raise ValueError("concat_bases: bases with implicit variables, controls or time are not supported")
        if b.n_states != n_raw:
            raise ValueError("concat_bases: all bases must be over the same states")
        if isinstance(b.features, FeatureSource):
            raise ValueError("concat_bases: expression libraries cannot be merged (write one expression list)")
        bf = b.features if b.features is not None else [(FEAT_IDENTITY, i, 1.0) for i in range(n_raw)]
        loc = []
        for f in bf:
            f = (int(f[0]), int(f[1]), float(f[2]))
            if f not in index:
                index[f] = len(feats)
                feats.append(f)
            loc.append(index[f])
        for j in range(len(b)):
            cols.append({loc[i]: int(b.exponents[i, j]) for i in range(b.exponents.shape[0]) if b.exponents[i, j]})
    E = np.zeros((len(feats), len(cols)), dtype=np.uint8)
    for j, c in enumerate(cols):
        for i, p in c.items():
            E[i, j] = p
    return Basis(E, features=feats, n_raw=n_raw)


def polynomial_basis(n_states: int, degree: int = 1) -> Basis:
    """``polynomial_basis(x, c)`` (``src/utils/basis_generators.jl:111-126``)."""
    if degree <= 0:
        raise ValueError("degree must be positive")
    return Basis(_exponent_table(int(n_states), int(degree)))


def monomial_basis(n_states: int, degree: int = 1) -> Basis:
    """``monomial_basis(x, c)`` = ``[1, x1, x1^2, ..., x1^c, x2, ...]`` (``basis_generators.jl:134-149``)."""
    if degree <= 0:
        raise ValueError("degree must be positive")
    n = int(n_states)
    e = np.zeros((n, n * degree + 1), dtype=np.uint8)
    for i in range(n):
        for j in range(degree):
            e[i, i * degree + j + 1] = j + 1
    return Basis(e)


# ----------------------------------------------------------------------------------------------
# Problems
# ----------------------------------------------------------------------------------------------
@dataclass
class DataDrivenProblem:
    """``DataDrivenProblem{T, ctrl, probType}`` (``src/problem/type.jl:92-113``) without controls /
    parameters (the backend rejects them)."""

    X: np.ndarray
    kind: str  # "continuous" | "discrete" | "direct"
    DX: Optional[np.ndarray] = None
    Y: Optional[np.ndarray] = None
    t: Optional[np.ndarray] = None
    U: Optional[np.ndarray] = None  # control inputs (n_u x m): DMDc, and libraries with `controls`

    def __post_init__(self):
        self.X = np.asarray(self.X, dtype=np.float64)
        for nm in ("DX", "Y"):
            v = getattr(self, nm)
            if v is not None:
                v = np.atleast_2d(np.asarray(v, dtype=np.float64))
                if v.shape[1] != self.X.shape[1]:
                    raise ValueError("One or more measurements are not sized equally.")  # check_lengths, type.jl:421
                setattr(self, nm, v)
        for v in (self.X, self.DX, self.Y):  # check_domain, type.jl:418-420
            if v is not None and not np.all(np.isfinite(v)):
                raise ValueError("One or more measurements contain `NaN` or `Inf`.")

    # get_implicit_data (type.jl:433-435) and get_oop_args (:444-455)
    def fit_inputs(self):
        if self.kind == "continuous":
            return self.X, self.DX
        if self.kind == "direct":
            return self.X, self.Y
        return self.X[:, :-1], self.X[:, 1:]


def ContinuousDataDrivenProblem(X, t=None, DX=None, U=None) -> DataDrivenProblem:
    if DX is None:
        raise ValueError("the B200 backend needs DX (collocation is out of scope: SURVEY.md section 2)")
    return DataDrivenProblem(X=X, kind="continuous", DX=DX, t=t, U=U)


def DiscreteDataDrivenProblem(X, t=None, U=None) -> DataDrivenProblem:
    return DataDrivenProblem(X=X, kind="discrete", t=t, U=U)


def DirectDataDrivenProblem(X, Y, t=None, U=None) -> DataDrivenProblem:
    return DataDrivenProblem(X=X, kind="direct", Y=Y, t=t, U=U)


# ----------------------------------------------------------------------------------------------
# Options and algorithms
# ----------------------------------------------------------------------------------------------
@dataclass
class DataDrivenCommonOptions:
    """``src/utils/common_options.jl:22-54``.  Fields the backend cannot honour must keep their
    defaults."""

    maxiters: int = 100
    abstol: float = math.sqrt(EPS)
    reltol: float = math.sqrt(EPS)
    progress: bool = False
    verbose: bool = False
    denoise: bool = False
    normalize: Optional[str] = None  # None = DataNormalization{Nothing}; "ZScoreTransform" (data_processing.jl:75-80)
    split: float = 1.0  # DataProcessing.split: (rough) fraction of training samples (data_processing.jl:17)
    batchsize: int = 0  # DataProcessing.batchsize: 0 = one batch
    partial: bool = True  # DataProcessing.partial: keep a last, smaller batch
    shuffle_seed: Optional[int] = 0  # seed of the DataLoader's shuffle (its rng field); None = keep the order
    perm: Optional[np.ndarray] = None  # explicit order of the training samples (overrides shuffle_seed)
    roundingmode: str = "RoundToZero"
    digits: int = 10
    selector: str = "bic"


class SoftThreshold:
    name = "SoftThreshold"


class HardThreshold:
    name = "HardThreshold"


class ClippedAbsoluteDeviation:
    name = "ClippedAbsoluteDeviation"

    def __init__(self, rho: float = float("nan")):
        self.rho = float(rho)


def _thr(t) -> np.ndarray:
    a = np.atleast_1d(np.asarray(t, dtype=np.float64)).copy()
    if not np.all(a > 0):
        raise AssertionError("Threshold must be positive definite")
    return a


class STLSQ:
    """``STLSQ(thresholds, rho = 0)`` (``algorithms/STLSQ.jl:48-60``)"""

    name = "STLSQ"

    def __init__(self, thresholds=1e-1, rho: float = 0.0):
        self.thresholds = _thr(thresholds)
        if rho < 0:
            raise AssertionError("Ridge regression parameter must be positive definite!")
        self.rho = float(rho)
        self.proximal = SoftThreshold()


class ADMM:
    """``ADMM(thresholds, rho = 1)`` (``algorithms/ADMM.jl:32-43``)"""

    name = "ADMM"

    def __init__(self, thresholds=1e-1, rho: float = 1.0):
        self.thresholds = _thr(thresholds)
        if not rho > 0:
            raise AssertionError("Augmented Lagrangian parameter should be positive definite")
        self.rho = float(rho)
        self.proximal = SoftThreshold()


class SR3:
    """``SR3(thresholds, nu = 1, R = HardThreshold())`` (``algorithms/SR3.jl:45-73``): with the hard
    threshold the stored thresholds are ``thr^2 / 2``."""

    name = "SR3"

    def __init__(self, thresholds=1e-1, nu: float = 1.0, proximal=None):
        proximal = HardThreshold() if proximal is None else proximal
        t = _thr(thresholds)
        if not nu > 0:
            raise AssertionError("Relaxation must be positive definite")
        self.thresholds = t**2 / 2 if isinstance(proximal, HardThreshold) else t
        self.rho = float(nu)
        self.proximal = proximal


class ImplicitOptimizer:
    """``ImplicitOptimizer(threshold = 1e-1, opt = STLSQ)`` / ``ImplicitOptimizer(opt)``
    (``algorithms/Implicit.jl:37-51``): wraps an explicit optimizer."""

    name = "ImplicitOptimizer"

    def __init__(self, optimizer=None, opt=STLSQ):
        if optimizer is None:
            optimizer = opt(1e-1)
        elif not isinstance(optimizer, (STLSQ, ADMM, SR3)):
            optimizer = opt(optimizer)  # a threshold or threshold grid
        self.optimizer = optimizer

    @property
    def thresholds(self):
        return self.optimizer.thresholds


class B200Sparse:
    """The drop-in algorithm type: wraps one of the reference's sparse regression algorithms and runs
    it on B200 GPUs.  Three ways to use several GPUs, all ending in the same two NCCL all-reduces inside the
    library: ``ngpus = N`` -- ONE process drives N GPUs of the node (``ddb_group``: the samples are cut into N
    ranges inside the library; this is what the Julia glue's ``B200Sparse(inner; ngpus = N)`` does);
    ``group`` -- an initialised ``torch.distributed`` process group with one process per GPU, every rank
    passing ITS row shard of the samples to ``solve`` and receiving the full result.
    """

    def __init__(self, inner, device: int = 0, group=None, ngpus: int = 1, devices=None):
        if not isinstance(inner, (STLSQ, ADMM, SR3, ImplicitOptimizer)):
            raise ValueError("B200Sparse wraps STLSQ, ADMM, SR3 or an ImplicitOptimizer of one of them")
        if ngpus < 1 or (ngpus > 1 and group is not None):
            raise ValueError("B200Sparse: ngpus >= 1, and either ngpus > 1 (one process) or group (one process per GPU)")
        self.inner = inner
        self.device = int(device)
        self.group = group
        self.ngpus = int(ngpus)
        self.devices = devices
        self._ctx: Optional[Context] = None
        self._grp = None

    def context(self) -> Context:
        if self._ctx is None:
            self._ctx = Context(self.device)
        return self._ctx

    def gpu_group(self):
        if self._grp is None:
            from .backend import Group

            self._grp = Group(self.ngpus, self.devices)
        return self._grp


# ----------------------------------------------------------------------------------------------
# Results
# ----------------------------------------------------------------------------------------------
@dataclass
class SparseRegressionResult:
    """``lib/DataDrivenSparse/src/result.jl:1-17``"""

    coefficients: np.ndarray
    dof: int
    lambda_: np.ndarray
    iterations: np.ndarray
    testerror: Optional[float]
    trainerror: float
    retcode: int = 1


@dataclass
class DataDrivenSolution:
    """The numeric content of ``DataDrivenSolution`` (``src/solution.jl:30-68``): the recovered model
    as a rounded coefficient matrix over the input basis, plus the statistics StatsAPI exposes."""

    basis: Basis
    coefficients: np.ndarray  # after __construct_basis rounding
    results: List[SparseRegressionResult]
    rss: float
    dof: int
    nobs: int
    retcode: int
    timings: dict = field(default_factory=dict)
    paths: Optional[SweepOutput] = None
    # sum(abs2, Y .- mean(Y, dims = 2)) from the device pass ddb_target_stats (None when it was not computed)
    null_ss: Optional[float] = None

    # -- StatsAPI statistics of a solution (src/solution.jl:105-157) ------------------------------------------
    def loglikelihood(self) -> float:
        """``-nobs / 2 * log(rss / nobs)`` (``src/solution.jl:117-121``)."""
        with np.errstate(divide="ignore"):
            return float(-self.nobs / 2.0 * np.log(np.float64(self.rss) / self.nobs))

    def nullloglikelihood(self) -> float:
        """``-nobs / 2 * log(sum(abs2, Y .- mean(Y, dims = 2)) / nobs)`` (``src/solution.jl:139-145``)."""
        if self.null_ss is None:
            raise ValueError("the target statistics were not computed for this solution")
        with np.errstate(divide="ignore"):
            return float(-self.nobs / 2.0 * np.log(np.float64(self.null_ss) / self.nobs))

    def r2(self) -> float:
        """Cox-Snell pseudo R^2 (``src/solution.jl:147-157`` -> StatsAPI ``r2(sol, :CoxSnell)``):
        ``1 - exp(2 (nullloglikelihood - loglikelihood) / nobs)``."""
        return float(1.0 - np.exp(2.0 * (self.nullloglikelihood() - self.loglikelihood()) / self.nobs))

    def aic(self) -> float:
        return -2.0 * self.loglikelihood() + 2.0 * self.dof

    def aicc(self) -> float:
        k, n = self.dof, self.nobs
        return self.aic() + 2.0 * k * (k + 1.0) / (n - k - 1.0)

    def bic(self) -> float:
        return -2.0 * self.loglikelihood() + self.dof * float(np.log(self.nobs))

    def get_parameter_values(self) -> np.ndarray:
        """Non-zero coefficients in equation-major order (``__build_eqs`` numbering)."""
        out = []
        for i in range(self.coefficients.shape[0]):
            out.extend(self.coefficients[i][self.coefficients[i] != 0.0].tolist())
        return np.asarray(out)

    def equations(self) -> List[str]:
        names = self.basis.term_strings()
        eqs = []
        for i in range(self.coefficients.shape[0]):
            terms = [f"{self.coefficients[i, j]:.10g}*{names[j]}" for j in np.nonzero(self.coefficients[i])[0]]
            eqs.append(" + ".join(terms) if terms else "0")
        return eqs


def construct_coefficients(C: np.ndarray, digits: int = 10) -> np.ndarray:
    """``round.(X, RoundToZero, digits = digits)`` then zero ``abs <= eps``
    (``src/utils/build_basis.jl:154-157``)."""
    og = 10.0**digits
    out = np.trunc(np.asarray(C, dtype=np.float64) * og) / og
    out[np.abs(out) <= EPS] = 0.0
    return out


# ----------------------------------------------------------------------------------------------
# get_fit_targets / solve
# ----------------------------------------------------------------------------------------------
def get_fit_targets(alg: B200Sparse, prob: DataDrivenProblem, basis: Basis):
    """``DataDrivenDiffEq.get_fit_targets`` for the B200 backend (``src/commonsolve.jl:74-81``):
    returns the RAW states and targets instead of (Theta, Y) so that Theta is never built."""
    X, Y = prob.fit_inputs()
    if X.shape[0] != basis.n_states:
        raise ValueError(f"basis has {basis.n_states} states but the problem has {X.shape[0]}")
    if basis.controls or basis.use_time:
        # the library is a function of (states, controls, t): evaluated on the stacked rows (get_oop_args,
        # src/problem/type.jl:444-455); to the kernels controls and time are just more "states"
        if basis.features is not None:
            raise ValueError("controls / time in a library over a feature map are not supported")
        m = X.shape[1]
        rows = [X]
        if basis.controls:
            if prob.U is None or np.atleast_2d(prob.U).shape[0] != basis.controls:
                raise ValueError(f"basis has {basis.controls} control inputs but the problem provides "
                                 f"{0 if prob.U is None else np.atleast_2d(prob.U).shape[0]}")
            rows.append(np.atleast_2d(np.asarray(prob.U, dtype=np.float64))[:, :m])
        if basis.use_time:
            if prob.t is None:
                raise ValueError("the basis depends on the independent variable but the problem has no time points")
            rows.append(np.asarray(prob.t, dtype=np.float64).reshape(1, -1)[:, :m])
        X = np.vstack(rows)
    if basis.implicits:
        # the library is evaluated on (implicits = get_implicit_data(prob), states); type.jl:433-455
        if Y.shape[0] != basis.implicits:
            raise ValueError(f"basis has {basis.implicits} implicit variables but the problem has {Y.shape[0]} targets")
        return np.vstack([X, Y]), Y
    return X, Y


def _check_options(options: DataDrivenCommonOptions):
    if options.denoise:
        raise ValueError("B200Sparse: denoise=true (optimal shrinkage of Theta) is not supported: Theta is never materialised")
    if options.normalize not in (None, "ZScoreTransform", "UnitRangeTransform"):
        raise ValueError("B200Sparse: normalize must be None (DataNormalization{Nothing}), 'ZScoreTransform' or "
                         "'UnitRangeTransform'")
    if options.roundingmode != "RoundToZero":
        raise ValueError("B200Sparse: only RoundToZero post-processing is supported")
    if options.selector not in ("bic", "aic", "aicc", "rss"):
        raise ValueError("B200Sparse: selector must be one of bic, aic, aicc, rss")


def _params_for(ctx: Context, inner, options: DataDrivenCommonOptions):
    prox = inner.proximal
    return ctx.make_params(
        algorithm=inner.name,
        rho=inner.rho,
        proximal=prox.name,
        selector=options.selector,
        maxiters=options.maxiters,
        abstol=options.abstol,
        reltol=options.reltol,
        cad_rho=getattr(prox, "rho", float("nan")),
    )


def _aicc(rss: float, dof: int, nobs: int) -> float:
    """StatsAPI ``aicc`` of a sparse-regression cache (``DataDrivenSparse.jl:171-201``):
    ll = -n/2 log(rss/n), aic = -2 ll + 2 k, aicc = aic + 2 k (k + 1) / (n - k - 1)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        ll = -nobs / 2.0 * np.log(rss / nobs)
        return float(-2.0 * ll + 2.0 * dof + 2.0 * dof * (dof + 1.0) / (nobs - dof - 1.0))


def implicit_candidate_matrix(implicit_idx: np.ndarray) -> np.ndarray:
    """``candidate_matrix`` (``lib/DataDrivenSparse/src/commonsolve.jl:65-74``): term i is a candidate for
    implicit variable j when it depends on j or on no other implicit variable."""
    implicit_idx = np.asarray(implicit_idx, dtype=bool)
    others = implicit_idx.sum(axis=1, keepdims=True) - implicit_idx
    return implicit_idx | (others == 0)


def implicit_regression(ctx: Context, idx: np.ndarray, necessary: np.ndarray, prm, lambdas, m_total: int, group=None):
    """``(x::ImplicitOptimizer)(X[idx, :], Y; necessary_idx)`` (``Implicit.jl:57-108``) on the Gram blocks
    already in ``ctx``: every selected row is regressed on the others in ONE batched sweep
    (``ddb_implicit_select``), then the best left-hand side is chosen on the host exactly as the reference
    does (dof >= 2, touches a necessary term, smallest AICc, first wins ties).
    Returns ``(x_opt over idx, optimal_threshold, optimal_iterations, rss of the chosen model, SweepOutput)``."""
    idx = np.asarray(idx, dtype=np.int64)
    necessary = np.asarray(necessary, dtype=bool)
    ctx.implicit_select(idx)
    ctx.sweep(prm, lambdas, m_total)
    if group is None:
        ctx.rss()
    else:
        import torch

        from . import dist

        rss = torch.empty(max(1, ctx.rss_count()), dtype=torch.float64, device=f"cuda:{ctx.device}")
        ctx.rss(rss.data_ptr())
        dist.allreduce_sum(rss, group)
    out = ctx.select()
    kk = idx.size
    best = -1
    best_score = 0.0
    for e in range(kk):
        inds = np.ones(kk, dtype=bool)
        inds[e] = False
        c = out.coefficients[e]
        if out.dof[e] >= 2 and np.any(c[inds][necessary[inds]] != 0.0):
            score = _aicc(float(out.rss[e]), int(out.dof[e]), m_total)
            if best < 0 or score < best_score:
                best, best_score = e, score
    if best < 0:
        raise IndexError("ImplicitOptimizer: no candidate with dof >= 2 touches a necessary term (Implicit.jl:100-102)")
    x_opt = out.coefficients[best].copy()
    x_opt[best] = -1.0
    return x_opt, float(out.lambda_opt[best]), int(out.iterations[best]), float(out.rss[best]), out


def _solve_implicit(prob, basis, alg, options, X, Y):
    """``__sparse_regression(ps::InternalDataDrivenProblem{<:ImplicitOptimizer}, X, Y)``
    (``lib/DataDrivenSparse/src/commonsolve.jl:58-114``)."""
    if not basis.implicits:
        raise AssertionError("The provided `Basis` does not have implicit variables!")
    if basis.features is not None:
        raise ValueError("implicit libraries over a feature map are not supported")
    ctx = alg.context()
    ctx.set_features(0, None)
    ctx.set_library(basis.exponents)
    inner = alg.inner.optimizer
    prm = _params_for(ctx, inner, options)
    ctx.upload(X, Y)
    m_total = X.shape[1]
    if alg.group is None:
        ctx.gram()
    else:
        import torch

        from . import dist

        m_total = dist.total_samples(X.shape[1], alg.group, device=f"cuda:{ctx.device}")
        packed = torch.empty(ctx.gram_count(), dtype=torch.float64, device=f"cuda:{ctx.device}")
        ctx.gram(packed.data_ptr())
        dist.allreduce_sum(packed, alg.group)
        k, n_t, base = ctx.k, ctx.n_t, packed.data_ptr()
        ctx.gram_upload_ptr(base, base + 8 * k * k, base + 8 * (k * k + k * n_t), n_t)
    implicit_idx = basis.implicit_idx
    cm = implicit_candidate_matrix(implicit_idx)
    C = np.zeros((cm.shape[1], cm.shape[0]))
    lams, iters, rss_total = [], [], 0.0
    for i in range(cm.shape[1]):
        idx = np.nonzero(cm[:, i])[0]
        x_opt, lam, it, rss_i, _ = implicit_regression(ctx, idx, implicit_idx[idx, i], prm, inner.thresholds, m_total, alg.group)
        C[i, idx] = x_opt
        lams.append(lam)
        iters.append(it)
        rss_total += rss_i  # sum(abs2, C[i, :] * Theta) of the chosen row: its exact streaming rss
    ctx.implicit_select(None)
    res = SparseRegressionResult(coefficients=C, dof=int(np.sum(np.abs(C) > 0.0)), lambda_=np.asarray(lams),
                                 iterations=np.asarray(iters), testerror=None, trainerror=float(rss_total), retcode=1)
    Cr = construct_coefficients(C, options.digits)
    return DataDrivenSolution(basis=basis, coefficients=Cr, results=[res], rss=res.trainerror, dof=int(np.sum(Cr != 0.0)),
                              nobs=int(Y.shape[0] * m_total), retcode=1, timings=ctx.timings())


def _needs_processing(options: DataDrivenCommonOptions) -> bool:
    return options.normalize is not None or options.split != 1.0 or options.batchsize > 0


def zscore_scales(css: np.ndarray, m: int) -> np.ndarray:
    """Scales of ``fit(ZScoreTransform, Theta, dims = 2, center = false)`` (``data_processing.jl:75-80``):
    sigma_j = sqrt(css_j / (m - 1)) with the CENTRED sums of squares ``css_j = sum (theta_j - mean_j)^2`` of
    ``ddb_term_css`` (two passes, like ``std`` of the materialised row in the reference: the one-pass form
    ``sum theta^2 - (sum theta)^2 / m`` cancels for terms whose mean dwarfs their spread); an exactly zero spread --
    a constant row -- is scaled with 1, the reference's ``iszero`` rule."""
    sig = np.sqrt(np.maximum(css, 0.0) / (m - 1))
    sig[sig == 0.0] = 1.0
    return sig


def _solve_processed(prob, basis, alg, options, X, Y, paths=False):
    """``init`` + ``solve!`` with the reference's pre-processing options (``src/commonsolve.jl:92-139``,
    ``lib/DataDrivenSparse/src/commonsolve.jl:1-24``): ZScore normalisation fitted on ALL samples, train/test
    split, shuffled mini-batches; every batch is regressed separately and the result with the smallest test
    (or train) error wins.  Train part, test part and batches are views of the resident data."""
    if alg.group is not None:
        return _solve_processed_sharded(prob, basis, alg, options, X, Y)
    ctx = alg.context()
    ctx.set_features(basis.n_raw, basis.features)
    ctx.set_library(basis.exponents)
    prm = _params_for(ctx, alg.inner, options)
    lambdas = alg.inner.thresholds
    m = X.shape[1]
    k, n_t = len(basis), Y.shape[0]
    split = min(max(float(options.split), 0.0), 1.0)
    m_train = int(round(split * m))  # MLUtils.splitobs(at = split)
    if m_train <= 0:
        raise ValueError("B200Sparse: the training split is empty")
    unit = options.normalize == "UnitRangeTransform"
    if unit:
        # a row of ones rides along as one more target: R[:, n_t] = Theta 1 of every batch view, which the affine
        # transform of the blocks needs; its own regression is dropped from the results
        Y = np.vstack([Y, np.ones((1, m))])
    ctx.upload(X, Y)
    sigma = None
    umin = udiv = None
    if unit:
        # fit(UnitRangeTransform, Theta, dims = 2) on ALL samples (data_processing.jl:68-73; commonsolve.jl:128-130)
        umin, umax = ctx.term_minmax()
        rng = umax - umin
        udiv = np.where(rng > 0.0, rng, 1.0)  # divisor = 1 / scale; an infinite scale (constant row) becomes 1
    if options.normalize == "ZScoreTransform":
        ctx.gram()  # over ALL samples: the transform is fitted before the split (commonsolve.jl:128-130)
        G, _R, _yy = ctx.gram_download()
        const = np.nonzero(basis.exponents.sum(axis=0) == 0)[0]
        if const.size:
            sums = G[const[0]].copy()  # row of the constant term = sum over samples of every term
        else:  # one more pass with a row of ones as the target: R[:, 0] = sum theta
            ctx.upload(X, np.ones((1, m)))
            ctx.gram()
            sums = ctx.gram_download()[1][:, 0].copy()
            ctx.upload(X, Y)
        sigma = zscore_scales(ctx.term_css(sums / m), m)
    # order of the training samples: shuffled by the DataLoader (data_processing.jl:38)
    batchsize = m_train if options.batchsize <= 0 else int(options.batchsize)
    perm = None
    if options.perm is not None:
        perm = np.asarray(options.perm, dtype=np.int64)
    elif options.shuffle_seed is not None and batchsize < m_train:
        perm = np.random.default_rng(options.shuffle_seed).permutation(m_train)
    if perm is not None:
        if perm.size != m_train:
            raise ValueError("perm must order the training samples")
        ctx.permute_samples(np.concatenate([perm, np.arange(m_train, m)]))
    starts = list(range(0, m_train, batchsize))
    if not options.partial and m_train % batchsize:
        starts = starts[:-1]
    results, outs = [], []
    for b0 in starts:
        b1 = min(b0 + batchsize, m_train)
        ctx.set_sample_range(b0, b1)
        ctx.gram()
        if sigma is not None:
            ctx.gram_scale_terms(sigma)
        if unit:
            tsum = ctx.gram_download()[1][:, n_t].copy()       # Theta 1 over this view (the ones target)
            ysum = ctx.target_stats(b1 - b0)[0] * (b1 - b0)     # Y 1
            ctx.gram_affine_terms(umin, udiv, tsum, ysum, b1 - b0)
        ctx.sweep(prm, lambdas, b1 - b0)
        ctx.rss()
        out = ctx.select(paths=paths)
        if unit:  # drop the auxiliary ones target
            out.coefficients, out.lambda_opt, out.iterations = out.coefficients[:n_t], out.lambda_opt[:n_t], out.iterations[:n_t]
            out.rss, out.dof, out.flags = out.rss[:n_t], out.dof[:n_t], out.flags[:n_t]
        Cs = out.coefficients  # in the transformed space when a normalisation is set
        Cu = Cs / sigma[None, :] if sigma is not None else (Cs / udiv[None, :] if unit else Cs)
        testerror = None
        if m_train < m:
            ctx.set_sample_range(m_train, m)
            if unit:  # residual of the transformed-space model on the transformed test rows, evaluated on the raw data
                off = np.concatenate([-(Cu @ umin), [0.0]])
                testerror = float(np.sum(ctx.residual_affine(np.vstack([Cu, np.zeros((1, k))]), off)[:n_t]))
            else:
                testerror = float(np.sum(ctx.residual(Cu)))
        results.append(SparseRegressionResult(coefficients=Cs, dof=int(np.sum(np.abs(Cs) > 0.0)), lambda_=out.lambda_opt,
                                              iterations=out.iterations, testerror=testerror,
                                              trainerror=float(np.sum(out.rss)), retcode=1 if not np.any(out.flags & 1) else 2))
        outs.append(out)
    ctx.set_sample_range(0, -1)
    order = sorted(range(len(results)), key=lambda i: results[i].testerror if results[i].testerror is not None
                   else results[i].trainerror)  # sort!(results, by = l2error), stable
    results = [results[i] for i in order]
    best = results[0]
    if unit:
        # lib commonsolve.jl:19-21 runs the data transform on the transposed coefficients: (C - min) .* scale, as written
        C = (best.coefficients - umin[None, :]) / udiv[None, :]
    else:
        C = best.coefficients / sigma[None, :] if sigma is not None else best.coefficients  # apply_transform (commonsolve.jl:19-21)
    Cr = construct_coefficients(C, options.digits)
    # what DataDrivenSolution recomputes on all data (src/solution.jl:37-56)
    rss_all = float(np.sum(ctx.residual(np.vstack([Cr, np.zeros((1, k))]))[:n_t])) if unit else float(np.sum(ctx.residual(Cr)))
    t = ctx.timings()
    null_ss = float(np.sum(ctx.target_stats(m)[1][:n_t]))
    return DataDrivenSolution(basis=basis, coefficients=Cr, results=results, rss=rss_all, dof=int(np.sum(Cr != 0.0)),
                              nobs=int(n_t * m), retcode=best.retcode, timings=t,
                              paths=outs[order[0]] if paths else None, null_ss=null_ss)


def _solve_processed_sharded(prob, basis, alg, options, X, Y):
    """:func:`_solve_processed` on the row-sharded path (one process per GPU; ``X``, ``Y`` are THIS rank's contiguous
    shard in rank order).  The global sample order, split point and batches are those of the single-process run: the
    split is contiguous (``splitobs(shuffle = false)``), the batches are cut from a global shuffle every rank derives
    from the same seed / ``perm``; a rank's part of a batch or of the test set is a contiguous view of its resident
    samples after one local gather (possibly EMPTY), and every statistic is a sum over the ranks (NCCL all-reduce inside
    the library).  ZScore scales come from the all-reduced blocks of all samples."""
    from . import dist

    ctx = alg.context()
    grp = alg.group
    if not dist._has_comm(ctx):
        dist.init_comm(ctx, grp)
    ctx.set_features(basis.n_raw, basis.features)
    ctx.set_library(basis.exponents)
    prm = _params_for(ctx, alg.inner, options)
    lambdas = alg.inner.thresholds
    m_local = X.shape[1]
    lo, m = dist.shard_offsets(m_local, grp, device=f"cuda:{ctx.device}")
    k, n_t = len(basis), Y.shape[0]
    split = min(max(float(options.split), 0.0), 1.0)
    m_train = int(round(split * m))
    if m_train <= 0:
        raise ValueError("B200Sparse: the training split is empty")
    lt = min(max(m_train - lo, 0), m_local)  # this rank's training samples: local [0, lt); test part: [lt, m_local)
    unit = options.normalize == "UnitRangeTransform"
    if unit:  # a row of ones rides along as one more target (see _solve_processed)
        Y = np.vstack([Y, np.ones((1, m_local))])
    ctx.upload(X, Y)
    sigma = None
    umin = udiv = None
    if unit:
        # extrema over ALL samples: the only exchange primitive of the ABI is a sum, so every rank writes its own
        # (min, max) into its slot of a zero vector and the sum is the gather (adding zeros is exact)
        import torch.distributed as tdist

        rank, world = tdist.get_rank(grp), tdist.get_world_size(grp)
        mn, mx = ctx.term_minmax()
        slots = np.zeros((world, 2, k))
        slots[rank, 0], slots[rank, 1] = mn, mx
        slots = dist.allreduce_host_sum(ctx, slots.ravel(), grp).reshape(world, 2, k)
        umin, umax = slots[:, 0].min(axis=0), slots[:, 1].max(axis=0)
        rng = umax - umin
        udiv = np.where(rng > 0.0, rng, 1.0)
    if options.normalize == "ZScoreTransform":
        ctx.gram()
        ctx.allreduce_gram()
        G, R, _yy = ctx.gram_download()
        const = np.nonzero(basis.exponents.sum(axis=0) == 0)[0]
        if const.size:
            sums = G[const[0]].copy()
        else:
            ctx.upload(X, np.ones((1, m_local)))
            ctx.gram()
            ctx.allreduce_gram()
            sums = ctx.gram_download()[1][:, 0].copy()
            ctx.upload(X, Y)
        sigma = zscore_scales(dist.allreduce_host_sum(ctx, ctx.term_css(sums / m), grp), m)
    batchsize = m_train if options.batchsize <= 0 else int(options.batchsize)
    perm = None
    if options.perm is not None:
        perm = np.asarray(options.perm, dtype=np.int64)
    elif options.shuffle_seed is not None and batchsize < m_train:
        perm = np.random.default_rng(options.shuffle_seed).permutation(m_train)
    pos_local = np.arange(lo, lo + lt)  # position of my training samples in the (shuffled) global training order
    if perm is not None:
        if perm.size != m_train:
            raise ValueError("perm must order the training samples")
        inv = np.empty(m_train, dtype=np.int64)
        inv[perm] = np.arange(m_train)
        pos = inv[lo:lo + lt]
        order = np.argsort(pos, kind="stable")
        pos_local = pos[order]
        ctx.permute_samples(np.concatenate([order, np.arange(lt, m_local)]))
    starts = list(range(0, m_train, batchsize))
    if not options.partial and m_train % batchsize:
        starts = starts[:-1]
    results = []
    for b0 in starts:
        b1 = min(b0 + batchsize, m_train)
        l0, l1 = int(np.searchsorted(pos_local, b0)), int(np.searchsorted(pos_local, b1))
        ctx.set_sample_range(l0, l1)  # possibly empty
        ctx.gram()
        ctx.allreduce_gram()
        if sigma is not None:
            ctx.gram_scale_terms(sigma)
        if unit:
            tsum = ctx.gram_download()[1][:, n_t].copy()       # Theta 1 over this batch, all ranks (the ones target)
            ysum = ctx.target_stats(b1 - b0)[0] * (b1 - b0)     # Y 1 (all-reduced inside)
            ctx.gram_affine_terms(umin, udiv, tsum, ysum, b1 - b0)
        ctx.sweep(prm, lambdas, b1 - b0)
        ctx.rss()
        ctx.allreduce_rss()
        out = ctx.select()
        if unit:  # drop the auxiliary ones target
            out.coefficients, out.lambda_opt, out.iterations = out.coefficients[:n_t], out.lambda_opt[:n_t], out.iterations[:n_t]
            out.rss, out.dof, out.flags = out.rss[:n_t], out.dof[:n_t], out.flags[:n_t]
        Cs = out.coefficients
        Cu = Cs / sigma[None, :] if sigma is not None else (Cs / udiv[None, :] if unit else Cs)
        testerror = None
        if m_train < m:
            ctx.set_sample_range(lt, m_local)
            if unit:
                off = np.concatenate([-(Cu @ umin), [0.0]])
                part = ctx.residual_affine(np.vstack([Cu, np.zeros((1, k))]), off)[:n_t]
            else:
                part = ctx.residual(Cu)
            testerror = float(np.sum(dist.allreduce_host_sum(ctx, part, grp)))
        results.append(SparseRegressionResult(coefficients=Cs, dof=int(np.sum(np.abs(Cs) > 0.0)), lambda_=out.lambda_opt,
                                              iterations=out.iterations, testerror=testerror,
                                              trainerror=float(np.sum(out.rss)), retcode=1 if not np.any(out.flags & 1) else 2))
    ctx.set_sample_range(0, -1)
    order = sorted(range(len(results)), key=lambda i: results[i].testerror if results[i].testerror is not None
                   else results[i].trainerror)
    results = [results[i] for i in order]
    best = results[0]
    if unit:
        C = (best.coefficients - umin[None, :]) / udiv[None, :]  # lib commonsolve.jl:19-21, as written
    else:
        C = best.coefficients / sigma[None, :] if sigma is not None else best.coefficients
    Cr = construct_coefficients(C, options.digits)
    part = ctx.residual(np.vstack([Cr, np.zeros((1, k))]))[:n_t] if unit else ctx.residual(Cr)
    rss_all = float(np.sum(dist.allreduce_host_sum(ctx, part, grp)))
    t = ctx.timings()
    null_ss = float(np.sum(ctx.target_stats(m)[1][:n_t]))
    return DataDrivenSolution(basis=basis, coefficients=Cr, results=results, rss=rss_all, dof=int(np.sum(Cr != 0.0)),
                              nobs=int(n_t * m), retcode=best.retcode, timings=t, null_ss=null_ss)


def solve(prob: DataDrivenProblem, basis: Basis, alg: B200Sparse, options: Optional[DataDrivenCommonOptions] = None,
          paths: bool = False) -> DataDrivenSolution:
    """``solve(prob, basis, alg; options)`` = ``solve!(init(...))`` for the B200 backend.

    Single GPU: one ``ddb_solve_sparse`` call with host buffers.  With ``alg.group`` set, every rank
    passes its own row shard; the packed normal-equation blocks and the candidate-rss table are
    summed with one all-reduce each (``dist.py``)."""
    options = options or DataDrivenCommonOptions()
    _check_options(options)
    X, Y = get_fit_targets(alg, prob, basis)
    if isinstance(alg.inner, ImplicitOptimizer):
        if _needs_processing(options):
            raise ValueError("B200Sparse: normalisation / split / batches are not supported with an ImplicitOptimizer")
        return _solve_implicit(prob, basis, alg, options, X, Y)
    if basis.implicits:
        raise ValueError("a Basis with implicit variables needs an ImplicitOptimizer (src/basis/type.jl:15)")
    if _needs_processing(options):
        return _solve_processed(prob, basis, alg, options, X, Y, paths)
    if alg.ngpus > 1:
        # one process, N GPUs: ddb_group_solve_sparse shards the host buffers itself
        if paths:
            raise ValueError("B200Sparse: paths = True is not available with ngpus > 1")
        grp = alg.gpu_group()
        grp.set_features(basis.n_raw, basis.features)
        grp.set_library(basis.exponents)
        out = grp.solve_sparse(X, Y, _params_for(grp, alg.inner, options), alg.inner.thresholds)
        t = grp.timings()
        null_ss = float(np.sum(grp.target_stats(Y.shape[0])[1]))
        return _finish_solution(basis, options, out, Y.shape[0], X.shape[1], Y.size, t, None, null_ss)
    ctx = alg.context()
    ctx.set_features(basis.n_raw, basis.features)
    ctx.set_library(basis.exponents)
    prm = _params_for(ctx, alg.inner, options)
    lambdas = alg.inner.thresholds
    if alg.group is None:
        if paths:
            ctx.upload(X, Y)
            ctx.gram()
            ctx.sweep(prm, lambdas)
            ctx.rss()
            out = ctx.select(paths=True)
        else:
            out = ctx.solve_sparse(X, Y, prm, lambdas)
        m_total = X.shape[1]
    else:
        from . import dist

        out, m_total = dist.solve_sharded(ctx, X, Y, prm, lambdas, alg.group, paths=paths)
    nobs = int(Y.size if alg.group is None else Y.shape[0] * m_total)
    t = ctx.timings()
    null_ss = float(np.sum(ctx.target_stats(m_total)[1]))  # device pass over the resident targets (collective if sharded)
    return _finish_solution(basis, options, out, Y.shape[0], m_total, nobs, t, out if paths else None, null_ss)


def _finish_solution(basis, options, out, n_t, m_total, nobs, timings, paths_out, null_ss=None):
    C = out.coefficients
    # __sparse_regression (commonsolve.jl:37-55): trainerror = sum(abs2, Y - C*Theta) = sum of the selected rss
    res = SparseRegressionResult(
        coefficients=C,
        dof=int(np.sum(np.abs(C) > 0.0)),
        lambda_=out.lambda_opt,
        iterations=out.iterations,
        testerror=None,
        trainerror=float(np.sum(out.rss)),
        retcode=1 if not np.any(out.flags & 1) else 2,
    )
    Cr = construct_coefficients(C, options.digits)
    return DataDrivenSolution(
        basis=basis,
        coefficients=Cr,
        results=[res],
        rss=res.trainerror,
        dof=int(np.sum(Cr != 0.0)),
        nobs=nobs,
        retcode=res.retcode,
        timings=timings,
        paths=paths_out,
        null_ss=null_ss,
    )
