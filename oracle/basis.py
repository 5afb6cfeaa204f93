"""Oracle: candidate-library generators and evaluation (test infrastructure only).

Follows ``src/utils/basis_generators.jl:84-149`` (term enumeration) and
``src/problem/type.jl:404-407`` -> ``src/basis/build_function.jl:186-205`` (evaluation of every
library term at every sample; the reference calls a Symbolics-generated scalar function once per
column and ``hcat``s the results, which is the same k x m matrix computed here in one shot).
"""

from __future__ import annotations

import numpy as np


def bounded_exponents(n_variables: int, degree: int) -> list[list[int]]:
    """All exponent vectors with total degree <= ``degree`` in the reference's order.

    ``_append_exponents!`` (``src/utils/basis_generators.jl:84-97``) recurses from the LAST
    variable down to the first, so the first variable's exponent varies fastest and the constant
    term comes first.  ``_bounded_exponents`` is ``:99-103``.
    """
    out: list[list[int]] = []
    current = [0] * n_variables

    def rec(index: int, remaining: int) -> None:
        if index == 0:
            out.append(list(current))
            return
        for e in range(remaining + 1):
            current[index - 1] = e
            rec(index - 1, remaining - e)

    rec(n_variables, degree)
    return out


def polynomial_exponents(n_variables: int, degree: int) -> np.ndarray:
    """Exponent table (n x k, uint8) of ``polynomial_basis`` (``basis_generators.jl:111-126``).

    Column j holds the powers of term j; ``exps[i, j]`` is the power of state i.  The order is
    pinned by ``test/Core/generators.jl:12-18``.
    """
    assert degree > 0
    e = np.asarray(bounded_exponents(n_variables, degree), dtype=np.uint8)
    return np.ascontiguousarray(e.T)


def monomial_exponents(n_variables: int, degree: int) -> np.ndarray:
    """Exponent table of ``monomial_basis`` (``basis_generators.jl:134-149``):
    ``[1, x1, x1^2, ..., x1^d, x2, ...]``."""
    assert degree > 0
    k = n_variables * degree + 1
    e = np.zeros((n_variables, k), dtype=np.uint8)
    for i in range(n_variables):
        for j in range(degree):
            e[i, i * degree + j + 1] = j + 1
    return e


def evaluate_library(exps: np.ndarray, X: np.ndarray) -> np.ndarray:
    """Theta (k x m) = every monomial of ``exps`` (n x k) at every column of ``X`` (n x m).

    ``(b::Basis)(prob)`` (``src/problem/type.jl:404-407``) -> ``_apply_vec_function``
    (``src/basis/build_function.jl:186-205``).  Each factor is ``x_i ^ e`` exactly as
    ``polynomial_basis`` writes it (``term *= xi^exponent``, ``basis_generators.jl:117-123``).
    """
    exps = np.asarray(exps)
    X = np.asarray(X, dtype=np.float64)
    n, k = exps.shape
    assert X.shape[0] == n
    m = X.shape[1]
    theta = np.ones((k, m), dtype=np.float64)
    for j in range(k):
        row = theta[j]
        for i in range(n):
            e = int(exps[i, j])
            if e == 1:
                row *= X[i]
            elif e > 1:
                row *= np.power(X[i], e)
    return theta


# ----------------------------------------------------------------------------------------------
# non-polynomial generators (src/utils/basis_generators.jl:30-82) as FEATURES of the states
# ----------------------------------------------------------------------------------------------
FEAT_IDENTITY, FEAT_SIN, FEAT_COS, FEAT_EXP, FEAT_CHEBYSHEV = 0, 1, 2, 3, 4


def _coefficients(c):
    """``f(x, terms::Int) = f(x, 1:terms)`` (``basis_generators.jl:40,55,70,82``)."""
    return list(range(1, int(c) + 1)) if np.isscalar(c) else [v for v in c]


def generator_features(kind: str, n_variables: int, coefficients) -> list:
    """Feature list ``(kind, source, coef)`` of ``sin_basis`` / ``cos_basis`` / ``chebyshev_basis`` /
    ``fourier_basis`` in the reference's order: coefficient-major, variable fastest
    (``_generateBasis!``, ``basis_generators.jl:15-22``)."""
    out = []
    for t in _coefficients(coefficients):
        for i in range(n_variables):
            if kind == "sin":  # sin.(t .* x), :48
                out.append((FEAT_SIN, i, float(t)))
            elif kind == "cos":  # cos.(t .* x), :63
                out.append((FEAT_COS, i, float(t)))
            elif kind == "chebyshev":  # cos.(t .* acos.(x)), :33
                out.append((FEAT_CHEBYSHEV, i, float(t)))
            elif kind == "fourier":  # iseven(t) ? cos.(t .* x ./ 2) : sin.(t .* x ./ 2), :78
                out.append((FEAT_COS if int(t) % 2 == 0 else FEAT_SIN, i, float(t) / 2.0))
            else:
                raise ValueError(kind)
    return out


def evaluate_features(features, X: np.ndarray) -> np.ndarray:
    """Feature matrix (n_feat x m) of raw states X (n_raw x m)."""
    X = np.asarray(X, dtype=np.float64)
    out = np.empty((len(features), X.shape[1]))
    for f, (kind, src, coef) in enumerate(features):
        x = X[src]
        if kind == FEAT_SIN:
            out[f] = np.sin(coef * x)
        elif kind == FEAT_COS:
            out[f] = np.cos(coef * x)
        elif kind == FEAT_EXP:
            out[f] = np.exp(coef * x)
        elif kind == FEAT_CHEBYSHEV:
            out[f] = np.cos(coef * np.arccos(x))
        else:
            out[f] = coef * x
    return out
