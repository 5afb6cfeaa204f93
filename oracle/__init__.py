"""CPU oracle for the sparse-identification hot path of DataDrivenDiffEq.jl.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker (or as the timed CPU baseline), never
as the thing shipped.  The product path (``datadrivendiffeq.jl_b200``) never imports this package
and fails loudly when its CUDA library is missing.

What it is: a NumPy/SciPy restatement, function by function, of the reference's algorithm for
the path ``Basis`` evaluation -> normal-equation blocks -> thresholded lambda sweep
(STLSQ / ADMM / SR3 through ``SparseLinearSolver``) and the DataDrivenDMD operators.  Every
function cites the reference ``file:line`` it follows (paths relative to ``/root/reference``).

Where the arithmetic lives: the reference is pure Julia and delegates all dense linear algebra
to the Julia stdlib ``LinearAlgebra`` (OpenBLAS/LAPACK), which is *not vendored* in
``/root/reference`` (no ``Manifest.toml``; compat bounds ``julia = "1.10"``,
``Symbolics = "7.18.1"``, ``Project.toml:58-61``).  The published algorithms restated here are:
``A \\ B`` / ``B / A`` for non-square A = column-pivoted QR (``geqp3``) with rank truncation at
``rcond = min(size(A)) * eps`` and a minimum-norm solve (LAPACK ``gelsy`` is the same algorithm
and is what SciPy calls); square A = LU (``getrf``/``getrs``); ``cholesky`` = ``potrf``;
``svd`` = ``gesdd``; ``eigen`` = ``geev``.

Parity status: **pinned to the reference's own known-answer tests, to their tolerances** (term
order of ``polynomial_basis``, proximal-operator values, the "Fat"/"Skinny" statistics for
STLSQ/ADMM/SR3, the DMD ground-truth operators; see ``tests/test_oracle_reference_kats.py`` and
SURVEY.md section 8c).  Julia is not installed in the build container or on the GPU box, so the
reference itself cannot be executed; there are no bit-level golden vectors in the reference for
this path (its tests assert statistics within tolerances only).
"""

from .basis import (  # noqa: F401
    bounded_exponents,
    polynomial_exponents,
    monomial_exponents,
    evaluate_library,
    generator_features,
    evaluate_features,
)
from .sparse import (  # noqa: F401
    STLSQ,
    ADMM,
    SR3,
    SoftThreshold,
    HardThreshold,
    ClippedAbsoluteDeviation,
    CommonOptions,
    SparseLinearSolver,
    apply_active_set,
    sparse_regression,
    construct_coefficients,
    ImplicitOptimizer,
    implicit_candidate_matrix,
    implicit_sparse_regression,
    solve_processed,
    zscore_fit,
    bic,
    aic,
    aicc,
    rss,
    dof,
    r2,
)
from .dmd import (  # noqa: F401
    DMDPINV,
    DMDSVD,
    TOTALDMD,
    FBDMD,
    dmdpinv_controlled,
    dmdsvd_controlled,
    operator_matrix,
    truncated_svd,
    koopman_solve,
    koopman_solve_lifted,
    monomial_jacobian_times,
)
