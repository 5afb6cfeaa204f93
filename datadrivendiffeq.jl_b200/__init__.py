"""B200-native backend for the sparse-identification hot path of DataDrivenDiffEq.jl.

``import ddb200`` (root shim) loads this directory as a package.  The numerical work is done by
``libddb200.so`` (``csrc/``, sm_100a only) behind the C ABI of ``include/ddb.h``; this package is
the host-side mirror of the reference's plug-in interface.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .backend import Context, Group, SweepOutput  # noqa: F401
from .api import (  # noqa: F401
    ADMM,
    B200Sparse,
    Basis,
    ClippedAbsoluteDeviation,
    ContinuousDataDrivenProblem,
    DataDrivenCommonOptions,
    DataDrivenProblem,
    DataDrivenSolution,
    DirectDataDrivenProblem,
    DiscreteDataDrivenProblem,
    HardThreshold,
    ImplicitOptimizer,
    implicit_candidate_matrix,
    implicit_regression,
    SR3,
    STLSQ,
    SoftThreshold,
    SparseRegressionResult,
    construct_coefficients,
    get_fit_targets,
    monomial_basis,
    sin_basis,
    cos_basis,
    chebyshev_basis,
    fourier_basis,
    concat_bases,
    expression_basis,
    FeatureSource,
    polynomial_basis,
    solve,
)
from .dmd import B200DMD, DMDPINV, DMDSVD, FBDMD, TOTALDMD, KoopmanResult, solve_dmd  # noqa: F401
from . import dist, systems  # noqa: F401
