"""B200-native backend for the sparse-identification hot path of DataDrivenDiffEq.jl."""
from . import _lib  # noqa: F401
from .backend import Context, SweepOutput  # noqa: F401
