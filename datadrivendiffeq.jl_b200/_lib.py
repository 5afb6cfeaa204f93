"""ctypes binding of ``libddb200.so`` (the C ABI declared in ``include/ddb.h``).

This is the tested stand-in for the Julia ``ccall`` glue (``julia/DataDrivenB200.jl``): same entry
points, same argument meaning.  There is no CPU fallback — if the shared library is missing or no
B200 is present, loading / context creation raises.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libddb200.so")


class DDBError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"ddb status {status}: {message}")
        self.status = status


STATUS = {
    0: "DDB_OK",
    -1: "DDB_INVALID_ARG",
    -2: "DDB_UNSUPPORTED",
    -3: "DDB_CUDA_ERROR",
    -4: "DDB_NCCL_ERROR",
    -5: "DDB_NOT_POSITIVE_DEFINITE",
    -6: "DDB_NONFINITE_INPUT",
    -7: "DDB_OUT_OF_MEMORY",
    -8: "DDB_BAD_STATE",
}


class SweepParams(C.Structure):
    """``ddb_sweep_params``"""

    _fields_ = [
        ("algorithm", C.c_int32),
        ("proximal", C.c_int32),
        ("selector", C.c_int32),
        ("maxiters", C.c_int32),
        ("abstol", C.c_double),
        ("reltol", C.c_double),
        ("rho", C.c_double),
        ("cad_rho", C.c_double),
    ]


class Timings(C.Structure):
    """``ddb_timings``"""

    _fields_ = [
        ("h2d_ms", C.c_double),
        ("gram_ms", C.c_double),
        ("factor_ms", C.c_double),
        ("sweep_ms", C.c_double),
        ("rss_ms", C.c_double),
        ("select_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("gram_launches", C.c_int32),
        ("sweep_launches", C.c_int32),
        ("big_steps", C.c_int32),
        ("candidates", C.c_int32),
        ("gram_dmma_flops", C.c_double),
        ("gram_schedule", C.c_int32),
        ("allreduce_calls", C.c_int32),
        ("allreduce_ms", C.c_double),
        ("allreduce_bytes", C.c_double),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_P = C.c_void_p
_D = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/ddb.h declares
PROTOTYPES = {
    "ddb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "ddb_ctx_destroy": (C.c_int, [_P]),
    "ddb_last_error": (C.c_char_p, []),
    "ddb_version": (C.c_char_p, []),
    "ddb_debug_check_guards": (C.c_int, []),
    "ddb_set_library_monomial": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "ddb_upload": (C.c_int, [_P, _P, _P, C.c_int, C.c_int64]),
    "ddb_bind_device": (C.c_int, [_P, _P, _P, C.c_int, C.c_int64]),
    "ddb_generate": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_uint64, C.c_double]),
    "ddb_download_data": (C.c_int, [_P, _P, _P]),
    "ddb_gram_count": (C.c_int64, [_P]),
    "ddb_gram": (C.c_int, [_P, _P]),
    "ddb_set_sample_range": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "ddb_permute_samples": (C.c_int, [_P, _P]),
    "ddb_gram_scale_terms": (C.c_int, [_P, _P]),
    "ddb_residual": (C.c_int, [_P, _P, _P]),
    "ddb_residual_affine": (C.c_int, [_P, _P, _P, _P]),
    "ddb_term_minmax": (C.c_int, [_P, _P, _P]),
    "ddb_term_css": (C.c_int, [_P, _P, _P]),
    "ddb_gram_affine_terms": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64]),
    "ddb_target_stats": (C.c_int, [_P, C.c_int64, _P, _P]),
    "ddb_group_target_stats": (C.c_int, [_P, _P, _P]),
    "ddb_set_features": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "ddb_set_features_source": (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p, _P, C.c_int]),
    "ddb_features_source_check": (C.c_int, [C.c_char_p, _P, C.c_int]),
    "ddb_upload_gram": (C.c_int, [_P, _P, _P, C.c_int, C.c_int64, _P]),
    "ddb_gram_buffer": (C.c_int, [_P, C.POINTER(_P)]),
    "ddb_gram_download": (C.c_int, [_P, _P, _P, _P]),
    "ddb_moment_wave_describe": (C.c_int64, [C.c_int, C.c_int, _P, C.c_int, C.c_int64, C.c_int, _P, C.c_int64, _P]),
    "ddb_moment_builder_describe": (C.c_int64, [C.c_int, C.c_int, _P, C.c_int, _P, C.c_int64, _P, C.c_int64, _P, C.c_int64]),
    "ddb_moment_plan_describe": (C.c_int64, [C.c_int, C.c_int, _P, C.c_int, _P, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P]),
    "ddb_gram_upload": (C.c_int, [_P, _P, _P, _P, C.c_int]),
    "ddb_implicit_select": (C.c_int, [_P, _P, C.c_int]),
    "ddb_sweep": (C.c_int, [_P, C.POINTER(SweepParams), _P, C.c_int, C.c_int64]),
    "ddb_rss_count": (C.c_int64, [_P]),
    "ddb_rss": (C.c_int, [_P, _P]),
    "ddb_rss_buffer": (C.c_int, [_P, C.POINTER(_P)]),
    "ddb_select": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ddb_solve_sparse": (
        C.c_int,
        [_P, _P, _P, C.c_int, C.c_int64, C.POINTER(SweepParams), _P, C.c_int, _P, _P, _P, _P, _P, _P, _P],
    ),
    "ddb_comm_unique_id": (C.c_int, [_P]),
    "ddb_comm_init_rank": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "ddb_comm_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ddb_allreduce_gram": (C.c_int, [_P]),
    "ddb_allreduce_rss": (C.c_int, [_P]),
    "ddb_allreduce_sum": (C.c_int, [_P, _P, C.c_int64]),
    "ddb_solve_sparse_sharded": (
        C.c_int,
        [_P, _P, _P, C.c_int, C.c_int64, C.c_int64, C.POINTER(SweepParams), _P, C.c_int, _P, _P, _P, _P, _P, _P, _P],
    ),
    "ddb_group_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "ddb_group_destroy": (C.c_int, [_P]),
    "ddb_group_size": (C.c_int, [_P]),
    "ddb_group_ctx": (_P, [_P, C.c_int]),
    "ddb_group_set_library_monomial": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "ddb_group_set_features": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "ddb_group_shard": (C.c_int, [_P, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ddb_group_solve_sparse": (
        C.c_int,
        [_P, _P, _P, C.c_int, C.c_int64, C.POINTER(SweepParams), _P, C.c_int, _P, _P, _P, _P, _P, _P, _P],
    ),
    "ddb_group_get_timings": (C.c_int, [_P, C.POINTER(Timings)]),
    "ddb_dmd_gram": (C.c_int, [_P, _P, _P, C.c_int, C.c_int64, _P, _P]),
    "ddb_dmd_operator": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "ddb_dmd_solve": (C.c_int, [_P, _P, C.c_int, C.c_int64, _P, _P, _P]),
    "ddb_dmd_residual": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, _P]),
    "ddb_dmd_eig_sym": (C.c_int, [_P, _P, C.c_int, _P]),
    "ddb_dmd_eig": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ddb_gemm_nt": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, _P, C.c_int64, C.c_int, C.c_int, C.c_int]),
    "ddb_pow_factors_describe": (C.c_int, [C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    "ddb_gram_plan_describe": (C.c_int64, [C.c_int, C.c_int64, C.c_int, _P, C.c_int64, _P]),
    "ddb_get_timings": (C.c_int, [_P, C.POINTER(Timings)]),
    "ddb_get_stream": (C.c_int, [_P, C.POINTER(_P)]),
    "ddb_synchronize": (C.c_int, [_P]),
}

_lib = None


def load() -> C.CDLL:
    """Load ``libddb200.so`` (built in-tree by ``csrc/Makefile`` / ``__graft_entry__.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().ddb_last_error().decode("utf-8", "replace")
        raise DDBError(status, f"{STATUS.get(status, '?')}: {msg}")


def ptr(a) -> C.c_void_p:
    """Host pointer of a contiguous NumPy array, or None."""
    if a is None:
        return None
    return C.c_void_p(a.ctypes.data)
