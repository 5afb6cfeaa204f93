"""Thin object wrapper over the C ABI (``include/ddb.h``): one ``Context`` per GPU.

Everything numerical happens in ``libddb200.so``; this module only marshals NumPy arrays / device
pointers.  Layout conventions are Julia's: matrices are column-major, states ``n x m`` with one
sample per column.  A NumPy array of shape ``(n, m)`` in Fortran order has exactly that layout.
"""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib

ALG = {"STLSQ": 0, "ADMM": 1, "SR3": 2}
PROX = {"SoftThreshold": 0, "HardThreshold": 1, "ClippedAbsoluteDeviation": 2}
SELECTOR = {"bic": 0, "aic": 1, "aicc": 2, "rss": 3}


def _f(a: np.ndarray) -> np.ndarray:
    """Column-major float64 view/copy of a 2-D array."""
    return np.asfortranarray(a, dtype=np.float64)


@dataclass
class SweepOutput:
    """What ``ddb_select`` returns (un-rounded ``SparseRegressionResult`` content + paths)."""

    coefficients: np.ndarray  # n_t x k
    lambda_opt: np.ndarray  # n_t
    iterations: np.ndarray  # n_t  (total iteration counter, solver.jl:117)
    rss: np.ndarray  # n_t  rss of the selected model
    dof: np.ndarray  # n_t
    flags: np.ndarray  # n_t
    support_path: Optional[np.ndarray] = None  # n_t x L x k bool
    coef_path: Optional[np.ndarray] = None  # n_t x L x k
    iters_path: Optional[np.ndarray] = None  # n_t x L


class Context:
    """One ``ddb_ctx`` (one GPU).  Raises ``DDBError`` on any failure; there is no CPU fallback."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        _lib.check(self._lib.ddb_ctx_create(int(device), C.byref(self._ctx)))
        self.device = int(device)
        self.n = self.k = self.n_t = 0
        self.m = 0
        self._keep = []  # keeps arrays the library retains pointers to alive

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.ddb_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- library ----------------------------------------------------------------------------
    def set_library(self, exponents: np.ndarray):
        """``exponents``: integer array (n x k), entry [i, j] = power of state i in term j."""
        e = np.asfortranarray(exponents, dtype=np.uint8)
        self.n, self.k = e.shape
        self._imp_k = 0
        _lib.check(self._lib.ddb_set_library_monomial(self._ctx, self.n, self.k, _lib.ptr(e)))

    def set_features(self, n_raw: int, features: Optional[Sequence] = None):
        """``ddb_set_features``: the library's states become features ``g(coef * x[src])`` of ``n_raw`` raw
        states; ``features`` = sequence of ``(kind, source, coef)``; ``None`` / empty removes the map.
        Uploads and binds then take RAW states."""
        self._imp_k = 0
        if features is not None and hasattr(features, "body"):  # api.FeatureSource: NVRTC-compiled expressions
            prm = np.ascontiguousarray(features.params, dtype=np.float64)
            _lib.check(self._lib.ddb_set_features_source(self._ctx, int(n_raw), len(features), features.body.encode(),
                                                          _lib.ptr(prm) if prm.size else None, int(prm.size)))
            self.n_raw = int(n_raw)
            return
        if not features:
            _lib.check(self._lib.ddb_set_features(self._ctx, 0, 0, None, None, None))
            self.n_raw = 0
            return
        kind = np.ascontiguousarray([f[0] for f in features], dtype=np.int32)
        src = np.ascontiguousarray([f[1] for f in features], dtype=np.int32)
        coef = np.ascontiguousarray([f[2] for f in features], dtype=np.float64)
        _lib.check(self._lib.ddb_set_features(self._ctx, int(n_raw), kind.size, _lib.ptr(kind), _lib.ptr(src), _lib.ptr(coef)))
        self.n_raw = int(n_raw)

    # -- data -------------------------------------------------------------------------------
    def upload(self, X: np.ndarray, Y: np.ndarray):
        X, Y = _f(X), _f(Y)
        assert X.shape[0] == (getattr(self, "n_raw", 0) or self.n) and X.shape[1] == Y.shape[1]
        self.n_t, self.m = Y.shape
        _lib.check(self._lib.ddb_upload(self._ctx, _lib.ptr(X), _lib.ptr(Y), self.n_t, self.m))

    def upload_raw(self, x_ptr: int, y_ptr: int, n_t: int, m: int):
        """Host pointers (e.g. pinned torch tensors) of column-major X (n x m) and Y (n_t x m)."""
        self.n_t, self.m = int(n_t), int(m)
        _lib.check(self._lib.ddb_upload(self._ctx, C.c_void_p(x_ptr), C.c_void_p(y_ptr), self.n_t, self.m))

    def bind_device(self, dX: int, dY: int, n_t: int, m: int):
        """Device pointers (e.g. ``torch.Tensor.data_ptr()``); the caller keeps the memory alive."""
        self.n_t, self.m = int(n_t), int(m)
        _lib.check(self._lib.ddb_bind_device(self._ctx, C.c_void_p(dX), C.c_void_p(dY), self.n_t, self.m))

    def generate(self, system_id: int, n_states: int, m: int, first_sample: int = 0, seed: int = 0, noise_std: float = 0.0):
        _lib.check(
            self._lib.ddb_generate(self._ctx, system_id, n_states, int(m), int(first_sample), int(seed), float(noise_std))
        )
        self.n_t = 3 if system_id == 0 else n_states
        self.m = int(m)

    def download_data(self):
        X = np.empty((self.n, self.m), order="F")
        Y = np.empty((self.n_t, self.m), order="F")
        _lib.check(self._lib.ddb_download_data(self._ctx, _lib.ptr(X), _lib.ptr(Y)))
        return X, Y

    # -- sample views / permutation / residuals ---------------------------------------------
    def set_sample_range(self, first: int = 0, last: int = -1):
        """``ddb_set_sample_range``: Gram / rss / residual see samples [first, last) only (-1 = to the end)."""
        self._imp_k = 0
        _lib.check(self._lib.ddb_set_sample_range(self._ctx, int(first), int(last)))

    def permute_samples(self, perm: Sequence[int]):
        """``ddb_permute_samples``: new sample s = old sample perm[s]."""
        p = np.ascontiguousarray(perm, dtype=np.int64)
        self._imp_k = 0
        _lib.check(self._lib.ddb_permute_samples(self._ctx, _lib.ptr(p)))

    def gram_scale_terms(self, scale: np.ndarray):
        """``ddb_gram_scale_terms``: the sweep then runs on Theta_j / scale_j (ZScore normalisation)."""
        s = np.ascontiguousarray(scale, dtype=np.float64)
        assert s.size == self.k
        _lib.check(self._lib.ddb_gram_scale_terms(self._ctx, _lib.ptr(s)))

    def residual(self, coef: np.ndarray) -> np.ndarray:
        """``ddb_residual``: rss per target of an arbitrary coefficient matrix (n_t x k) over the current view."""
        c = np.asfortranarray(coef, dtype=np.float64)
        assert c.shape == (self.n_t, self.k)
        out = np.zeros(self.n_t)
        _lib.check(self._lib.ddb_residual(self._ctx, _lib.ptr(c), _lib.ptr(out)))
        return out

    def term_minmax(self):
        """``ddb_term_minmax``: ``(min, max)`` of every library term over the current sample view."""
        mn, mx = np.zeros(self.k), np.zeros(self.k)
        _lib.check(self._lib.ddb_term_minmax(self._ctx, _lib.ptr(mn), _lib.ptr(mx)))
        return mn, mx

    def term_css(self, center) -> np.ndarray:
        """``ddb_term_css``: sum over the current sample view of ``(theta_j - center_j)^2`` for every library term."""
        c = np.ascontiguousarray(center, dtype=np.float64)
        assert c.size == self.k
        out = np.zeros(self.k)
        _lib.check(self._lib.ddb_term_css(self._ctx, _lib.ptr(c), _lib.ptr(out)))
        return out

    def gram_affine_terms(self, shift, scale, tsum, ysum, m_total: int):
        """``ddb_gram_affine_terms``: the sweep then runs on ``(Theta_j - shift_j) / scale_j`` (UnitRangeTransform)."""
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (shift, scale, tsum, ysum)]
        assert a[0].size == a[1].size == a[2].size == self.k and a[3].size == self.n_t
        _lib.check(self._lib.ddb_gram_affine_terms(self._ctx, _lib.ptr(a[0]), _lib.ptr(a[1]), _lib.ptr(a[2]), _lib.ptr(a[3]), int(m_total)))

    def residual_affine(self, coef: np.ndarray, offset: np.ndarray) -> np.ndarray:
        """``ddb_residual_affine``: rss per target of ``Y - (coef Theta + offset)``."""
        c = np.asfortranarray(coef, dtype=np.float64)
        o = np.ascontiguousarray(offset, dtype=np.float64)
        assert c.shape == (self.n_t, self.k) and o.size == self.n_t
        out = np.zeros(self.n_t)
        _lib.check(self._lib.ddb_residual_affine(self._ctx, _lib.ptr(c), _lib.ptr(o), _lib.ptr(out)))
        return out

    def target_stats(self, m_total: Optional[int] = None):
        """``ddb_target_stats``: per-target ``(mean, centred sum of squares)`` of the bound targets (collective on a
        sharded run; ``m_total`` = samples of all ranks)."""
        mean = np.zeros(self.n_t)
        css = np.zeros(self.n_t)
        _lib.check(self._lib.ddb_target_stats(self._ctx, int(m_total or self.m), _lib.ptr(mean), _lib.ptr(css)))
        return mean, css

    # -- normal-equation blocks -------------------------------------------------------------
    def gram_count(self) -> int:
        return int(self._lib.ddb_gram_count(self._ctx))

    def gram(self, d_packed: Optional[int] = None):
        self._imp_k = 0
        _lib.check(self._lib.ddb_gram(self._ctx, C.c_void_p(d_packed) if d_packed else None))

    def upload_gram(self, X: np.ndarray, Y: np.ndarray, d_packed: Optional[int] = None):
        """``ddb_upload_gram``: upload + Gram with the copies overlapped by the kernels (chunked)."""
        X, Y = _f(X), _f(Y)
        assert X.shape[0] == (getattr(self, "n_raw", 0) or self.n) and X.shape[1] == Y.shape[1]
        self.n_t, self.m = Y.shape
        self._imp_k = 0
        _lib.check(self._lib.ddb_upload_gram(self._ctx, _lib.ptr(X), _lib.ptr(Y), self.n_t, self.m,
                                             C.c_void_p(d_packed) if d_packed else None))

    def upload_gram_raw(self, x_ptr: int, y_ptr: int, n_t: int, m: int, d_packed: Optional[int] = None):
        """Same from raw host pointers (e.g. pinned torch tensors)."""
        self.n_t, self.m = int(n_t), int(m)
        self._imp_k = 0
        _lib.check(self._lib.ddb_upload_gram(self._ctx, C.c_void_p(x_ptr), C.c_void_p(y_ptr), self.n_t, self.m,
                                             C.c_void_p(d_packed) if d_packed else None))

    def gram_buffer(self) -> int:
        p = C.c_void_p()
        _lib.check(self._lib.ddb_gram_buffer(self._ctx, C.byref(p)))
        return int(p.value)

    def gram_download(self):
        G = np.empty((self.k, self.k), order="F")
        R = np.empty((self.k, self.n_t), order="F")
        yy = np.empty(self.n_t)
        _lib.check(self._lib.ddb_gram_download(self._ctx, _lib.ptr(G), _lib.ptr(R), _lib.ptr(yy)))
        return G, R, yy

    def gram_upload(self, G: np.ndarray, R: np.ndarray, yy: np.ndarray):
        G, R, yy = _f(G), _f(R), np.ascontiguousarray(yy, dtype=np.float64)
        self.n_t = R.shape[1]
        self._imp_k = 0
        _lib.check(self._lib.ddb_gram_upload(self._ctx, _lib.ptr(G), _lib.ptr(R), _lib.ptr(yy), self.n_t))

    def gram_upload_ptr(self, pG: int, pR: int, pyy: int, n_t: int):
        """Install blocks from raw (host or device) pointers, e.g. an all-reduced torch tensor."""
        self.n_t = int(n_t)
        self._imp_k = 0
        _lib.check(self._lib.ddb_gram_upload(self._ctx, C.c_void_p(pG), C.c_void_p(pR), C.c_void_p(pyy), self.n_t))

    # -- sweep ------------------------------------------------------------------------------
    @staticmethod
    def make_params(algorithm="STLSQ", rho=0.0, proximal="SoftThreshold", selector="bic", maxiters=100,
                    abstol=math.sqrt(np.finfo(float).eps), reltol=math.sqrt(np.finfo(float).eps),
                    cad_rho=float("nan")) -> _lib.SweepParams:
        p = _lib.SweepParams()
        p.algorithm = ALG[algorithm]
        p.proximal = PROX[proximal]
        p.selector = SELECTOR[selector]
        p.maxiters = int(maxiters)
        p.abstol = float(abstol)
        p.reltol = float(reltol)
        p.rho = float(rho)
        p.cad_rho = float(cad_rho)
        return p

    def sweep(self, params: _lib.SweepParams, lambdas: Sequence[float], m_total: Optional[int] = None):
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        self._L = lam.size
        _lib.check(self._lib.ddb_sweep(self._ctx, C.byref(params), _lib.ptr(lam), lam.size, int(m_total or self.m)))

    def rss_count(self) -> int:
        return int(self._lib.ddb_rss_count(self._ctx))

    def rss(self, d_rss: Optional[int] = None):
        _lib.check(self._lib.ddb_rss(self._ctx, C.c_void_p(d_rss) if d_rss else None))

    def rss_buffer(self) -> int:
        p = C.c_void_p()
        _lib.check(self._lib.ddb_rss_buffer(self._ctx, C.byref(p)))
        return int(p.value)

    def implicit_select(self, idx: Optional[Sequence[int]]):
        """``ddb_implicit_select``: the following sweep / rss / select calls regress EACH of the library rows
        ``idx`` on the other selected rows (ImplicitOptimizer).  ``None`` or empty returns to explicit mode.
        A new Gram, data set or library resets the selection on the device side as well."""
        if idx is None or len(idx) == 0:
            _lib.check(self._lib.ddb_implicit_select(self._ctx, None, 0))
            self._imp_k = 0
            return
        a = np.ascontiguousarray(idx, dtype=np.int32)
        _lib.check(self._lib.ddb_implicit_select(self._ctx, _lib.ptr(a), a.size))
        self._imp_k = int(a.size)

    def select(self, paths: bool = False) -> SweepOutput:
        n_t, k, L = self.n_t, self.k, self._L
        if getattr(self, "_imp_k", 0):
            n_t = k = self._imp_k
        W = (k + 31) // 32
        coef = np.zeros((n_t, k), order="F")
        lam = np.zeros(n_t)
        iters = np.zeros(n_t, dtype=np.int32)
        rss = np.zeros(n_t)
        dof = np.zeros(n_t, dtype=np.int32)
        flags = np.zeros(n_t, dtype=np.int32)
        sp = np.zeros((n_t, L, W), dtype=np.uint32) if paths else None
        cp = np.zeros((n_t, L, k)) if paths else None
        ip = np.zeros((n_t, L), dtype=np.int32) if paths else None
        _lib.check(
            self._lib.ddb_select(self._ctx, _lib.ptr(coef), _lib.ptr(lam), _lib.ptr(iters), _lib.ptr(rss), _lib.ptr(dof),
                                 _lib.ptr(sp), _lib.ptr(cp), _lib.ptr(ip), _lib.ptr(flags))
        )
        out = SweepOutput(np.ascontiguousarray(coef), lam, iters, rss, dof, flags)
        if paths:
            bits = np.unpackbits(sp.view(np.uint8).reshape(n_t, L, W * 4), axis=-1, bitorder="little")
            out.support_path = bits[:, :, :k].astype(bool)
            out.coef_path = cp
            out.iters_path = ip
        return out

    def solve_sparse(self, X: np.ndarray, Y: np.ndarray, params: _lib.SweepParams, lambdas: Sequence[float]) -> SweepOutput:
        """The one-call host path ``ddb_solve_sparse`` (host buffers in and out)."""
        X, Y = _f(X), _f(Y)
        lam_in = np.ascontiguousarray(lambdas, dtype=np.float64)
        n_t, m = Y.shape
        self.n_t, self.m, self._L = n_t, m, lam_in.size
        coef = np.zeros((n_t, self.k), order="F")
        lam = np.zeros(n_t)
        iters = np.zeros(n_t, dtype=np.int32)
        rss = np.zeros(n_t)
        dof = np.zeros(n_t, dtype=np.int32)
        flags = np.zeros(n_t, dtype=np.int32)
        _lib.check(
            self._lib.ddb_solve_sparse(self._ctx, _lib.ptr(X), _lib.ptr(Y), n_t, m, C.byref(params), _lib.ptr(lam_in),
                                       lam_in.size, _lib.ptr(coef), _lib.ptr(lam), _lib.ptr(iters), _lib.ptr(rss),
                                       _lib.ptr(dof), None, _lib.ptr(flags))
        )
        return SweepOutput(np.ascontiguousarray(coef), lam, iters, rss, dof, flags)

    # -- multi-GPU: communicator inside the library (comm.cu) -------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """``ddb_comm_unique_id``: 128 bytes created on ONE rank and carried to the others by the host program."""
        buf = (C.c_char * 128)()
        _lib.check(_lib.load().ddb_comm_unique_id(buf))
        return bytes(buf.raw)

    def comm_init_rank(self, nranks: int, rank: int, unique_id: bytes):
        """``ddb_comm_init_rank``: attach an NCCL communicator (one process or thread per GPU)."""
        assert len(unique_id) == 128
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        _lib.check(self._lib.ddb_comm_init_rank(self._ctx, int(nranks), int(rank), buf))

    def comm_info(self):
        """``(nranks, rank, nccl_version)`` of the context's communicator (``(1, 0, v)`` without one)."""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _lib.check(self._lib.ddb_comm_info(self._ctx, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def allreduce_gram(self):
        _lib.check(self._lib.ddb_allreduce_gram(self._ctx))

    def allreduce_rss(self):
        _lib.check(self._lib.ddb_allreduce_rss(self._ctx))

    def allreduce_sum(self, d_ptr: int, count: int):
        _lib.check(self._lib.ddb_allreduce_sum(self._ctx, C.c_void_p(d_ptr), int(count)))

    def solve_sparse_sharded(self, X, Y, params: _lib.SweepParams, lambdas: Sequence[float], m_total: int,
                             x_ptr: Optional[int] = None, y_ptr: Optional[int] = None, n_t: Optional[int] = None,
                             m_local: Optional[int] = None) -> SweepOutput:
        """``ddb_solve_sparse_sharded``: this rank's host shard in, the full result out (collective).  Either
        NumPy arrays ``X``, ``Y`` or raw host pointers (``x_ptr``, ``y_ptr``, ``n_t``, ``m_local``)."""
        if x_ptr is None:
            X, Y = _f(X), _f(Y)
            n_t, m_local = Y.shape
            xp, yp = _lib.ptr(X), _lib.ptr(Y)
        else:
            xp, yp = C.c_void_p(x_ptr), C.c_void_p(y_ptr)
        lam_in = np.ascontiguousarray(lambdas, dtype=np.float64)
        self.n_t, self.m, self._L = int(n_t), int(m_local), lam_in.size
        self._imp_k = 0
        coef = np.zeros((n_t, self.k), order="F")
        lam = np.zeros(n_t)
        iters = np.zeros(n_t, dtype=np.int32)
        rss = np.zeros(n_t)
        dof = np.zeros(n_t, dtype=np.int32)
        flags = np.zeros(n_t, dtype=np.int32)
        _lib.check(
            self._lib.ddb_solve_sparse_sharded(self._ctx, xp, yp, int(n_t), int(m_local), int(m_total), C.byref(params),
                                               _lib.ptr(lam_in), lam_in.size, _lib.ptr(coef), _lib.ptr(lam), _lib.ptr(iters),
                                               _lib.ptr(rss), _lib.ptr(dof), None, _lib.ptr(flags))
        )
        return SweepOutput(np.ascontiguousarray(coef), lam, iters, rss, dof, flags)

    # -- DataDrivenDMD Gram path ------------------------------------------------------------
    def dmd_gram(self, dX: int, dY: int, n: int, m: int, dP: int, dQ: int):
        """P = X X', Q = Y X' for device snapshot matrices (n x m, column-major) into device n x n buffers."""
        _lib.check(self._lib.ddb_dmd_gram(self._ctx, C.c_void_p(dX), C.c_void_p(dY), int(n), int(m), C.c_void_p(dP),
                                          C.c_void_p(dQ)))

    def dmd_operator(self, dP: int, dQ: int, n: int, dK: int):
        """K = Q P^-1 (DMDPINV through the normal equations), all device n x n column-major buffers."""
        _lib.check(self._lib.ddb_dmd_operator(self._ctx, C.c_void_p(dP), C.c_void_p(dQ), int(n), C.c_void_p(dK)))

    def dmd_residual(self, dX: int, dZ: int, dM: int, n_in: int, n_out: int, m: int) -> float:
        """``ddb_dmd_residual``: ``|| Z - M X ||_F^2`` over resident snapshots (device pointers, column-major)."""
        out = C.c_double(0.0)
        _lib.check(self._lib.ddb_dmd_residual(self._ctx, C.c_void_p(dX), C.c_void_p(dZ), C.c_void_p(dM), int(n_in), int(n_out),
                                              int(m), C.byref(out)))
        return float(out.value)

    def dmd_eig_sym(self, dA: int, n: int, dW: int):
        """``ddb_dmd_eig_sym``: eigenvectors overwrite the symmetric matrix at ``dA``, ascending eigenvalues to ``dW``."""
        _lib.check(self._lib.ddb_dmd_eig_sym(self._ctx, C.c_void_p(dA), int(n), C.c_void_p(dW)))

    def dmd_eig(self, dA: int, n: int, dW: int, dVR: Optional[int]):
        """``ddb_dmd_eig``: eigenvalues (complex interleaved) and packed right eigenvectors of a general real matrix."""
        _lib.check(self._lib.ddb_dmd_eig(self._ctx, C.c_void_p(dA), int(n), C.c_void_p(dW), C.c_void_p(dVR) if dVR else None))

    def gemm_nt(self, dD: int, ldd: int, dA: int, lda: int, dB: int, ldb: int, rows: int, cols: int, K: int):
        """``ddb_gemm_nt``: row-major ``D = A B'`` on the library's DMMA kernel."""
        _lib.check(self._lib.ddb_gemm_nt(self._ctx, C.c_void_p(dD), int(ldd), C.c_void_p(dA), int(lda), C.c_void_p(dB), int(ldb),
                                         int(rows), int(cols), int(K)))

    def dmd_solve(self, x: np.ndarray, want_K: bool = True):
        """``ddb_dmd_solve``: host snapshots ``x`` (n x (m+1)) -> ``(P, Q, K)`` with K = Q P^-1 (DMDPINV)."""
        x = _f(x)
        n, mp1 = x.shape
        P = np.empty((n, n), order="F")
        Q = np.empty((n, n), order="F")
        K = np.empty((n, n), order="F") if want_K else None
        _lib.check(self._lib.ddb_dmd_solve(self._ctx, _lib.ptr(x), n, mp1, _lib.ptr(P), _lib.ptr(Q), _lib.ptr(K)))
        return P, Q, K

    # -- introspection ----------------------------------------------------------------------
    def timings(self) -> dict:
        t = _lib.Timings()
        _lib.check(self._lib.ddb_get_timings(self._ctx, C.byref(t)))
        return t.as_dict()

    def stream(self) -> int:
        p = C.c_void_p()
        _lib.check(self._lib.ddb_get_stream(self._ctx, C.byref(p)))
        return int(p.value or 0)

    def synchronize(self):
        _lib.check(self._lib.ddb_synchronize(self._ctx))


class Group:
    """One process driving N GPUs: ``ddb_group`` = N contexts + ``ncclCommInitAll``.  ``solve_sparse`` is the
    N-GPU form of ``Context.solve_sparse`` (host buffers in, host buffers out; the samples are cut into N
    contiguous ranges inside the library)."""

    def __init__(self, ngpus: int, devices: Optional[Sequence[int]] = None):
        self._lib = _lib.load()
        self._g = C.c_void_p()
        dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        _lib.check(self._lib.ddb_group_create(int(ngpus), _lib.ptr(dev), C.byref(self._g)))
        self.ngpus = int(ngpus)
        self.n = self.k = 0

    def close(self):
        if getattr(self, "_g", None) and self._g.value:
            self._lib.ddb_group_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_library(self, exponents: np.ndarray):
        e = np.asfortranarray(exponents, dtype=np.uint8)
        self.n, self.k = e.shape
        _lib.check(self._lib.ddb_group_set_library_monomial(self._g, self.n, self.k, _lib.ptr(e)))

    def set_features(self, n_raw: int, features: Optional[Sequence] = None):
        if features is not None and hasattr(features, "body"):
            prm = np.ascontiguousarray(features.params, dtype=np.float64)
            for r in range(self.ngpus):  # one compiled module per context
                ctx = self._lib.ddb_group_ctx(self._g, r)
                _lib.check(self._lib.ddb_set_features_source(ctx, int(n_raw), len(features), features.body.encode(),
                                                              _lib.ptr(prm) if prm.size else None, int(prm.size)))
            return
        if not features:
            _lib.check(self._lib.ddb_group_set_features(self._g, 0, 0, None, None, None))
            return
        kind = np.ascontiguousarray([f[0] for f in features], dtype=np.int32)
        src = np.ascontiguousarray([f[1] for f in features], dtype=np.int32)
        coef = np.ascontiguousarray([f[2] for f in features], dtype=np.float64)
        _lib.check(self._lib.ddb_group_set_features(self._g, int(n_raw), kind.size, _lib.ptr(kind), _lib.ptr(src), _lib.ptr(coef)))

    def shard(self, m: int, rank: int):
        a, b = C.c_int64(), C.c_int64()
        _lib.check(self._lib.ddb_group_shard(self._g, int(m), int(rank), C.byref(a), C.byref(b)))
        return a.value, b.value

    make_params = staticmethod(Context.make_params)

    def solve_sparse(self, X, Y, params: _lib.SweepParams, lambdas: Sequence[float], x_ptr: Optional[int] = None,
                     y_ptr: Optional[int] = None, n_t: Optional[int] = None, m: Optional[int] = None) -> SweepOutput:
        if x_ptr is None:
            X, Y = _f(X), _f(Y)
            n_t, m = Y.shape
            xp, yp = _lib.ptr(X), _lib.ptr(Y)
        else:
            xp, yp = C.c_void_p(x_ptr), C.c_void_p(y_ptr)
        lam_in = np.ascontiguousarray(lambdas, dtype=np.float64)
        coef = np.zeros((n_t, self.k), order="F")
        lam = np.zeros(n_t)
        iters = np.zeros(n_t, dtype=np.int32)
        rss = np.zeros(n_t)
        dof = np.zeros(n_t, dtype=np.int32)
        flags = np.zeros(n_t, dtype=np.int32)
        _lib.check(
            self._lib.ddb_group_solve_sparse(self._g, xp, yp, int(n_t), int(m), C.byref(params), _lib.ptr(lam_in), lam_in.size,
                                             _lib.ptr(coef), _lib.ptr(lam), _lib.ptr(iters), _lib.ptr(rss), _lib.ptr(dof), None,
                                             _lib.ptr(flags))
        )
        return SweepOutput(np.ascontiguousarray(coef), lam, iters, rss, dof, flags)

    def target_stats(self, n_t: int):
        """``ddb_group_target_stats`` over the shards the last ``solve_sparse`` left resident."""
        mean = np.zeros(n_t)
        css = np.zeros(n_t)
        _lib.check(self._lib.ddb_group_target_stats(self._g, _lib.ptr(mean), _lib.ptr(css)))
        return mean, css

    def timings(self) -> dict:
        t = _lib.Timings()
        _lib.check(self._lib.ddb_group_get_timings(self._g, C.byref(t)))
        return t.as_dict()
