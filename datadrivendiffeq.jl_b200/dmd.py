"""Host-side mirror of the DataDrivenDMD entry point for the Gram path.

``solve(prob, DMDPINV())`` / ``solve(prob, DMDSVD())`` on a ``DiscreteDataDrivenProblem`` with the
identity lifting (``src/commonsolve.jl:84-90`` -> ``unit_basis``; ``lib/DataDrivenDMD/src/solve.jl:36-133``).
The sample-dimension contractions ``P = X X'`` and ``Q = Y X'`` (``solve.jl:128-129``) run in the
hand-written DMMA kernel (``csrc/dmd.cu``), row-sharded over ranks with one all-reduce of ``[P | Q]``.
``K = Q P^-1`` (DMDPINV, ``algorithms.jl:80-83``) is solved on the device through the equilibrated blocked
Cholesky.  The dense eigen-decompositions (``eigen(K)``; for DMDSVD ``P = U S^2 U'``, ``Atilde = U' Q U S^-2``,
``algorithms.jl:147-158`` restated through the Gram identities of SURVEY.md 3.4) go through the C ABI as well
(``ddb_dmd_eig_sym`` / ``ddb_dmd_eig``: cuSOLVER bound by the library), the real n x n products through the
library's DMMA kernel (``ddb_gemm_nt``) and the residual of the fitted model through ``ddb_dmd_residual`` -- a
Julia caller gets all of it without torch.  torch is left with device buffers and the complex-valued glue
(``Matrix(::Eigen)``, FBDMD's matrix square root).

Lifted bases (``solve(prob, basis, alg)``, ``lib/DataDrivenDMD/src/solve.jl:3-65``) and continuous problems:
``Theta = basis(X)``, ``Ytilde = basis(X shifted)`` resp. ``J_basis(X) DX``, ``Z`` = the raw targets.  Every block
the driver needs (``P = Theta Theta'``, ``Q = Ytilde Theta'``, ``Ytilde Ytilde'``, ``Z Ytilde'``) is a sub-block of
ONE fused Theta+Gram call of the sparse path (``ddb_gram``) on a doubled monomial library over the stacked states
``[x(s); x(s+1)]`` resp. ``[x; dx]``; the output map ``C = Z / Ytilde`` (``solve.jl:130``) is a second operator solve
and the residual is the sparse path's exact streaming pass (``ddb_residual``).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .api import DataDrivenProblem
from .backend import Context


class DMDPINV:
    name = "DMDPINV"


class DMDSVD:
    name = "DMDSVD"

    def __init__(self, truncation: float = 0.0):
        self.truncation = truncation


class TOTALDMD:
    """``TOTALDMD(truncation, alg)`` (``lib/DataDrivenDMD/src/algorithms.jl:215-248``)."""

    name = "TOTALDMD"

    def __init__(self, truncation=0.0, alg=None):
        self.truncation = truncation
        self.alg = alg if alg is not None else DMDPINV()
        if not isinstance(self.alg, (DMDPINV, DMDSVD)):
            raise ValueError("TOTALDMD wraps DMDPINV or DMDSVD")


class FBDMD:
    """``FBDMD(truncation)`` (``algorithms.jl:276-298``)."""

    name = "FBDMD"

    def __init__(self, truncation=0.0):
        self.alg = DMDSVD(truncation)


class B200DMD:
    """Drop-in Koopman algorithm type (``<: AbstractKoopmanAlgorithm``, ``DataDrivenDMD.jl:71``)."""

    def __init__(self, inner=None, device: int = 0, group=None):
        self.inner = inner or DMDSVD()
        self.device, self.group = int(device), group
        self._ctx: Optional[Context] = None

    def context(self) -> Context:
        if self._ctx is None:
            self._ctx = Context(self.device)
        return self._ctx


@dataclass
class KoopmanResult:
    """Numeric content of ``KoopmanResult`` (``lib/DataDrivenDMD/src/result.jl:20-60``)."""

    K: np.ndarray
    eigenvalues: np.ndarray
    modes: np.ndarray
    C: np.ndarray
    Q: np.ndarray
    P: np.ndarray
    rss: float
    dof: int
    nobs: int
    timings: dict
    B: Optional[np.ndarray] = None  # input map (controlled problems)


# ------------------------------------------------------------------------------------------------
# Everything below works on the Gram blocks of the snapshot matrices.  Every DataDrivenDMD algorithm only
# needs products of the form A B' over the snapshot index: for A = U S V',
#   U, S^2 = eigen(A A'),   V = A' U S^-1,   B V = (B A') U S^-1,
# so the m-long factors of the reference's SVDs never have to exist.
# ------------------------------------------------------------------------------------------------
def _gram_pair(ctx, A, B, group=None):
    """``(A A', B A')`` for device tensors A, B of shape (m, r) (row-major == r x m column-major), r even."""
    import torch

    m, r = A.shape
    assert B.shape == (m, r) and r % 2 == 0 and A.is_contiguous() and B.is_contiguous()
    PQ = torch.empty((2, r, r), dtype=torch.float64, device=A.device)
    ctx.dmd_gram(A.data_ptr(), B.data_ptr(), r, m, PQ[0].data_ptr(), PQ[1].data_ptr())
    if group is not None:
        from . import dist

        if dist._has_comm(ctx):  # NCCL inside the library, on the compute stream
            ctx.allreduce_sum(PQ.data_ptr(), PQ.numel())
            ctx.synchronize()
        else:
            dist.allreduce_sum(PQ, group)
    return PQ[0].T, PQ[1].T, PQ  # column-major buffers seen from row-major torch: transposes


def _truncate(s2, truncation):
    """``truncated_svd`` (``algorithms.jl:2-20``) on squared singular values sorted in descending order."""
    import torch

    if isinstance(truncation, (int, np.integer)) and not isinstance(truncation, bool):
        keep = (torch.arange(s2.numel(), device=s2.device) < int(truncation)) & (s2 > 0)
    else:
        t = min(float(truncation), 1.0)
        keep = (s2 > (t * t) * s2[0]) & (s2 > 0)
    return keep


def _mm(ctx, A, B):
    """``A @ B`` for real device matrices on the library's DMMA kernel (``ddb_gemm_nt``: row-major D = A B2' with
    B2 = B')."""
    import torch

    if A.is_complex() or B.is_complex():
        return A @ B  # complex glue stays with torch
    A = A.contiguous()
    Bt = B.T.contiguous()
    D = torch.empty((A.shape[0], Bt.shape[0]), dtype=torch.float64, device=A.device)
    ctx.gemm_nt(D.data_ptr(), D.shape[1], A.data_ptr(), A.shape[1], Bt.data_ptr(), Bt.shape[1], A.shape[0], Bt.shape[0], A.shape[1])
    return D


def _eigh(ctx, P):
    """``eigen(Symmetric(P))`` through ``ddb_dmd_eig_sym``: ascending eigenvalues, eigenvectors as columns."""
    import torch

    n = P.shape[0]
    A = P.contiguous().clone()  # symmetric: row-major == column-major
    w = torch.empty(n, dtype=torch.float64, device=P.device)
    ctx.dmd_eig_sym(A.data_ptr(), n, w.data_ptr())
    return w, A.T  # the column-major result seen from row-major torch


def _eig(ctx, K):
    """``eigen(K)`` of a general real matrix through ``ddb_dmd_eig``: complex eigenvalues and right eigenvectors."""
    import torch

    if K.is_complex():
        return torch.linalg.eig(K.clone())
    n = K.shape[0]
    A = K.T.contiguous().clone()  # column-major K
    w = torch.empty(2 * n, dtype=torch.float64, device=K.device)
    vr = torch.empty((n, n), dtype=torch.float64, device=K.device)
    ctx.dmd_eig(A.data_ptr(), n, w.data_ptr(), vr.data_ptr())
    lam = torch.complex(w[0::2], w[1::2])
    V = vr.T  # V[:, j] = packed column j
    phi = V.to(torch.complex128).clone()
    im = w[1::2]
    first = torch.nonzero(im > 0).flatten()  # a complex pair (j, j + 1): Re v in column j, Im v in column j + 1
    if first.numel():
        re, imv = V[:, first], V[:, first + 1]
        phi[:, first] = torch.complex(re, imv)
        phi[:, first + 1] = torch.complex(re, -imv)
    return lam, phi


def _left_factors(ctx, P, truncation):
    """Left singular vectors and squared singular values of A from P = A A' (descending, truncated)."""
    import torch

    s2, U = _eigh(ctx, P)
    s2, U = torch.flip(s2, dims=[0]), torch.flip(U, dims=[1])
    keep = _truncate(s2, truncation)
    return U[:, keep], s2[keep]


def _opmat(lam, phi):
    """``Matrix(::Eigen)``; pseudo-inverse for rank-truncated modes; real when the imaginary part is noise."""
    import torch

    K = (phi * lam[None, :]) @ torch.linalg.pinv(phi)
    if K.is_complex() and float(K.imag.abs().max()) <= 1e-9 * max(1.0, float(K.real.abs().max())):
        K = K.real
    return K


def _dmdsvd_blocks(ctx, P, Q, truncation):
    """Exact DMD from P = X X', Q = Y X' (``algorithms.jl:147-158``): B = Y V S^-1 = Q U S^-2."""
    U, s2 = _left_factors(ctx, P, truncation)
    B = _mm(ctx, Q, U) / s2[None, :]
    lam, om = _eig(ctx, _mm(ctx, U.T, B))
    return lam, B.to(om.dtype) @ om


def _dense_alg(ctx, alg, Xs, Ys):
    """``alg(Xs, Ys)`` for SMALL dense matrices (the n x r projections of TOTALDMD)."""
    import torch

    if isinstance(alg, DMDPINV):
        K = Ys @ torch.linalg.pinv(Xs)
        return _eig(ctx, K)
    # truncated SVD of the small projection through the Gram matrix of its SHORT side (min(n, r) singular values,
    # like `svd(Xs)`; the long side's Gram matrix would be rank deficient)
    if Xs.shape[0] <= Xs.shape[1]:
        U, s2 = _left_factors(ctx, _mm(ctx, Xs, Xs.T), alg.truncation)
        S = torch.sqrt(s2)
        V = _mm(ctx, Xs.T, U) / S[None, :]
    else:
        V, s2 = _left_factors(ctx, _mm(ctx, Xs.T, Xs), alg.truncation)
        S = torch.sqrt(s2)
        U = _mm(ctx, Xs, V) / S[None, :]
    B = _mm(ctx, Ys, V) / S[None, :]
    lam, om = _eig(ctx, _mm(ctx, U.T, B))
    return lam, B.to(om.dtype) @ om


def _operator(ctx, P, Qrows):
    """``Qrows P^-1`` (``Y / X`` through the normal equations, ``algorithms.jl:80-83``) on the device: the equilibrated
    blocked Cholesky + DMMA substitution behind ``ddb_dmd_operator``.  ``Qrows``: (rows x r) with rows <= r (padded with
    zero rows, which are independent right-hand sides)."""
    import torch

    r = P.shape[0]
    rows = Qrows.shape[0]
    Qp = torch.zeros((r, r), dtype=torch.float64, device=P.device)
    Qp[:rows] = Qrows
    Qc = Qp.T.contiguous()  # column-major Q
    Kd = torch.empty((r, r), dtype=torch.float64, device=P.device)
    ctx.dmd_operator(P.contiguous().data_ptr(), Qc.data_ptr(), r, Kd.data_ptr())
    return Kd.T[:rows].contiguous()


def _allreduce_scalar(ctx, value, grp, dev):
    import torch

    from . import dist

    t = torch.tensor([value], dtype=torch.float64, device=dev)
    if dist._has_comm(ctx):
        ctx.allreduce_sum(t.data_ptr(), 1)
        ctx.synchronize()
    else:
        dist.allreduce_sum(t, grp)
    return float(t.item())


def _koopman_from_blocks(ctx, inner, P, Q, S_fn):
    """The wrapped algorithm on the Gram blocks P = X X', Q = Y X' (``S_fn()`` returns Y Y' when a variant needs it).
    Returns ``(K, lam, phi)``."""
    import torch

    if isinstance(inner, DMDPINV):
        K = _operator(ctx, P, Q)
        lam, phi = _eig(ctx, K)
        return K, lam, phi
    if isinstance(inner, DMDSVD):
        lam, phi = _dmdsvd_blocks(ctx, P, Q, inner.truncation)
        return _opmat(lam, phi), lam, phi
    S = S_fn()
    if isinstance(inner, FBDMD):  # algorithms.jl:284-292
        A1 = _opmat(*_dmdsvd_blocks(ctx, P, Q, inner.alg.truncation))
        A2 = _opmat(*_dmdsvd_blocks(ctx, S, Q.T, inner.alg.truncation))
        Mx = (A1.to(torch.complex128) @ torch.linalg.inv(A2.to(torch.complex128)))
        mu, V = torch.linalg.eig(Mx)
        At = (V * torch.sqrt(mu)[None, :]) @ torch.linalg.inv(V)  # principal square root
        At = At.abs() * torch.sign(A1.real if A1.is_complex() else A1)
        lam, phi = _eig(ctx, At)
    else:  # TOTALDMD (algorithms.jl:225-228): right singular subspace of [X; Y] from its Gram matrix
        Mz = torch.cat([torch.cat([P, Q.T], dim=1), torch.cat([Q, S], dim=1)], dim=0)
        Uz, s2 = _left_factors(ctx, Mz, inner.truncation)
        sz = torch.sqrt(s2)
        Xs = _mm(ctx, torch.cat([P, Q.T], dim=1), Uz) / sz[None, :]  # X V = X [X; Y]' U S^-1
        Ys = _mm(ctx, torch.cat([Q, S], dim=1), Uz) / sz[None, :]
        lam, phi = _dense_alg(ctx, inner.alg, Xs, Ys)
    return _opmat(lam, phi), lam, phi


def _jacobian_library(E: np.ndarray):
    """For a monomial library with exponent table E (n x k): the distinct monomials over the stacked variables
    ``[x; dx]`` that appear in ``J_theta(x) dx`` and the integer matrix D (k x n_phi) with
    ``(J dx)_j = sum_q D[j, q] phi_q``  (d/dx_i of x^e is e_i x^(e - 1_i), times dx_i)."""
    n, k = E.shape
    cols, index, entries = [], {}, []
    for j in range(k):
        for i in range(n):
            if E[i, j] == 0:
                continue
            e = np.zeros(2 * n, dtype=np.int64)
            e[:n] = E[:, j]
            e[i] -= 1
            e[n + i] = 1
            key = e.tobytes()
            if key not in index:
                index[key] = len(cols)
                cols.append(e)
            entries.append((j, index[key], int(E[i, j])))
    Phi = np.stack(cols, axis=1) if cols else np.zeros((2 * n, 0), dtype=np.int64)
    D = np.zeros((k, Phi.shape[1]))
    for j, q, v in entries:
        D[j, q] += v
    return Phi, D


def _solve_dmd_lifted(prob: DataDrivenProblem, alg: "B200DMD", basis) -> "KoopmanResult":
    """``solve(prob, basis, alg)`` with a monomial lifting (``lib/DataDrivenDMD/src/solve.jl:3-65, 102-133``)."""
    import torch

    from .api import Basis  # noqa: F401

    if basis.features is not None or basis.controls or basis.use_time or basis.implicits:
        raise ValueError("B200DMD: lifted bases must be plain monomial libraries of the states")
    if alg.group is not None:
        raise ValueError("B200DMD: lifted bases are not available on the sharded path")
    E = np.asarray(basis.exponents, dtype=np.int64)
    n, k = E.shape
    x = np.asarray(prob.X, dtype=np.float64)
    if x.shape[0] != n:
        raise ValueError(f"basis has {n} states but the problem has {x.shape[0]}")
    ctx = alg.context()
    dev = torch.device(f"cuda:{ctx.device}")
    Ez = np.zeros_like(E)
    if prob.kind == "discrete":
        # stacked variables [x(s); x(s + 1)]: Theta(s) = basis(x(s)), Ytilde(s) = basis(x(s + 1)), Z = x(s + 1)
        m = x.shape[1] - 1
        stacked = np.vstack([x[:, :m], x[:, 1:]])
        Z = np.ascontiguousarray(x[:, 1:])
        lib = np.hstack([np.vstack([E, Ez]), np.vstack([Ez, E])])
        D = np.eye(k)
    elif prob.kind == "continuous":
        # stacked variables [x; dx]: Ytilde = J_basis(x) dx = D Phi(x, dx), Z = dx  (solve.jl:3-34)
        m = x.shape[1]
        stacked = np.vstack([x, prob.DX])
        Z = np.ascontiguousarray(prob.DX)
        Phi, D = _jacobian_library(E)
        lib = np.hstack([np.vstack([E, Ez]), Phi])
    else:
        raise ValueError("B200DMD: discrete or continuous problems only")
    if int(lib.sum(axis=0).max()) > 8:
        raise ValueError("B200DMD: total degree of the lifted library exceeds 8")
    ctx.set_features(0, None)
    ctx.set_library(lib)
    ctx.upload(stacked, Z)
    ctx.gram()  # ONE fused Theta + Gram pass: every block below is a sub-block of it
    t_gram = ctx.timings()["gram_ms"]
    G, R, _yy = ctx.gram_download()
    Gt = torch.from_numpy(np.ascontiguousarray(G)).to(dev)
    Dt = torch.from_numpy(D).to(dev)
    P = Gt[:k, :k].contiguous()                       # Theta Theta'
    Q = _mm(ctx, Dt, Gt[k:, :k].contiguous())         # Ytilde Theta' = D Phi Theta'
    Smat = [None]

    def S_fn():
        if Smat[0] is None:
            Smat[0] = _mm(ctx, _mm(ctx, Dt, Gt[k:, k:].contiguous()), Dt.T)  # Ytilde Ytilde'
        return Smat[0]

    K, lam, phi = _koopman_from_blocks(ctx, alg.inner, P, Q, S_fn)
    ZYt = torch.from_numpy(np.ascontiguousarray((D @ R[k:, :]).T)).to(dev)  # Z Ytilde'  (n x k)
    Cmap = _operator(ctx, S_fn(), ZYt)                 # C = Z / Ytilde  (solve.jl:130)
    Kh = K.cpu().numpy()
    Kr = np.real(Kh)
    Ch = Cmap.cpu().numpy()
    # rss = || Z - C K Theta ||^2: the sparse path's exact streaming pass with the coefficient matrix C K on the Theta
    # half of the installed library (result.jl:46-59)
    coef = np.zeros((n, lib.shape[1]))
    coef[:, :k] = Ch @ Kr
    rss = float(np.sum(ctx.residual(coef)))
    return KoopmanResult(K=Kh, eigenvalues=lam.cpu().numpy(), modes=phi.cpu().numpy(), C=Ch, Q=Q.cpu().numpy(), P=P.cpu().numpy(),
                         rss=rss, nobs=n * m, dof=int(np.sum(np.round(Kr, 10) != 0)),
                         timings={"gram_ms": t_gram, **{kk: v for kk, v in ctx.timings().items() if kk != "gram_ms"}})


def solve_dmd(prob: DataDrivenProblem, alg: B200DMD, control_input: Optional[np.ndarray] = None, basis=None) -> KoopmanResult:
    """``solve(prob, [basis,] alg; control_input)`` (``lib/DataDrivenDMD/src/solve.jl:3-133``).  Without a basis the
    lifting is the identity (``src/commonsolve.jl:84-90`` -> ``unit_basis``) and the snapshot contractions run in the
    dedicated TMA-fed kernel; with a monomial ``basis`` see :func:`_solve_dmd_lifted`.  ``prob.U`` (optional,
    n_u x (m + 1)) are control inputs; ``control_input`` is a known input matrix B (``solve.jl:121-126`` ->
    ``alg(X, Y, U, B)``)."""
    import torch

    if basis is not None:
        e = np.asarray(basis.exponents)
        identity = (basis.features is None and e.shape[0] == e.shape[1] and np.array_equal(e, np.eye(e.shape[0], dtype=e.dtype)))
        if not identity:
            if getattr(prob, "U", None) is not None or control_input is not None:
                raise ValueError("B200DMD: control inputs with a lifted basis are not supported")
            return _solve_dmd_lifted(prob, alg, basis)
    if prob.kind not in ("discrete", "continuous"):
        raise ValueError("B200DMD: discrete or continuous problems only")
    x = np.asarray(prob.X, dtype=np.float64)
    n = x.shape[0]
    if n % 2:
        raise ValueError("B200DMD: the number of states must be even (16-byte TMA rows)")
    inner = alg.inner
    ctx = alg.context()
    dev = torch.device(f"cuda:{ctx.device}")
    grp = alg.group
    if prob.kind == "discrete":
        m = x.shape[1] - 1
        xt = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)  # (m+1, n) row-major == n x (m+1) column-major
        Xd, Yd = xt[:m], xt[1:]  # views: Y is X shifted by one snapshot, one resident copy serves both
    else:  # continuous, identity lifting: Theta = X, Ytilde = J dx = DX, Z = DX (solve.jl:3-34)
        m = x.shape[1]
        Xd = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
        Yd = torch.from_numpy(np.ascontiguousarray(np.asarray(prob.DX, dtype=np.float64).T)).to(dev)
    Ud = None
    if getattr(prob, "U", None) is not None:
        u = np.atleast_2d(np.asarray(prob.U, dtype=np.float64))
        if u.shape[1] != x.shape[1]:
            raise ValueError("One or more measurements are not sized equally.")
        Ud = torch.from_numpy(np.ascontiguousarray(u.T)).to(dev)[:m]
    Bmap = None
    if Ud is not None and control_input is not None:
        # alg(X, Y, U, B) (algorithms.jl:58-66, 237-248): the known input is removed from the targets
        Bmap = torch.as_tensor(np.atleast_2d(np.asarray(control_input, dtype=np.float64)), device=dev)
        Yfit = (Yd - Ud @ Bmap.T).contiguous()
        Ufit = None
    else:
        Yfit, Ufit = Yd, Ud
    if isinstance(inner, FBDMD) and Ufit is not None:
        inner = inner.alg  # algorithms.jl:294-298: FBDMD falls back to DMDSVD with exogenous signals
    t0 = 0.0

    def blocks(A, B):
        nonlocal t0
        out = _gram_pair(ctx, A if A.is_contiguous() else A.contiguous(), B if B.is_contiguous() else B.contiguous(), grp)
        t0 += ctx.timings()["gram_ms"]
        return out

    m_total = m
    if grp is not None:
        from . import dist

        m_total = dist.total_samples(m, grp, device=str(dev))
    P, Q, PQ = blocks(Xd, Yfit)  # P = X X', Q = Y X'
    Bout = None
    if Ufit is not None:
        # controlled forms: everything over the stacked rows [X; U] (padded to an even count)
        nu = Ufit.shape[1]
        r = n + nu + ((n + nu) % 2)
        A = torch.zeros((m, r), dtype=torch.float64, device=dev)
        A[:, :n], A[:, n:n + nu] = Xd, Ufit
        Ypad = torch.zeros((m, r), dtype=torch.float64, device=dev)
        Ypad[:, :n] = Yfit
        Pt, Qt, PQt = blocks(A, Ypad)
        if r > n + nu:  # the padding row is identically zero: keep the blocks non-singular
            Pt = Pt.clone()
            Pt[r - 1, r - 1] = 1.0
        if isinstance(inner, TOTALDMD):
            raise ValueError("B200DMD: TOTALDMD with unknown input matrix is not supported (pass control_input)")
        if isinstance(inner, DMDPINV):  # Ktilde = Y / [X; U] (algorithms.jl:84-93)
            Kt = _operator(ctx, Pt, Qt[:n].contiguous())
            K = Kt[:, :n].contiguous()
            Bout = Kt[:, n:n + nu].contiguous()
            lam, phi = _eig(ctx, K)
        else:  # DMDc (algorithms.jl:160-176)
            Ut, s2 = _left_factors(ctx, Pt, inner.truncation)
            S, _, _ = blocks(Yfit, Yfit)  # Y Y': left singular vectors of Y
            sh, Uh = _eigh(ctx, S)
            Uh = torch.flip(Uh, dims=[1])
            U1, U2 = Ut[:n], Ut[n:n + nu]
            Cm = _mm(ctx, Qt[:n].contiguous(), Ut) / s2[None, :]
            CU = _mm(ctx, _mm(ctx, Cm, U1.T), Uh)
            lam, om = _eig(ctx, _mm(ctx, Uh.T, CU))
            phi = CU.to(om.dtype) @ om
            Bout = _mm(ctx, Cm, U2.T)
            K = _opmat(lam, phi)
    elif isinstance(inner, DMDPINV):
        Kd = torch.empty((n, n), dtype=torch.float64, device=dev)
        ctx.dmd_operator(PQ[0].data_ptr(), PQ[1].data_ptr(), n, Kd.data_ptr())
        K = Kd.T.contiguous()
        lam, phi = _eig(ctx, K)
    else:
        K, lam, phi = _koopman_from_blocks(ctx, inner, P, Q, lambda: blocks(Yfit, Xd)[0])
    Kh = K.cpu().numpy()
    Bfin = Bmap if Bmap is not None else Bout
    # rss = ||Z - C (K X + B U)||^2 with C = I (Z = Y for the identity lifting): the exact residual, prediction formed
    # tile by tile on the tensor pipe (ddb_dmd_residual)
    Kr = torch.as_tensor(np.ascontiguousarray(np.real(Kh)), device=dev)
    if Bfin is not None and Ud is not None:
        nu = Ud.shape[1]
        r = n + nu + ((n + nu) % 2)
        XU = torch.zeros((m, r), dtype=torch.float64, device=dev)
        XU[:, :n], XU[:, n:n + nu] = Xd, Ud
        M = torch.zeros((n, r), dtype=torch.float64, device=dev)
        M[:, :n], M[:, n:n + nu] = Kr, Bfin
        Zc = Yd.contiguous()
        rss = ctx.dmd_residual(XU.data_ptr(), Zc.data_ptr(), M.T.contiguous().data_ptr(), r, n, m)
    else:
        # Xd / Yd are views of the resident snapshots (row-major m x n == column-major n x m, leading dimension n)
        rss = ctx.dmd_residual(Xd.data_ptr(), Yd.data_ptr(), Kr.T.contiguous().data_ptr(), n, n, m)
    if grp is not None:  # every rank scored its own snapshots: the statistic is the sum over the ranks
        rss = _allreduce_scalar(ctx, rss, grp, dev)
    return KoopmanResult(
        K=Kh, eigenvalues=lam.cpu().numpy(), modes=phi.cpu().numpy(), C=np.eye(n), Q=Q.cpu().numpy(), P=P.cpu().numpy(),
        rss=rss, nobs=n * m_total,
        # dof = non-zero parameters of the recovered model: operator and input map (convert_to_basis, solve.jl:86-100)
        dof=int(np.sum(np.round(np.real(Kh), 10) != 0)) + (0 if Bfin is None else int(np.sum(np.round(Bfin.cpu().numpy(), 10) != 0))),
        timings={"gram_ms": t0, **{k: v for k, v in ctx.timings().items() if k != "gram_ms"}},
        B=None if Bfin is None else Bfin.cpu().numpy(),
    )
