"""Host-side mirror of the DataDrivenDMD entry point for the Gram path.

``solve(prob, DMDPINV())`` / ``solve(prob, DMDSVD())`` on a ``DiscreteDataDrivenProblem`` with the
identity lifting (``src/commonsolve.jl:84-90`` -> ``unit_basis``; ``lib/DataDrivenDMD/src/solve.jl:36-133``).
The sample-dimension contractions ``P = X X'`` and ``Q = Y X'`` (``solve.jl:128-129``) run in the
hand-written DMMA kernel (``csrc/dmd.cu``), row-sharded over ranks with one all-reduce of ``[P | Q]``.
``K = Q P^-1`` (DMDPINV, ``algorithms.jl:80-83``) is solved on the device through the equilibrated blocked
Cholesky.  The small dense eigen-decompositions (``eigen(K)``; for DMDSVD ``P = U S^2 U'``,
``Atilde = U' Q U S^-2``, ``algorithms.jl:147-158`` restated through the Gram identities of SURVEY.md 3.4)
are library calls (``torch.linalg`` = cuSOLVER) -- plumbing around the path, not part of it.
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


def _left_factors(P, truncation):
    """Left singular vectors and squared singular values of A from P = A A' (descending, truncated)."""
    import torch

    s2, U = torch.linalg.eigh(P.contiguous().clone())
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


def _dmdsvd_blocks(P, Q, truncation):
    """Exact DMD from P = X X', Q = Y X' (``algorithms.jl:147-158``): B = Y V S^-1 = Q U S^-2."""
    import torch

    U, s2 = _left_factors(P, truncation)
    B = (Q @ U) / s2[None, :]
    lam, om = torch.linalg.eig((U.T @ B).clone())
    return lam, B.to(om.dtype) @ om


def _dense_alg(alg, Xs, Ys):
    """``alg(Xs, Ys)`` for SMALL dense matrices (the n x r projections of TOTALDMD): library calls."""
    import torch

    if isinstance(alg, DMDPINV):
        K = Ys @ torch.linalg.pinv(Xs)
        lam, phi = torch.linalg.eig(K.clone())
        return lam, phi
    U, S, Vh = torch.linalg.svd(Xs, full_matrices=False)
    keep = _truncate(S * S, alg.truncation)
    U, S, V = U[:, keep], S[keep], Vh[keep].T
    B = (Ys @ V) / S[None, :]
    lam, om = torch.linalg.eig((U.T @ B).clone())
    return lam, B.to(om.dtype) @ om


def solve_dmd(prob: DataDrivenProblem, alg: B200DMD, control_input: Optional[np.ndarray] = None) -> KoopmanResult:
    """``solve(prob, alg; control_input)`` for a ``DiscreteDataDrivenProblem`` with the identity lifting
    (``lib/DataDrivenDMD/src/solve.jl:36-133``).  ``prob.U`` (optional, n_u x (m + 1)) are control inputs;
    ``control_input`` is a known input matrix B (``solve.jl:121-126`` -> ``alg(X, Y, U, B)``)."""
    import torch

    if prob.kind != "discrete":
        raise ValueError("B200DMD: only discrete problems with the identity lifting are supported")
    x = np.asarray(prob.X, dtype=np.float64)
    n, mp1 = x.shape
    m = mp1 - 1
    if n % 2:
        raise ValueError("B200DMD: the number of states must be even (16-byte TMA rows)")
    inner = alg.inner
    ctx = alg.context()
    dev = torch.device(f"cuda:{ctx.device}")
    grp = alg.group
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)  # (m+1, n) row-major == n x (m+1) column-major
    Xd, Yd = xt[:m], xt[1:]  # views: Y is X shifted by one snapshot, one resident copy serves both
    Ud = None
    if getattr(prob, "U", None) is not None:
        u = np.atleast_2d(np.asarray(prob.U, dtype=np.float64))
        if u.shape[1] != mp1:
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
            Kt = torch.linalg.solve(Pt, Qt[:n].T).T
            K = Kt[:, :n].contiguous()
            Bout = Kt[:, n:n + nu].contiguous()
            lam, phi = torch.linalg.eig(K.clone())
        else:  # DMDc (algorithms.jl:160-176)
            Ut, s2 = _left_factors(Pt, inner.truncation)
            S, _, _ = blocks(Yfit, Yfit)  # Y Y': left singular vectors of Y
            sh, Uh = torch.linalg.eigh(S.contiguous().clone())
            Uh = torch.flip(Uh, dims=[1])
            U1, U2 = Ut[:n], Ut[n:n + nu]
            Cm = (Qt[:n] @ Ut) / s2[None, :]
            lam, om = torch.linalg.eig((Uh.T @ Cm @ U1.T @ Uh).clone())
            phi = (Cm @ U1.T @ Uh).to(om.dtype) @ om
            Bout = Cm @ U2.T
            K = _opmat(lam, phi)
    elif isinstance(inner, DMDPINV):
        Kd = torch.empty((n, n), dtype=torch.float64, device=dev)
        ctx.dmd_operator(PQ[0].data_ptr(), PQ[1].data_ptr(), n, Kd.data_ptr())
        K = Kd.T.contiguous()
        lam, phi = torch.linalg.eig(K.clone())  # geev overwrites its input; never hand it the operator itself
    elif isinstance(inner, DMDSVD):
        lam, phi = _dmdsvd_blocks(P, Q, inner.truncation)
        K = _opmat(lam, phi)
    else:
        S, Qt_, _ = blocks(Yfit, Xd)  # S = Y Y', X Y' = Q'
        if isinstance(inner, FBDMD):  # algorithms.jl:284-292
            A1 = _opmat(*_dmdsvd_blocks(P, Q, inner.alg.truncation))
            A2 = _opmat(*_dmdsvd_blocks(S, Q.T, inner.alg.truncation))
            Mx = (A1.to(torch.complex128) @ torch.linalg.inv(A2.to(torch.complex128)))
            mu, V = torch.linalg.eig(Mx)
            At = (V * torch.sqrt(mu)[None, :]) @ torch.linalg.inv(V)  # principal square root
            At = At.abs() * torch.sign(A1.real if A1.is_complex() else A1)
            lam, phi = torch.linalg.eig(At.clone())
        else:  # TOTALDMD (algorithms.jl:225-228): right singular subspace of [X; Y] from its Gram matrix
            Mz = torch.cat([torch.cat([P, Q.T], dim=1), torch.cat([Q, S], dim=1)], dim=0)
            Uz, s2 = _left_factors(Mz, inner.truncation)
            sz = torch.sqrt(s2)
            Xs = (torch.cat([P, Q.T], dim=1) @ Uz) / sz[None, :]  # X V = X [X; Y]' U S^-1
            Ys = (torch.cat([Q, S], dim=1) @ Uz) / sz[None, :]
            lam, phi = _dense_alg(inner.alg, Xs, Ys)
        K = _opmat(lam, phi)
    Kh = K.cpu().numpy()
    Bfin = Bmap if Bmap is not None else Bout
    # rss = ||Z - C (K X + B U)||^2 with C = I (Z = Y for the identity lifting): chunked, exact residual
    rss = 0.0
    Kr = torch.as_tensor(np.real(Kh), device=dev)
    step = max(1, min(m, (1 << 24) // max(1, n)))
    for s0 in range(0, m, step):
        s1 = min(m, s0 + step)
        pred = Xd[s0:s1] @ Kr.T
        if Bfin is not None and Ud is not None:
            pred = pred + Ud[s0:s1] @ Bfin.T
        rss += float(((Yd[s0:s1] - pred) ** 2).sum())
    if grp is not None:  # every rank scored its own snapshots: the statistic is the sum over the ranks
        from . import dist

        rss = float(dist.allreduce_sum(torch.tensor([rss], dtype=torch.float64, device=dev), grp).item())
    return KoopmanResult(
        K=Kh, eigenvalues=lam.cpu().numpy(), modes=phi.cpu().numpy(), C=np.eye(n), Q=Q.cpu().numpy(), P=P.cpu().numpy(),
        rss=rss, nobs=n * m_total,
        # dof = non-zero parameters of the recovered model: operator and input map (convert_to_basis, solve.jl:86-100)
        dof=int(np.sum(np.round(np.real(Kh), 10) != 0)) + (0 if Bfin is None else int(np.sum(np.round(Bfin.cpu().numpy(), 10) != 0))),
        timings={"gram_ms": t0, **{k: v for k, v in ctx.timings().items() if k != "gram_ms"}},
        B=None if Bfin is None else Bfin.cpu().numpy(),
    )
