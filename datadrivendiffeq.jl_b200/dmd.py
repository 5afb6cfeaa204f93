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


def solve_dmd(prob: DataDrivenProblem, alg: B200DMD) -> KoopmanResult:
    import torch

    if prob.kind != "discrete":
        raise ValueError("B200DMD: only discrete problems with the identity lifting are supported")
    x = np.asarray(prob.X, dtype=np.float64)
    n, mp1 = x.shape
    m = mp1 - 1
    if n % 2:
        raise ValueError("B200DMD: the number of states must be even (16-byte TMA rows)")
    ctx = alg.context()
    dev = torch.device(f"cuda:{ctx.device}")
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)  # (m+1, n) row-major == n x (m+1) column-major
    PQ = torch.empty((2, n, n), dtype=torch.float64, device=dev)
    ctx.dmd_gram(xt.data_ptr(), xt.data_ptr() + 8 * n, n, m, PQ[0].data_ptr(), PQ[1].data_ptr())
    t_gram = ctx.timings()["gram_ms"]
    m_total = m
    if alg.group is not None:
        from . import dist

        dist.allreduce_sum(PQ, alg.group)
        m_total = dist.total_samples(m, alg.group, device=str(dev))
    # PQ[i] holds a column-major n x n matrix; as a torch (row-major) tensor that is its transpose
    P = PQ[0].T
    Q = PQ[1].T
    if isinstance(alg.inner, DMDPINV):
        Kd = torch.empty((n, n), dtype=torch.float64, device=dev)
        ctx.dmd_operator(PQ[0].data_ptr(), PQ[1].data_ptr(), n, Kd.data_ptr())
        K = Kd.T.contiguous()
        lam, phi = torch.linalg.eig(K.clone())  # geev overwrites its input; never hand it the operator itself
    else:
        # exact DMD through the Gram identities: P = U S^2 U', V = X' U S^-1, B = Y V S^-1 = Q U S^-2,
        # Atilde = U' B, modes = B omega  (algorithms.jl:147-158)
        s2, U = torch.linalg.eigh(P.contiguous().clone())
        s2 = torch.flip(s2, dims=[0])
        U = torch.flip(U, dims=[1])
        trunc = min(float(alg.inner.truncation), 1.0)
        keep = s2 > (trunc**2) * s2[0]
        keep &= s2 > 0
        U, s2 = U[:, keep], s2[keep]
        B = (Q @ U) / s2[None, :]
        At = U.T @ B
        lam, om = torch.linalg.eig(At.clone())
        phi = B.to(om.dtype) @ om
        K = (phi * lam[None, :]) @ torch.linalg.pinv(phi)
        K = K.real if float(K.imag.abs().max()) <= 1e-9 * max(1.0, float(K.real.abs().max())) else K
    Kh = K.cpu().numpy()
    # rss = ||Z - C K X||^2 with C = I (Z = Y for the identity lifting): chunked, exact residual
    rss = 0.0
    Kr = torch.as_tensor(np.real(Kh), device=dev)
    step = max(1, min(m, (1 << 24) // max(1, n)))
    for s0 in range(0, m, step):
        Xc = xt[s0 : min(m, s0 + step)]
        Yc = xt[s0 + 1 : min(m, s0 + step) + 1]
        rss += float(((Yc - Xc @ Kr.T) ** 2).sum())
    return KoopmanResult(
        K=Kh, eigenvalues=lam.cpu().numpy(), modes=phi.cpu().numpy(), C=np.eye(n), Q=Q.cpu().numpy(), P=P.cpu().numpy(),
        rss=rss, dof=int(np.sum(np.round(np.real(Kh), 10) != 0)), nobs=n * m_total,
        timings={"gram_ms": t_gram, **{k: v for k, v in ctx.timings().items() if k != "gram_ms"}},
    )
