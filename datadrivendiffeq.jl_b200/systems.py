"""Synthetic trajectories of the benchmark systems, host side, NumPy.

These produce the *inputs* (states X, exact derivatives DX) exactly the way
``DataDrivenProblem(sol::ODESolution)`` does (``src/problem/type.jl:555-591``: DX is the ODE
right-hand side evaluated at the saved states).  They are small-size companions of the device
generator ``ddb_generate`` (``csrc/generate.cu``), used by tests and by the CPU-baseline sample.
"""

from __future__ import annotations

import numpy as np


def lorenz_rhs(u: np.ndarray, sigma=10.0, rho=28.0, beta=8.0 / 3.0) -> np.ndarray:
    """u: (3, ...) -> du (3, ...)"""
    x, y, z = u[0], u[1], u[2]
    return np.stack([sigma * (y - x), x * (rho - z) - y, x * y - beta * z])


def lorenz96_rhs(u: np.ndarray, F=8.0) -> np.ndarray:
    """u: (n, ...) periodic; du_i = (u_{i+1} - u_{i-2}) u_{i-1} - u_i + F"""
    return (np.roll(u, -1, axis=0) - np.roll(u, 2, axis=0)) * np.roll(u, 1, axis=0) - u + F


def _rk4(rhs, u, h):
    k1 = rhs(u)
    k2 = rhs(u + 0.5 * h * k1)
    k3 = rhs(u + 0.5 * h * k2)
    k4 = rhs(u + h * k3)
    return u + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)


def lorenz_readme(m: int = 1001, dt: float = 0.1, h: float = 1e-3):
    """Config C1: README Lorenz, u0 = (1,0,0), samples every ``dt`` on [0, (m-1) dt]
    (``README.md:40-56``); fixed-step RK4 stands in for Tsit5.  Returns X (3 x m), DX (3 x m)."""
    sub = int(round(dt / h))
    u = np.array([1.0, 0.0, 0.0])
    X = np.empty((3, m))
    X[:, 0] = u
    for s in range(1, m):
        for _ in range(sub):
            u = _rk4(lorenz_rhs, u, h)
        X[:, s] = u
    return X, lorenz_rhs(X)


def ensemble(rhs, u0: np.ndarray, n_traj: int, n_samples: int, h: float, stride: int, spinup: int):
    """Vectorised ensemble: ``u0`` is (n, n_traj).  Returns X (n, n_traj * n_samples) with
    trajectory-major sample order."""
    u = u0.copy()
    for _ in range(spinup):
        u = _rk4(rhs, u, h)
    n = u.shape[0]
    X = np.empty((n, n_traj, n_samples))
    for s in range(n_samples):
        X[:, :, s] = u
        for _ in range(stride):
            u = _rk4(rhs, u, h)
    return X.reshape(n, n_traj * n_samples)


def lorenz_ensemble(m: int, seed: int = 0, samples_per_traj: int = 1000):
    """Config C2 at reduced or full size: independent Lorenz trajectories sampled every 0.01
    (RK4 h = 0.01), initial conditions (1,0,0) + N(0,1), 1000 spin-up steps discarded."""
    n_traj = max(1, m // samples_per_traj)
    spt = m // n_traj
    rng = np.random.default_rng(seed)
    u0 = np.array([1.0, 0.0, 0.0])[:, None] + rng.standard_normal((3, n_traj))
    X = ensemble(lorenz_rhs, u0, n_traj, spt, 0.01, 1, 1000)
    return X, lorenz_rhs(X)


def lorenz96_ensemble(m: int, n: int = 64, seed: int = 0, samples_per_traj: int = 1000, F: float = 8.0):
    """Config C3/C4 at reduced or full size: Lorenz-96, RK4 h = 0.01, one sample every 5 steps,
    2000 spin-up steps from F + 0.01 N(0,1)."""
    n_traj = max(1, m // samples_per_traj)
    spt = m // n_traj
    rng = np.random.default_rng(seed)
    u0 = F + 0.01 * rng.standard_normal((n, n_traj))
    rhs = lambda u: lorenz96_rhs(u, F)  # noqa: E731
    X = ensemble(rhs, u0, n_traj, spt, 0.01, 5, 2000)
    return X, rhs(X)


def linear_snapshots(n: int, m: int, seed: int = 0, restart: int = 256):
    """Config C5: x_{j+1} = A x_j with A = Q blockdiag(2x2 scaled rotations, |lambda| in
    [0.95, 1.0]) Q', restarted from a fresh random state every ``restart`` steps so the snapshot
    matrix stays full rank.  Returns (A, X n x m, Y n x m) with Y the one-step image of X."""
    rng = np.random.default_rng(seed)
    Qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
    D = np.zeros((n, n))
    for b in range(n // 2):
        r = 0.95 + 0.05 * rng.random()
        th = np.pi * rng.random()
        c, s = r * np.cos(th), r * np.sin(th)
        D[2 * b : 2 * b + 2, 2 * b : 2 * b + 2] = [[c, -s], [s, c]]
    if n % 2:
        D[-1, -1] = 0.97
    A = Qm @ D @ Qm.T
    X = np.empty((n, m))
    Y = np.empty((n, m))
    x = rng.standard_normal(n)
    for j in range(m):
        if j % restart == 0:
            x = rng.standard_normal(n)
        X[:, j] = x
        x = A @ x
        Y[:, j] = x
    return A, X, Y


def linear_snapshots_blocked(n: int, m: int, seed: int = 0, restart: int = 64):
    """Same construction as :func:`linear_snapshots` (restarted trajectories of ``x_{j+1} = A x_j``), advanced as
    ONE matrix product per step over all trajectories at once -- for state counts where a matrix-vector product
    per snapshot would take minutes on the host (n = 4096).  Sample ``t * restart + s`` is step ``s`` of trajectory
    ``t``.  Returns (A, X n x m, Y n x m)."""
    rng = np.random.default_rng(seed)
    Qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
    D = np.zeros((n, n))
    for b in range(n // 2):
        r = 0.95 + 0.05 * rng.random()
        th = np.pi * rng.random()
        c, s = r * np.cos(th), r * np.sin(th)
        D[2 * b : 2 * b + 2, 2 * b : 2 * b + 2] = [[c, -s], [s, c]]
    if n % 2:
        D[-1, -1] = 0.97
    A = Qm @ D @ Qm.T
    ntraj = -(-m // restart)
    cur = rng.standard_normal((n, ntraj))
    X = np.empty((n, m), order="F")
    Y = np.empty((n, m), order="F")
    for s in range(restart):
        cols = np.arange(s, m, restart)
        nxt = A @ cur
        X[:, cols] = cur[:, : cols.size]
        Y[:, cols] = nxt[:, : cols.size]
        cur = nxt
    return A, X, Y

