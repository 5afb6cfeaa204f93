"""Small instances of every hot kernel for `compute-sanitizer` (tools/sanitize.sh): the self-building Gram kernel
(BT = 64 and BT = 128 through the moment staircase), the ring schedule, the persistent STLSQ sweep (shared-memory
and blocked paths), ADMM / SR3 through the inverse path, the exact-rss pass, the DMD Gram kernel and the blocked
triangular solve.  Sizes are tiny: the sanitizer serialises and instruments every access."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddb200  # noqa: E402
from ddb200 import systems  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "all"
lams = 10.0 ** np.arange(-3, -0.9, 0.5)


def sparse(n, deg, m, alg="STLSQ", noise=1e-2, n_t=None, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    X, D = (systems.lorenz_ensemble(m, seed=1, samples_per_traj=256) if n == 3 else
            systems.lorenz96_ensemble(m, n=n, seed=1, samples_per_traj=256))
    D = D + noise * np.random.default_rng(0).standard_normal(D.shape)
    if n_t:
        D = D[:n_t]
    basis = ddb200.polynomial_basis(n, deg)
    with ddb200.Context(0) as ctx:
        ctx.set_library(basis.exponents)
        ctx.upload(X, D)
        ctx.gram()
        sched = ctx.timings()["gram_schedule"]
        prm = {"STLSQ": lambda: ctx.make_params("STLSQ"), "ADMM": lambda: ctx.make_params("ADMM", rho=1.0),
               "SR3": lambda: ctx.make_params("SR3", rho=1.0, proximal="HardThreshold")}[alg]()
        ctx.sweep(prm, lams)
        ctx.rss()
        out = ctx.select()
        print(f"case n={n} deg={deg} m={m} {alg}: schedule {sched}, dof {out.dof.tolist()[:4]}, big_steps {ctx.timings()['big_steps']}")
    for k in (env or {}):
        os.environ.pop(k, None)


if case in ("all", "c2"):
    sparse(3, 5, 2048)                       # gram_kernel<64>, stlsq_small_kernel, rss_kernel
if case in ("all", "moment"):
    sparse(64, 2, 1024, n_t=2)               # gram_kernel<128> over the moment staircase, blocked STLSQ steps
if case in ("all", "ring"):
    sparse(64, 2, 1024, n_t=2, env={"DDB_GRAM_MOMENT": "0"})   # gram_ring_kernel + diagonal pairs
if case in ("all", "relaxed"):
    sparse(24, 2, 1024, alg="ADMM", n_t=3)   # k = 325: inverse path (trsm.cu GEMMs, ninv_apply_kernel)
    sparse(24, 2, 1024, alg="SR3", n_t=3)
if case in ("all", "dmd"):
    A, Xs, Ys = systems.linear_snapshots(130, 600, seed=2, restart=64)
    x = np.concatenate([Xs, Ys[:, -1:]], axis=1)
    with ddb200.Context(0) as ctx:
        P, Q, K = ctx.dmd_solve(x)
    print("case dmd: |K - A| =", float(np.abs(K - A).max()))
