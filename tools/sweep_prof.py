"""One C3-size sweep (k = 2145, 64 targets) on few samples: the launch list of the factorisation / sweep kernels
for `ncu --metrics gpu__time_duration.sum` (the Gram stage is small at 8192 samples)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddb200  # noqa: E402

alg = sys.argv[1] if len(sys.argv) > 1 else "STLSQ"
noise = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
lams = 10.0 ** np.arange(-5, -1 + 1e-9, 0.1)
basis = ddb200.polynomial_basis(64, 2)
with ddb200.Context(0) as ctx:
    ctx.set_library(basis.exponents)
    ctx.generate(1, 64, 8192, 0, 0, noise)
    ctx.gram()
    prm = {"STLSQ": lambda: ctx.make_params("STLSQ"), "ADMM": lambda: ctx.make_params("ADMM", rho=1.0),
           "SR3": lambda: ctx.make_params("SR3", rho=1.0, proximal="HardThreshold")}[alg]()
    for rep in range(3):
        ctx.sweep(prm, lams if alg != "SR3" else ddb200.SR3(lams, 1.0).thresholds)
        ctx.rss()
        out = ctx.select()
        t = ctx.timings()
        print(f"rep {rep}: factor {t['factor_ms']:.3f} ms, sweep {t['sweep_ms']:.3f} ms, rss {t['rss_ms']:.3f} ms, launches {t['sweep_launches']}, "
              f"big_steps {t['big_steps']}, dof {out.dof[:4].tolist()}")
