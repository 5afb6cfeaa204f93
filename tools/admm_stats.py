"""Development tool: how many candidates / distinct supports does an ADMM or SR3 sweep record at the C3 shape?"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, ddb200
lams = 10.0 ** np.arange(-5, -1 + 1e-9, 0.1)
ctx = ddb200.Context(0)
ctx.set_library(ddb200.polynomial_basis(64, 2).exponents)
ctx.generate(1, 64, 200_000, 0, 0, 0.0)
ctx.gram()
for alg, kw in (("STLSQ", {}), ("ADMM", dict(rho=1.0)), ("SR3", dict(rho=1.0, proximal="HardThreshold"))):
    ctx.sweep(ctx.make_params(alg, **kw), lams)
    ctx.rss()
    out = ctx.select(paths=True)
    t = ctx.timings()
    print(alg, "candidates", t["candidates"], "rss_ms", round(t["rss_ms"], 3), "sweep_ms", round(t["sweep_ms"], 2),
          "iters/target", float(out.iterations.mean()), "dof", out.dof[:4], "launches", t["sweep_launches"])
    sp = out.support_path  # n_t x L x k
    print("   distinct supports per target (over thresholds):", [len({r.tobytes() for r in sp[i]}) for i in range(4)])
