"""Dev check: GPU sweep vs oracle on small noisy problems (real tests live in tests/)."""
import sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np
import ddb200, oracle
from ddb200 import systems

def compare(name, X, Y, ex, alg_name, lambdas, rho=0.0, prox="SoftThreshold", maxiters=100):
    Th = oracle.evaluate_library(ex, X)
    if alg_name == "STLSQ": oalg = oracle.STLSQ(lambdas, rho)
    elif alg_name == "ADMM": oalg = oracle.ADMM(lambdas, rho)
    else:
        pr = {"HardThreshold": oracle.HardThreshold(), "SoftThreshold": oracle.SoftThreshold(), "ClippedAbsoluteDeviation": oracle.ClippedAbsoluteDeviation()}[prox]
        oalg = oracle.SR3(lambdas, rho, pr)
    traces = []
    t0 = time.time()
    ores = oracle.sparse_regression(oalg, Th, Y, oracle.CommonOptions(maxiters=maxiters), traces=traces)
    t_or = time.time() - t0
    ctx = ddb200.Context(0)
    ctx.set_library(ex)
    ctx.upload(X, Y)
    ctx.gram()
    prm = ctx.make_params(alg_name, rho=rho, proximal=prox, maxiters=maxiters)
    ctx.sweep(prm, oalg.thresholds)
    ctx.rss()
    out = ctx.select(paths=True)
    tm = ctx.timings()
    n_t = Y.shape[0]
    sup_ok = True; maxrel = 0.0; lam_ok = True; it_diff = 0
    for i in range(n_t):
        for j in range(len(lambdas)):
            so = traces[i].support_path[j]; sg = out.support_path[i, j]
            if not np.array_equal(so, sg): sup_ok = False
        co = ores.coefficients[i]; cg = out.coefficients[i]
        if not np.array_equal(co != 0, cg != 0): sup_ok = False
        nz = co != 0
        if nz.any(): maxrel = max(maxrel, float(np.max(np.abs(co[nz] - cg[nz]) / np.abs(co[nz]))))
        if abs(ores.lambdas[i] - out.lambda_opt[i]) > 1e-15 * abs(ores.lambdas[i]): lam_ok = False
        it_diff = max(it_diff, abs(ores.iterations[i] - int(out.iterations[i])))
    print(json.dumps({"case": name, "alg": alg_name, "support_identical": sup_ok, "max_rel_coef": maxrel, "lambda_opt_equal": lam_ok,
                      "max_iter_diff": it_diff, "flags": out.flags.tolist()[:4], "oracle_s": round(t_or, 3), "factor_ms": tm["factor_ms"],
                      "sweep_ms": tm["sweep_ms"], "rss_ms": tm["rss_ms"], "cands": tm["candidates"], "big_steps": tm["big_steps"],
                      "dofs": [int(d) for d in out.dof[:6]]}), flush=True)
    ctx.close()

def big_cases():
    rng = np.random.default_rng(5)
    X96, D96 = systems.lorenz96_ensemble(3000, n=16, seed=2, samples_per_traj=500)
    D96n = D96 + 5e-2 * rng.standard_normal(D96.shape)
    ex96 = oracle.polynomial_exponents(16, 2)
    compare("l96_n16_noisy_big", X96, D96n, ex96, "STLSQ", 10.0 ** np.arange(-7, 0.01, 0.25))
    X20, D20 = systems.lorenz96_ensemble(4000, n=20, seed=4, samples_per_traj=500)
    D20n = D20 + 5e-2 * rng.standard_normal(D20.shape)
    ex20 = oracle.polynomial_exponents(20, 2)
    compare("l96_n20_noisy_big", X20, D20n, ex20, "STLSQ", 10.0 ** np.arange(-7, 0.01, 0.25))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        big_cases(); sys.exit(0)
    rng = np.random.default_rng(1)
    lams = 10.0 ** np.arange(-5, -1 + 1e-9, 0.1)
    # C1: README Lorenz, noise-free
    X, DX = systems.lorenz_readme()
    ex = oracle.polynomial_exponents(3, 5)
    compare("lorenz_readme", X, DX, ex, "STLSQ", lams)
    # noisy Lorenz, degree 3 (supports shrink gradually)
    Xn, DXn = systems.lorenz_ensemble(4000, seed=3)
    DXn = DXn + 0.05 * rng.standard_normal(DXn.shape)
    ex3 = oracle.polynomial_exponents(3, 3)
    lam2 = 10.0 ** np.arange(-3, 0.5, 0.1)
    compare("lorenz_noisy_deg3", Xn, DXn, ex3, "STLSQ", lam2)
    compare("lorenz_noisy_deg3_ridge", Xn, DXn, ex3, "STLSQ", lam2, rho=1e-4)
    compare("lorenz_noisy_deg3", Xn, DXn, ex3, "ADMM", lam2, rho=1.0)
    compare("lorenz_noisy_deg3", Xn, DXn, ex3, "SR3", lam2, rho=1.0, prox="HardThreshold")
    compare("lorenz_noisy_deg3", Xn, DXn, ex3, "SR3", lam2, rho=1.0, prox="SoftThreshold")
    # Lorenz-96 n=16 degree 2 (k=153), noisy -> large active sets at small lambda (big path when > 128)
    X96, D96 = systems.lorenz96_ensemble(3000, n=16, seed=2, samples_per_traj=500)
    D96n = D96 + 1e-2 * rng.standard_normal(D96.shape)
    ex96 = oracle.polynomial_exponents(16, 2)
    compare("l96_n16_noisy", X96, D96n, ex96, "STLSQ", 10.0 ** np.arange(-4, 0.01, 0.2))
    compare("l96_n16_clean", X96, D96, ex96, "STLSQ", lams)
