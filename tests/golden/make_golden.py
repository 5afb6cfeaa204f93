"""Generates the committed golden fixtures under tests/golden/.

The reference (Julia) cannot be executed in the build container or on the GPU box, and its own tests
hold no bit-level vectors for this path (they assert statistics within tolerances).  The fixtures are
therefore (a) the deterministic known-answer data of the reference's tests, restated, and (b) outputs
of the oracle (``oracle/``, validated against those known answers) on seeded inputs, so that the GPU
tests can run without recomputing the slow pivoted-QR path every time and so that any drift of either
side shows up as a diff of this directory.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import ddb200  # noqa: E402  (only the NumPy data generators in ddb200.systems are used)
from ddb200 import systems  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LAMS_README = 10.0 ** np.arange(-5, -1 + 1e-9, 0.1)  # exp10.(-5:0.1:-1), README.md:55


def run_case(name, X, Y, exps, alg, **meta):
    Th = oracle.evaluate_library(exps, X)
    traces = []
    res = oracle.sparse_regression(alg, Th, Y, oracle.CommonOptions(), traces=traces)
    sp = np.stack([np.stack(t.support_path) for t in traces])  # n_t x L x k
    cp = np.stack([np.stack(t.coef_path) for t in traces])
    ip = np.stack([np.asarray(t.iters_path) for t in traces])
    margins = []
    for i in range(Y.shape[0]):
        for j, lam in enumerate(alg.thresholds):
            c = np.abs(cp[i, j])
            nz = c[c > 0]
            if nz.size:
                margins.append(float(np.min(np.abs(nz - lam) / lam)))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        X=X, Y=Y, exps=exps, thresholds=alg.thresholds, coefficients=res.coefficients,
        lambda_opt=np.asarray(res.lambdas), iterations=np.asarray(res.iterations), support_path=np.packbits(sp, axis=-1),
        coef_path=cp, iters_path=ip, k=exps.shape[1], min_margin=min(margins) if margins else np.inf,
        G=Th @ Th.T, R=Th @ Y.T, yy=(Y * Y).sum(1), **meta,
    )
    print(name, "dof", [int(s.sum()) for s in sp[:, -1]], "min relative margin |c|-lambda", min(margins) if margins else None)


def main():
    rng = np.random.default_rng(20261019)
    # C1: README Lorenz (README.md:40-56), degree-5 library, 41 thresholds, STLSQ rho = 0
    X, DX = systems.lorenz_readme()
    run_case("c1_readme_lorenz_stlsq", X, DX, oracle.polynomial_exponents(3, 5), oracle.STLSQ(LAMS_README))
    # noisy Lorenz, degree 3: the support shrinks step by step along the grid
    Xn, DXn = systems.lorenz_ensemble(2000, seed=3)
    DXn = DXn + 0.05 * rng.standard_normal(DXn.shape)
    lam2 = 10.0 ** np.arange(-3, 0.5, 0.1)
    ex3 = oracle.polynomial_exponents(3, 3)
    run_case("lorenz_noisy_deg3_stlsq", Xn, DXn, ex3, oracle.STLSQ(lam2))
    run_case("lorenz_noisy_deg3_stlsq_ridge", Xn, DXn, ex3, oracle.STLSQ(lam2, 1e-4))
    # Lorenz-96 n = 8, degree 2 (k = 45), noisy
    X8, D8 = systems.lorenz96_ensemble(1500, n=8, seed=2, samples_per_traj=500)
    D8 = D8 + 1e-2 * rng.standard_normal(D8.shape)
    run_case("l96_n8_noisy_stlsq", X8, D8, oracle.polynomial_exponents(8, 2), oracle.STLSQ(10.0 ** np.arange(-4, 0.01, 0.2)))
    # well-conditioned (standardised states) problems for the Cholesky-based algorithms
    Xs = rng.standard_normal((4, 1500))
    ex4 = oracle.polynomial_exponents(4, 2)
    Ctrue = np.zeros((3, ex4.shape[1]))
    Ctrue[0, [1, 5]] = [1.3, -0.7]
    Ctrue[1, [2, 7, 9]] = [0.9, 0.4, -1.1]
    Ctrue[2, [0, 3]] = [0.5, 2.0]
    Ys = Ctrue @ oracle.evaluate_library(ex4, Xs) + 0.02 * rng.standard_normal((3, 1500))
    lam3 = np.linspace(0.05, 1.0, 12)
    run_case("gauss_deg2_admm", Xs, Ys, ex4, oracle.ADMM(lam3, 1.0))
    run_case("gauss_deg2_sr3_hard", Xs, Ys, ex4, oracle.SR3(lam3, 1.0))
    run_case("gauss_deg2_sr3_soft", Xs, Ys, ex4, oracle.SR3(lam3, 1.0, oracle.SoftThreshold()))
    run_case("gauss_deg2_sr3_cad", Xs, Ys, ex4, oracle.SR3(lam3, 1.0, oracle.ClippedAbsoluteDeviation()))
    # the deterministic developer-API example of the reference
    # (lib/DataDrivenSparse/test/Core/developer_api.jl:4-14): 2 terms x 3 samples, STLSQ(0.1)
    np.savez(os.path.join(HERE, "ref_developer_api.npz"), Theta=np.array([[1.0, 2.0, 3.0], [1.0, 4.0, 9.0]]),
             Y=np.array([[2.0, 4.0, 6.0]]), threshold=0.1)


def c4_inputs():
    """Inputs of the C4-size fixtures (regenerated from the seed by the tests, not stored): the full Lorenz-96
    library of BASELINE config 4 (n = 64, degree 2, k = 2145 terms) on 6000 noisy samples, 4 of the 64 targets."""
    X, D = systems.lorenz96_ensemble(6000, n=64, seed=7, samples_per_traj=500)
    D = D + 1e-2 * np.random.default_rng(17).standard_normal(D.shape)
    return X, D[[0, 21, 42, 63]], oracle.polynomial_exponents(64, 2), 10.0 ** np.arange(-3, -0.9, 0.5)


def c4_size():
    """ADMM(rho = 1) and SR3(nu = 1, HardThreshold) step traces at k = 2145: the path that iterates through the explicit
    inverse of the equilibrated matrix (`ninv_apply_kernel`) at the size the bench runs it.  The oracle needs minutes
    here (one 2145 x 2145 `cho_solve` per iteration and target), so its traces are committed."""
    X, Y, ex, lams = c4_inputs()
    Th = oracle.evaluate_library(ex, X)
    for name, alg in (("c4_k2145_admm", oracle.ADMM(lams, 1.0)), ("c4_k2145_sr3_hard", oracle.SR3(lams, 1.0))):
        traces = []
        res = oracle.sparse_regression(alg, Th, Y, oracle.CommonOptions(), traces=traces)
        sp = np.stack([np.stack(t.support_path) for t in traces])
        cp = np.stack([np.stack(t.coef_path) for t in traces])
        ip = np.stack([np.asarray(t.iters_path) for t in traces])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), thresholds=alg.thresholds, coefficients=res.coefficients,
                            lambda_opt=np.asarray(res.lambdas), iterations=np.asarray(res.iterations),
                            support_path=np.packbits(sp, axis=-1), coef_path=cp, iters_path=ip, k=ex.shape[1],
                            x_checksum=float(np.sum(X)), y_checksum=float(np.sum(Y)))
        print(name, "dof", [int(s.sum()) for s in sp[:, -1]], "iters", ip.sum(1).tolist())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "c4":
        c4_size()
    else:
        main()
