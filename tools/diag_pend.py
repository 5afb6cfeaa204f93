import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, ddb200, oracle
from test_gpu_features import _pendulum_data, _theta
X, DX = _pendulum_data()
basis = ddb200.concat_bases([ddb200.polynomial_basis(2, 5), ddb200.sin_basis(2, 1), ddb200.cos_basis(2, 1)])
lams = np.arange(1e-2, 1e-1 + 1e-9, 1e-2)
alg = ddb200.B200Sparse(ddb200.STLSQ(lams))
sol = ddb200.solve(ddb200.ContinuousDataDrivenProblem(X, DX=DX), basis, alg, paths=True)
Th = _theta(basis, X)
print("cond Theta", np.linalg.cond(Th), "cond equilibrated", np.linalg.cond(Th / np.linalg.norm(Th, axis=1, keepdims=True)))
traces = []
ref = oracle.sparse_regression(oracle.STLSQ(lams), Th, DX, traces=traces)
for i, tr in enumerate(traces):
    sp = np.asarray(tr.support_path); cp = np.asarray(tr.coef_path)
    for j in range(len(lams)):
        d = np.nonzero(sol.paths.support_path[i][j] != sp[j])[0]
        if d.size:
            print("target", i, "lam", lams[j], "terms", d, "oracle coef", cp[j][d], "gpu coef", sol.paths.coef_path[i][j][d], "iters", tr.iters_path[j], sol.paths.iters_path[i][j])
            print("  oracle support", np.nonzero(sp[j])[0], "gpu", np.nonzero(sol.paths.support_path[i][j])[0])
