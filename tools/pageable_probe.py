"""End-to-end C3 solve from PAGEABLE host arrays for several bounce-ring thread counts (DDB_STAGE_THREADS)."""
import os, sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np
import ddb200, oracle
import torch
n, m = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = ddb200.Context(0)
ctx.set_library(oracle.polynomial_exponents(n, 2))
ctx.generate(1, n, m, first_sample=0, seed=0, noise_std=0.0)
X = np.empty((n, m), order="F"); Y = np.empty((n, m), order="F")
ddb200._lib.check(ctx._lib.ddb_download_data(ctx._ctx, X.ctypes.data, Y.ctypes.data))
prm = ctx.make_params("STLSQ")
lams = 10.0 ** np.arange(-4.0, 0.01, 0.1)
print("host cores:", os.cpu_count())
for t in (1, 2, 4, 8, 12, 16):
    os.environ["DDB_STAGE_THREADS"] = str(t)
    ts = []
    for r in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = ctx.solve_sparse(X, Y, prm, lams)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(json.dumps({"threads": t, "ms": [round(1e3 * x, 1) for x in ts], "h2d_ms": round(ctx.timings()["h2d_ms"], 1)}), flush=True)
