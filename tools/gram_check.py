"""Quick GPU check of the Gram stage against NumPy (development tool; the real tests are in tests/)."""
import sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np
import ddb200
from ddb200 import _lib
import oracle

lib = _lib.load()
import ctypes as C

def run(n, deg, m, n_t=None, seed=0, reps=3):
    rng = np.random.default_rng(seed)
    n_t = n if n_t is None else n_t
    X = np.asfortranarray(rng.standard_normal((n, m)))
    Y = np.asfortranarray(rng.standard_normal((n_t, m)))
    ex = np.asfortranarray(oracle.polynomial_exponents(n, deg))
    k = ex.shape[1]
    ctx = C.c_void_p()
    _lib.check(lib.ddb_ctx_create(0, C.byref(ctx)))
    _lib.check(lib.ddb_set_library_monomial(ctx, n, k, _lib.ptr(ex)))
    _lib.check(lib.ddb_upload(ctx, _lib.ptr(X), _lib.ptr(Y), n_t, m))
    tm = _lib.Timings()
    best = 1e30
    for r in range(reps):
        _lib.check(lib.ddb_gram(ctx, None))
        lib.ddb_get_timings(ctx, C.byref(tm))
        best = min(best, tm.gram_ms)
    G = np.zeros((k, k), order="F"); R = np.zeros((k, n_t), order="F"); yy = np.zeros(n_t)
    _lib.check(lib.ddb_gram_download(ctx, _lib.ptr(G), _lib.ptr(R), _lib.ptr(yy)))
    out = {"n": n, "deg": deg, "k": k, "m": m, "gram_ms": best}
    if k * m <= 4e8:
        Th = oracle.evaluate_library(ex, X)
        G0 = Th @ Th.T; R0 = Th @ Y.T; yy0 = (Y * Y).sum(1)
        sc = np.sqrt(np.outer(np.diag(G0), np.diag(G0)))
        out["errG"] = float(np.max(np.abs(G - G0) / sc))
        out["errR"] = float(np.max(np.abs(R - R0)) / np.max(np.abs(R0)))
        out["erryy"] = float(np.max(np.abs(yy - yy0) / yy0))
        out["symm"] = float(np.max(np.abs(G - G.T)))
    F = k * (k + 1) + 2 * k * n_t + 2 * n_t
    out["tflops_syrk"] = F * m / best * 1e-9
    _lib.check(lib.ddb_ctx_destroy(ctx))
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    run(3, 2, 1001)
    run(3, 5, 1001)
    run(3, 5, 100000)
    run(5, 3, 5000)
    run(64, 2, 3000)
    run(64, 2, 20000)
    run(3, 5, 10_000_000, reps=3)
    run(64, 2, 1_000_000, reps=2)
