"""Development tool: time the moment Gram schedule against the ring schedule and compare their blocks."""
import sys, os, json
sys.path.insert(0, "/root/repo")
import numpy as np
import ctypes as C
import ddb200
from ddb200 import _lib
import oracle

def run(n, deg, m, moment, reps=2):
    os.environ["DDB_GRAM_MOMENT"] = str(moment)
    lib = _lib.load()
    ex = np.asfortranarray(oracle.polynomial_exponents(n, deg))
    k = ex.shape[1]
    ctx = C.c_void_p()
    _lib.check(lib.ddb_ctx_create(0, C.byref(ctx)))
    _lib.check(lib.ddb_set_library_monomial(ctx, n, k, _lib.ptr(ex)))
    _lib.check(lib.ddb_generate(ctx, 1, n, m, 0, 0, 0.0))
    tm = _lib.Timings()
    best = 1e30
    G = np.zeros((k, k), order="F"); R = np.zeros((k, n), order="F"); yy = np.zeros(n)
    for r in range(reps):
        _lib.check(lib.ddb_gram(ctx, None))
        lib.ddb_get_timings(ctx, C.byref(tm))
        best = min(best, tm.gram_ms)
    _lib.check(lib.ddb_gram_download(ctx, _lib.ptr(G), _lib.ptr(R), _lib.ptr(yy)))
    F = k * (k + 1) + 2 * k * n + 2 * n
    print(json.dumps({"moment": moment, "n": n, "deg": deg, "k": k, "m": m, "gram_ms": best, "tflops_syrk": F * m / best * 1e-9}), flush=True)
    _lib.check(lib.ddb_ctx_destroy(ctx))
    return G, R, yy

if __name__ == "__main__":
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    deg = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    a = run(n, deg, m, 0)
    b = run(n, deg, m, 1)
    for x, y, name in zip(a, b, "GRy"):
        print(name, "max rel diff", float(np.max(np.abs(x - y) / (np.abs(x) + 1e-300))), "max abs", float(np.max(np.abs(x))))
