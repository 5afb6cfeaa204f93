"""Profiling driver: one Gram launch at the Lorenz-96 shape (n=64, degree 2) or the Lorenz shape."""
import sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, ctypes as C
import ddb200
from ddb200 import _lib
import oracle
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = int(sys.argv[3]) if len(sys.argv) > 3 else 200_000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ex = np.asfortranarray(oracle.polynomial_exponents(n, deg)); k = ex.shape[1]
ctx = C.c_void_p(); _lib.check(lib.ddb_ctx_create(0, C.byref(ctx)))
_lib.check(lib.ddb_set_library_monomial(ctx, n, k, _lib.ptr(ex)))
_lib.check(lib.ddb_generate(ctx, 0 if n == 3 else 1, n, m, 0, 0, 0.0))
tm = _lib.Timings()
for r in range(reps):
    _lib.check(lib.ddb_gram(ctx, None)); lib.ddb_get_timings(ctx, C.byref(tm))
    F = k * (k + 1) + 2 * k * n + 2 * n
    print(json.dumps({"n": n, "k": k, "m": m, "gram_ms": tm.gram_ms, "tflops_syrk": F * m / tm.gram_ms * 1e-9}), flush=True)
_lib.check(lib.ddb_ctx_destroy(ctx))
