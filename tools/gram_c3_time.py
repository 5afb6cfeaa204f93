"""Gram stage at the C3 shape (n = 64, degree 2) for m samples: best of 3."""
import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, ctypes as C
import ddb200
from ddb200 import _lib
import oracle
if os.environ.get('DDB_LIB'): _lib.LIB_PATH = os.environ['DDB_LIB']
lib = _lib.load()
n = 64
m = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ex = np.asfortranarray(oracle.polynomial_exponents(n, 2)); k = ex.shape[1]
ctx = C.c_void_p(); _lib.check(lib.ddb_ctx_create(0, C.byref(ctx)))
_lib.check(lib.ddb_set_library_monomial(ctx, n, k, _lib.ptr(ex)))
_lib.check(lib.ddb_generate(ctx, 1, n, m, 0, 0, 0.0))
tm = _lib.Timings()
best = 1e30
for r in range(4):
    _lib.check(lib.ddb_gram(ctx, None)); lib.ddb_get_timings(ctx, C.byref(tm))
    best = min(best, tm.gram_ms)
print(json.dumps({"n": n, "k": k, "m": m, "gram_ms": round(best, 3), "samples_per_s": m / best * 1e3}))
_lib.check(lib.ddb_ctx_destroy(ctx))
