"""A/B of the Gram stage at the C2 shape (Lorenz, degree 5, 1e7 samples): prefix-product builder against the factor builder."""
import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, ctypes as C
import ddb200
from ddb200 import _lib
import oracle
lib = _lib.load()
n, deg, m = 3, 5, 10_000_000
ex = np.asfortranarray(oracle.polynomial_exponents(n, deg)); k = ex.shape[1]
ctx = C.c_void_p(); _lib.check(lib.ddb_ctx_create(0, C.byref(ctx)))
_lib.check(lib.ddb_set_library_monomial(ctx, n, k, _lib.ptr(ex)))
_lib.check(lib.ddb_generate(ctx, 0, n, m, 0, 0, 0.0))
tm = _lib.Timings()
out = {}
for tr in ("1", "0", "1", "0"):
    os.environ["DDB_GRAM_TREE"] = tr
    best = 1e30
    for r in range(4):
        _lib.check(lib.ddb_gram(ctx, None)); lib.ddb_get_timings(ctx, C.byref(tm))
        best = min(best, tm.gram_ms)
    out.setdefault("tree" + tr, []).append(round(best, 4))
print(json.dumps({"n": n, "k": k, "m": m, **out}))
_lib.check(lib.ddb_ctx_destroy(ctx))
