"""A/B of the Gram stage at the C3 shape: tensor-pipe builder (DDB_GRAM_MB=1) against the FP64-pipe builder (=0)."""
import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, ctypes as C
import ddb200
from ddb200 import _lib
import oracle
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
ex = np.asfortranarray(oracle.polynomial_exponents(n, 2)); k = ex.shape[1]
ctx = C.c_void_p(); _lib.check(lib.ddb_ctx_create(0, C.byref(ctx)))
_lib.check(lib.ddb_set_library_monomial(ctx, n, k, _lib.ptr(ex)))
_lib.check(lib.ddb_generate(ctx, 1, n, m, 0, 0, 0.0))
tm = _lib.Timings()
out = {}
for mb in ("1", "0", "1", "0"):
    os.environ["DDB_GRAM_MB"] = mb
    best = 1e30
    for r in range(3):
        _lib.check(lib.ddb_gram(ctx, None)); lib.ddb_get_timings(ctx, C.byref(tm))
        best = min(best, tm.gram_ms)
    out.setdefault("mb" + mb, []).append(round(best, 3))
print(json.dumps({"n": n, "k": k, "m": m, **out}))
_lib.check(lib.ddb_ctx_destroy(ctx))
