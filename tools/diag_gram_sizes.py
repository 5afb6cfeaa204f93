import sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, ctypes as C
import ddb200
from ddb200 import _lib
import oracle
lib = _lib.load()
def run(n, deg, n_t, m):
    rng = np.random.default_rng(0)
    ex = np.asfortranarray(oracle.polynomial_exponents(n, deg)); k = ex.shape[1]
    X = np.asfortranarray(rng.standard_normal((n, m))); Y = np.asfortranarray(rng.standard_normal((n_t, m)))
    ctx = C.c_void_p(); _lib.check(lib.ddb_ctx_create(0, C.byref(ctx)))
    _lib.check(lib.ddb_set_library_monomial(ctx, n, k, _lib.ptr(ex)))
    _lib.check(lib.ddb_upload(ctx, _lib.ptr(X), _lib.ptr(Y), n_t, m))
    tm = _lib.Timings(); best = 1e9
    for r in range(3):
        _lib.check(lib.ddb_gram(ctx, None)); lib.ddb_get_timings(ctx, C.byref(tm)); best = min(best, tm.gram_ms)
    full_ms = m * 2.0 * 128 * 128 / 37e12 * 1e3
    print(json.dumps({"n": n, "k": k, "kaug": k + n_t, "m": m, "gram_ms": best, "full_pair_ms_at_peak": full_ms, "ratio": best / full_ms}))
    _lib.check(lib.ddb_ctx_destroy(ctx))
run(14, 2, 8, 4_000_000)    # kaug = 128: one diagonal pair
run(14, 2, 14, 4_000_000)   # kaug = 134: pairs (0,0) diag, (1,0) ragged 6 rows, (1,1) diag ragged
