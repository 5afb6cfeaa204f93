"""Development tool: time the ring Gram schedule against the self-building one at config C3."""
import sys, os, json
sys.path.insert(0, "/root/repo")
import numpy as np
import ctypes as C
import ddb200
from ddb200 import _lib
import oracle

def run(n, deg, m, ring, reps=2):
    os.environ["DDB_GRAM_RING"] = str(ring)
    lib = _lib.load()
    ex = np.asfortranarray(oracle.polynomial_exponents(n, deg))
    k = ex.shape[1]
    ctx = C.c_void_p()
    _lib.check(lib.ddb_ctx_create(0, C.byref(ctx)))
    _lib.check(lib.ddb_set_library_monomial(ctx, n, k, _lib.ptr(ex)))
    _lib.check(lib.ddb_generate(ctx, 1, n, m, 0, 0, 0.0))
    tm = _lib.Timings()
    best = 1e30
    for r in range(reps):
        _lib.check(lib.ddb_gram(ctx, None))
        lib.ddb_get_timings(ctx, C.byref(tm))
        best = min(best, tm.gram_ms)
    F = k * (k + 1) + 2 * k * n + 2 * n
    print(json.dumps({"ring": ring, "m": m, "gram_ms": best, "tflops_syrk": F * m / best * 1e-9}), flush=True)
    _lib.check(lib.ddb_ctx_destroy(ctx))

if __name__ == "__main__":
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    run(64, 2, m, 0)
    run(64, 2, m, 1, reps=1)
