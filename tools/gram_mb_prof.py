import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, ctypes as C
import ddb200
from ddb200 import _lib
import oracle
lib = _lib.load()
n, m = 64, 300_000
ex = np.asfortranarray(oracle.polynomial_exponents(n, 2)); k = ex.shape[1]
ctx = C.c_void_p(); _lib.check(lib.ddb_ctx_create(0, C.byref(ctx)))
_lib.check(lib.ddb_set_library_monomial(ctx, n, k, _lib.ptr(ex)))
_lib.check(lib.ddb_generate(ctx, 1, n, m, 0, 0, 0.0))
for mb in ("1", "0"):
    os.environ["DDB_GRAM_MB"] = mb
    _lib.check(lib.ddb_gram(ctx, None))
_lib.check(lib.ddb_ctx_destroy(ctx))
