"""Gram-stage time of the C2 shape (Lorenz, degree 5, 1e7 samples) only: python tools/gram_c2_time.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddb200, oracle
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
with ddb200.Context(0) as ctx:
    ctx.set_library(oracle.polynomial_exponents(3, 5))
    ctx.generate(0, 3, 10_000_000)
    ts = []
    for _ in range(reps):
        ctx.gram()
        ts.append(ctx.timings()["gram_ms"])
    print("DDB_GRAM_POW=%s DDB_GRAM_W4=%s" % (os.environ.get("DDB_GRAM_POW"), os.environ.get("DDB_GRAM_W4")),
          "gram_ms min %.4f median %.4f" % (min(ts), float(np.median(ts))), "schedule", ctx.timings()["gram_schedule"])
