"""Dev timing of the DMD Gram kernel at config C5 scale (n = 4096 states, m snapshots)."""
import sys, json
sys.path.insert(0, "/root/repo")
import torch, ddb200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dev = torch.device("cuda:0")
xt = torch.empty((m + 1, n), dtype=torch.float64, device=dev)
chunk = 65536
for s in range(0, m + 1, chunk):
    xt[s:s + chunk].normal_()
PQ = torch.zeros((2, n, n), dtype=torch.float64, device=dev)
with ddb200.Context(0) as ctx:
    for r in range(3):
        ctx.dmd_gram(xt.data_ptr(), xt.data_ptr() + 8 * n, n, m, PQ[0].data_ptr(), PQ[1].data_ptr())
        ms = ctx.timings()["gram_ms"]
        F = n * (n + 1) + 2 * n * n
        print(json.dumps({"n": n, "m": m, "gram_ms": ms, "tflops_algorithmic": F * m / ms * 1e-9, "GBs": 8 * n * m / ms * 1e-6}), flush=True)
    # spot check against torch on a slab
    Pc = (xt[:4096].T @ xt[:4096]); 
    K = torch.zeros((n, n), dtype=torch.float64, device=dev)
    ctx.dmd_operator(PQ[0].data_ptr(), PQ[1].data_ptr(), n, K.data_ptr())
    print(json.dumps({"operator_ms": ctx.timings()["factor_ms"]}))
