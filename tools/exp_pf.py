import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spark_rapids_ml_b200 import _native
n, d, k = 10_000_000, 128, 64
ctx = _native.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
ctr = torch.rand((k, d), generator=g, device="cuda") * 20 - 10
X = torch.empty((n, d), device="cuda")
for s in range(0, n, 1_000_000):
    e = min(n, s + 1_000_000)
    X[s:e] = ctr[torch.randint(0, k, (e - s,), generator=g, device="cuda")] + torch.randn((e - s, d), generator=g, device="cuda")
for pf in [int(a) for a in sys.argv[1:]] or [3, -1, 0, 1]:
    ctx.set_option("pf_dist", pf)
    C = X[:k].clone()
    ctx.kmeans_lloyd(X, C, 3, -1.0)
    ctx.set_option("time_kernels", 1)
    ctx.kmeans_lloyd(X, C, 30, -1.0)
    print("pf_dist", pf, "fused kernel ms", round(ctx.stats()["last_fused_ms"], 4), flush=True)
    ctx.set_option("time_kernels", 0)
