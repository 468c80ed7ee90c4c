"""A/B helper: fused-kernel ms for pair=0/1 on cfg2 (same box, same process)."""
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
# reference point for box-to-box variation: the TMA streaming microbenchmark
print("tma stream 9 slots: %.3f ms" % ctx.debug_tma_stream(X, 9, 0))
for pair in (1, 0, 1):
    ctx.set_option("pair", pair)
    C = X[:k].clone()
    ctx.kmeans_lloyd(X, C, 3, -1.0)
    ctx.set_option("time_kernels", 1)
    ctx.kmeans_lloyd(X, C, 40, -1.0)
    print("pair", pair, "fused kernel ms", round(ctx.stats()["last_fused_ms"], 4), flush=True)
    ctx.set_option("time_kernels", 0)
