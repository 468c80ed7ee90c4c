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
for probe in (0, 4):
    ctx.set_option("probe", probe)
    C = X[:k].clone()
    ctx.kmeans_lloyd(X, C, 3, -1.0)
    ctx.set_option("time_kernels", 1)
    ctx.kmeans_lloyd(X, C, 20, -1.0)
    print("probe", probe, "fused kernel ms", ctx.stats()["last_fused_ms"])
    ctx.set_option("time_kernels", 0)
ctx.set_option("profile_fused", 1)
ctx.kmeans_lloyd(X, C, 1, -1.0)
import numpy as np
P = ctx.fused_profile().astype(np.float64)
tpc = ((n + 127) // 128) / P.shape[0]
for nm, ws in {"convert": range(0, 4), "epilogue": range(4, 8), "update": range(8, 24), "tma": [24], "mma": [25]}.items():
    sub = P[:, list(ws), :]
    print(nm, "role", int(sub[:, :, 0].mean() / tpc), "blocked", [int(sub[:, :, 1 + i].mean() / tpc) for i in range(3)])
