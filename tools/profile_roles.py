"""Per-role blocked-cycle breakdown of the fused kernel (needs a GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spark_rapids_ml_b200 import _native
n, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 128, 64
ctx = _native.Context(0)
if len(sys.argv) > 2: ctx.set_option("pair", int(sys.argv[2]))
g = torch.Generator(device="cuda").manual_seed(1)
ctr = torch.rand((k, d), generator=g, device="cuda") * 20 - 10
X = torch.empty((n, d), device="cuda")
for s in range(0, n, 1_000_000):
    e = min(n, s + 1_000_000)
    X[s:e] = ctr[torch.randint(0, k, (e - s,), generator=g, device="cuda")] + torch.randn((e - s, d), generator=g, device="cuda")
C = X[:k].clone()
ctx.kmeans_lloyd(X, C, 3, -1.0)
ctx.set_option("profile_fused", 1)
ctx.kmeans_lloyd(X, C, 1, -1.0)
P = ctx.fused_profile().astype(np.float64)
roles = {"convert": range(0, 4), "epilogue": range(4, 8), "update": range(8, 24), "tma": [24], "mma": [25]}
names = {"convert": ["x_full", "a_empty", "xn_empty"], "epilogue": ["d_full", "xn_full", "lab_empty"],
         "update": ["lab_full", "x_full"], "tma": ["x_empty"], "mma": ["d_empty", "a_full"]}
ntiles = (n + 127) // 128
tiles_per_cta = ntiles / P.shape[0]
print(f"grid={P.shape[0]} tiles/CTA={tiles_per_cta:.1f}")
for r, ws in roles.items():
    sub = P[:, list(ws), :]
    tot = sub[:, :, 0].mean()
    line = f"{r:9s} role_cycles/tile={tot / tiles_per_cta:8.0f}  blocked:"
    blocked = 0
    for i, nm in enumerate(names[r]):
        v = sub[:, :, 1 + i].mean()
        blocked += v
        line += f" {nm}={v / tiles_per_cta:7.0f}"
    line += f"  busy={(tot - blocked) / tiles_per_cta:7.0f}"
    if r == "update":
        per_warp = (sub[:, :, 0] - sub[:, :, 1] - sub[:, :, 2]).mean(axis=0) / tiles_per_cta
        line += "  busy/warp=" + ",".join(f"{v:.0f}" for v in per_warp)
    print(line)
    if len(sys.argv) > 2 and r in ("convert","mma","epilogue"):
        for par in (0,1):
            sp = P[par::2][:, list(ws), :]
            print("      rank", par, "role", int(sp[:,:,0].mean()/tiles_per_cta), "blocked", [int(sp[:,:,1+i].mean()/tiles_per_cta) for i in range(3)])
