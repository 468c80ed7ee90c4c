"""Per-role blocked-cycle breakdown of the fused kernel (needs a GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
os.environ.setdefault("B2K_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spark_rapids_ml_b200", "libb2kmeans_trace.so"))
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
if _native.LIB_PATH.endswith("_trace.so"): P = P[:-4]   # diagnostic builds append a 4-CTA trace area (tools/trace_tiles.py)
roles = {"convert": range(0, 4), "epilogue": range(4, 8), "update": range(8, 24), "tma": [24], "mma": [25]}
names = {"convert": ["x_full", "a_empty", "xn_empty"], "epilogue": ["d_full", "xn_full", "lab_empty"],
         "update": ["lab_full"], "tma": ["x_empty"], "mma": ["d_empty", "a_full"]}
stage = {"convert": ["group(2 chunks)", "signalling"], "epilogue": ["ld+argmin", "sort+handoff"], "update": ["work", "bar2"]}
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
    print(line)
    if r in stage:   # stage timers pw[3], pw[4]: per warp, per tile (leader = first warp of the role)
        for i, nm in enumerate(stage[r]):
            v = sub[:, :, 4 + i] / tiles_per_cta
            print(f"      {nm:16s} leader={v[:, 0].mean():7.0f} (even CTAs {v[0::2, 0].mean():7.0f}, odd CTAs {v[1::2, 0].mean():7.0f}) mean={v.mean():7.0f} max-warp={v.mean(axis=0).max():7.0f}")
    if r in ("convert", "epilogue", "update"):
        lead = sub[:, 0, :]
        print("      leader blocked:", {nm: int(lead[:, 1 + i].mean() / tiles_per_cta) for i, nm in enumerate(names[r])})
    if len(sys.argv) > 2 and r in ("convert","mma","epilogue"):
        for par in (0,1):
            sp = P[par::2][:, list(ws), :]
            print("      rank", par, "role", int(sp[:,:,0].mean()/tiles_per_cta), "blocked", [int(sp[:,:,1+i].mean()/tiles_per_cta) for i in range(3)])
