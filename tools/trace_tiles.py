"""Event trace of 16 consecutive tiles of CTAs 0/1 (needs a diagnostic build whose profile carries a trace area)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
os.environ.setdefault("B2K_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spark_rapids_ml_b200", "libb2kmeans_trace.so"))
from spark_rapids_ml_b200 import _native
n, d, k = 10_000_000, 128, 64
ctx = _native.Context(0)
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
P = ctx.fused_profile()
grid = P.shape[0] - 4
T = P[grid:].reshape(-1)[:512].reshape(2, 16, 16).astype(np.int64)
names = ["tma_c0", "tma_c3", "cvt_g2_start", "cvt_g2_done", "epi_dfull", "epi_argmin", "epi_lfull", "upd_start", "upd_done", "cvt_g1_start", "mma_c2", "mma_c3", "mma_commit", "mma_dempty", "mma_c0"]
for cta in range(2):
    t = T[cta]
    t0 = t[0, 0]
    print(f"CTA {cta}: per tile, cycles relative to the first tile's tma_c0")
    print("tile " + " ".join(f"{nm:>12s}" for nm in names))
    for i in range(16):
        print(f"{i:4d} " + " ".join(f"{int(t[i, e] - t0):12d}" for e in range(len(names))))
    d_ = lambda a, b: float(np.mean(t[2:, a] - t[2:, b]))
    print("mean: period", float(np.mean(np.diff(t[:, 8]))), " load(c3 issue->cvt g2 start)", d_(2, 1), " cvt g2", d_(3, 2), " mma tail+wake", d_(4, 3),
          " argmin", d_(5, 4), " sort", d_(6, 5), " handoff", d_(7, 6), " update", d_(8, 7), " issue c3 -> release", d_(8, 1))
