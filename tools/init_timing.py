import sys, time, torch
sys.path.insert(0, ".")
from spark_rapids_ml_b200 import _native
ctx = _native.Context(0)
for (n, d, k) in ((2_000_000, 256, 256), (2_000_000, 128, 1024), (12_500_000, 256, 256), (10_000_000, 128, 64)):
    g = torch.Generator(device="cuda").manual_seed(1)
    ctr = torch.rand((k, d), generator=g, device="cuda") * 20 - 10
    X = (ctr[torch.randint(0, k, (n,), generator=g, device="cuda")] + torch.randn((n, d), generator=g, device="cuda")).contiguous()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = ctx.kmeans_fit(X, k, init="k-means||", max_iter=0, tol=1e-4, seed=1, compute_inertia=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out2 = ctx.kmeans_fit(X, k, init="k-means||", max_iter=10, tol=1e-4, seed=1)
    print(f"n={n} d={d} k={k}: k-means|| init {t1-t0:.2f} s; after 10 Lloyd iterations inertia/n = {out2['inertia_']/n:.2f} (blobs sigma^2*d = {d})", flush=True)
