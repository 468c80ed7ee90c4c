import sys, time, torch
sys.path.insert(0, ".")
from spark_rapids_ml_b200 import _native
ctx = _native.Context(0)
n, d, k = 12_500_000, 256, 256
g = torch.Generator(device="cuda").manual_seed(1)
ctr = torch.rand((k, d), generator=g, device="cuda") * 20 - 10
X = torch.empty((n, d), device="cuda")
for s in range(0, n, 1_000_000):
    e = min(n, s + 1_000_000)
    X[s:e] = ctr[torch.randint(0, k, (e - s,), generator=g, device="cuda")] + torch.randn((e - s, d), generator=g, device="cuda")
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.kmeans_fit(X, k, init="k-means||", max_iter=0, tol=1e-4, seed=1 + rep, compute_inertia=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = ctx.kmeans_fit(X, k, init="k-means||", max_iter=20, tol=-1.0, seed=1 + rep)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep {rep}: init only {t1-t0:.3f} s; full fit (init + 20 iterations + inertia) {t2-t1:.3f} s", flush=True)
