"""Kernel-time split of the k-means|| initialiser (CUPTI through torch.profiler)."""
import sys, os, time, torch
sys.path.insert(0, os.getcwd())
from spark_rapids_ml_b200 import _native
from torch.profiler import profile, ProfilerActivity
n, d, k = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (10_000_000, 128, 64)
ctx = _native.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
ctr = torch.rand((k, d), generator=g, device="cuda") * 20 - 10
X = torch.empty((n, d), device="cuda")
for s in range(0, n, 1_000_000):
    e = min(n, s + 1_000_000)
    X[s:e] = ctr[torch.randint(0, k, (e - s,), generator=g, device="cuda")] + torch.randn((e - s, d), generator=g, device="cuda")
ctx.kmeans_fit(X, k, init="k-means||", max_iter=0, tol=1e-4, seed=1, compute_inertia=False)
torch.cuda.synchronize(); t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ctx.kmeans_fit(X, k, init="k-means||", max_iter=0, tol=1e-4, seed=1, compute_inertia=False)
    torch.cuda.synchronize()
print("wall %.1f ms (under the profiler)" % ((time.perf_counter() - t0) * 1e3))
tot = 0.0
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if ev.device_time_total > 0:
        tot += ev.device_time_total
        print("   %-70s n=%3d total=%8.3f ms" % (ev.key[:70], ev.count, ev.device_time_total / 1e3))
print("device total %.1f ms" % (tot / 1e3))
