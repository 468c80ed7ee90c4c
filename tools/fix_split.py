"""Kernel-time split of the large-shape pass (main kernel vs the deferred rows' fix-up kernels) under CUPTI."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
from spark_rapids_ml_b200 import _native
from torch.profiler import profile, ProfilerActivity
n, d, k = 3_000_000, 256, 256
ctx = _native.Context(0)
ctx.set_option("collect_recheck", 1)
g = torch.Generator(device="cuda").manual_seed(1)
ctr = torch.rand((k, d), generator=g, device="cuda") * 20 - 10
X = torch.empty((n, d), device="cuda")
for s in range(0, n, 1_000_000):
    e = min(n, s + 1_000_000)
    X[s:e] = ctr[torch.randint(0, k, (e - s,), generator=g, device="cuda")] + torch.randn((e - s, d), generator=g, device="cuda")
C0 = (ctr + 0.25 * torch.randn((k, d), generator=g, device="cuda")).contiguous()
C1 = X[:k].clone()
# light: 0.5 % of the rows sit at the midpoint of two centres (near-ties between exactly two clusters)
XL = X.clone()
idx = torch.arange(0, n, 200, device="cuda")
pa = torch.randint(0, k, (idx.numel(),), generator=g, device="cuda")
pb = (pa + 1 + torch.randint(0, k - 1, (idx.numel(),), generator=g, device="cuda")) % k
XL[idx] = 0.5 * (C0[pa] + C0[pb]) + 0.001 * torch.randn((idx.numel(), d), generator=g, device="cuda")
cases = [("near_true", X, C0), ("light", XL, C0), ("first_k", X, C1)]
if len(sys.argv) > 1 and sys.argv[1] == "uniform":
    U = torch.rand((n, d), generator=g, device="cuda")
    cases.append(("uniform", U, U[:k].clone()))
for name, XX, C in cases:
    Cc = C.clone(); ctx.kmeans_lloyd(XX, Cc, 1, -1.0)
    Cc = C.clone()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):     # the SAME first iteration four times (the centres move away from the near-ties otherwise)
            Cc = C.clone(); ctx.kmeans_lloyd(XX, Cc, 1, -1.0)
        torch.cuda.synchronize()
    st = ctx.stats()
    print(name, "recheck_rows/iter", st["recheck_rows"], "cand/row", st["recheck_candidates"] / max(1, st["recheck_rows"]))
    for ev in prof.key_averages():
        if ev.device_time_total > 0:
            print("   %-60s n=%d avg=%.3f ms" % (ev.key[:60], ev.count, ev.device_time_total / ev.count / 1e3))
if len(sys.argv) > 1 and sys.argv[1] == "uniform":   # the adaptive switch: 12 iterations on uniform noise, both settings
    ctx.set_option("time_kernels", 1)
    name, XX, C = cases[-1]
    for adaptive in (1, 0, 1):
        ctx.set_option("adaptive_path", adaptive)
        Cc = C.clone(); ctx.kmeans_lloyd(XX, Cc, 12, -1.0)
        st = ctx.stats()
        print("uniform 12 iterations adaptive=%d: loop %.1f ms, switch at %d, last_path %d" % (adaptive, st["last_loop_ms"], st["path_switch_iter"], st["last_path"]))
    ctx.set_option("kernel_path", 1)
    Cc = C.clone(); ctx.kmeans_lloyd(XX, Cc, 12, -1.0)
    print("uniform 12 iterations generic only: loop %.1f ms" % ctx.stats()["last_loop_ms"])
    for nm, XX2, C2 in cases[:3]:
        Cc = C2.clone(); ctx.kmeans_lloyd(XX2, Cc, 12, -1.0)
        print(nm, "12 iterations generic only: loop %.1f ms" % ctx.stats()["last_loop_ms"])
