"""Bring-up of the CTA-pair (tcgen05 cta_group::2) instantiation: parity vs the single-CTA kernel + oracle, then speed."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import kmeans_oracle as ko
from spark_rapids_ml_b200 import _native
ctx = _native.Context(0)
ctx.set_option("kernel_path", 2)
ok = True
for n in (128, 256, 300, 1000, 40000, 300000):
    X, _ = ko.make_blobs(n, 128, 64, seed=7)
    C = X[np.random.default_rng(3).choice(n, 64, replace=False)].copy()
    Xd, Cd = torch.from_numpy(X).cuda(), torch.from_numpy(C).cuda()
    res = {}
    for pair in (0, 1):
        ctx.set_option("pair", pair)
        lab, md = ctx.kmeans_assign(Xd, Cd, want_mindist=True)
        C1 = Cd.clone()
        ctx.kmeans_lloyd(Xd, C1, 2, -1.0)
        res[pair] = (lab.cpu().numpy(), md.cpu().numpy(), C1.cpu().numpy())
    cmp = ko.compare_labels(X, C, res[1][0])
    same = np.array_equal(res[0][0], res[1][0])
    cerr = ko.max_center_rel_err(res[1][2], res[0][2])
    print(f"n={n}: pair labels==single {same}  outside_margin={cmp['n_mismatch_outside_margin']}  centers rel diff after 2 iters {cerr:.2e}", flush=True)
    ok &= cmp["n_mismatch_outside_margin"] == 0 and cerr < 1e-5
print("PAIR PARITY", "OK" if ok else "FAILED", flush=True)
n = 10_000_000
g = torch.Generator(device="cuda").manual_seed(1)
ctr = torch.rand((64, 128), generator=g, device="cuda") * 20 - 10
X = torch.empty((n, 128), device="cuda")
for s in range(0, n, 1_000_000):
    e = min(n, s + 1_000_000)
    X[s:e] = ctr[torch.randint(0, 64, (e - s,), generator=g, device="cuda")] + torch.randn((e - s, 128), generator=g, device="cuda")
for pair in (0, 1):
    ctx.set_option("pair", pair)
    C = X[:64].clone()
    ctx.kmeans_lloyd(X, C, 3, -1.0)
    ctx.set_option("time_kernels", 1)
    ctx.kmeans_lloyd(X, C, 30, -1.0)
    print("pair", pair, "fused kernel ms", round(ctx.stats()["last_fused_ms"], 4), flush=True)
    ctx.set_option("time_kernels", 0)
