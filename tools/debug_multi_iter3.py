import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import kmeans_oracle as ko
from spark_rapids_ml_b200 import _native
ctx = _native.Context(0)
g = np.load("tests/golden/lloyd_golden_blobs_d128_k64.npz")
X, C0 = g["X"], g["C0"]
lab, _, _ = ko.assign(X, C0)
S, w = ko.partial_sums(X, lab, 64)
for n in (4096, 2048, 1024, 512, 256, 128):
    Xs = X[:n]; labs = lab[:n]
    Ss, ws = ko.partial_sums(Xs, labs, 64)
    exp = C0.astype(np.float64).copy(); nz = ws > 0; exp[nz] = Ss[nz] / ws[nz][:, None]
    Cg = torch.from_numpy(C0).cuda()
    ctx.kmeans_lloyd(torch.from_numpy(Xs.copy()).cuda(), Cg, 1, -1.0)
    Cg = Cg.cpu().numpy().astype(np.float64)
    err = np.linalg.norm(Cg - exp, axis=1) / np.maximum(np.linalg.norm(exp, axis=1), 1e-30)
    bad = np.nonzero(err > 1e-5)[0]
    # per-tile max rows of one cluster
    mx = max(np.bincount(labs[t*128:(t+1)*128], minlength=64).max() for t in range((n + 127)//128))
    print(f"n={n}: max err {err.max():.2e}; bad clusters {bad.tolist()[:10]} sizes {ws[bad].astype(int).tolist()[:10]}; max rows of a cluster in one tile = {mx}", flush=True)
    if len(bad):
        j = bad[0]
        # infer implied sum difference in units of rows
        diff = Cg[j] * ws[j] - Ss[j]
        # which single row best explains the difference (missing or extra)?
        d2 = ((Xs.astype(np.float64) - diff) ** 2).sum(1); d3 = ((Xs.astype(np.float64) + diff) ** 2).sum(1)
        print("   cluster", j, "diff norm", np.linalg.norm(diff), "closest extra-row", int(d2.argmin()), float(d2.min()), "label", int(labs[d2.argmin()]), "closest missing-row", int(d3.argmin()), float(d3.min()), "label", int(labs[d3.argmin()]))
