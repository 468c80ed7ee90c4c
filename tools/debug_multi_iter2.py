import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import kmeans_oracle as ko
from spark_rapids_ml_b200 import _native
ctx = _native.Context(0)
g = np.load("tests/golden/lloyd_golden_blobs_d128_k64.npz")
X, C0 = g["X"], g["C0"]
Xd = torch.from_numpy(X).cuda()
for pair in (1, 0):
    ctx.set_option("pair", pair)
    for ce in (1, 4):
        ctx.set_option("check_every", ce)
        Co = C0.copy(); Cm = torch.from_numpy(C0).cuda()
        hist = []
        for it in range(8):
            Co, w, sh = ko.lloyd_iteration([X], Co)
            Cg = torch.from_numpy(C0).cuda()
            ctx.kmeans_lloyd(Xd, Cg, it + 1, -1.0)   # one call with it+1 iterations
            lab, _ = ctx.kmeans_assign(Xd, torch.from_numpy(Co).cuda())
            cmp = ko.compare_labels(X, Co, lab.cpu().numpy())
            hist.append((it + 1, ko.max_center_rel_err(Cg.cpu().numpy(), Co), cmp["n_mismatch"], cmp["n_mismatch_outside_margin"]))
        print(f"pair={pair} check_every={ce}:", " ".join(f"[{a}:{b:.1e},{c},{d}]" for a, b, c, d in hist), flush=True)
# inertia path
X2, tc = ko.make_blobs(20000, 128, 64, seed=3)
C2 = (tc + 0.25 * np.random.default_rng(0).normal(size=tc.shape)).astype(np.float32)
ref = ko.lloyd([X2], C2, 5, 1e-4)
out = ctx.kmeans_fit(torch.from_numpy(X2).cuda(), 64, init=C2, max_iter=5, tol=1e-4)
print("fit n_iter", out["n_iter_"], ref["n_iter"], "inertia", out["inertia_"], ref["inertia"], "cerr", ko.max_center_rel_err(out["cluster_centers_"].cpu().numpy(), ref["centers"]))
lab, md = ctx.kmeans_assign(torch.from_numpy(X2).cuda(), out["cluster_centers_"], want_mindist=True)
_, mdo, _ = ko.assign(X2, ref["centers"])
print("mindist sum gpu", float(md.double().sum()), "oracle", mdo.sum(), "max abs diff", float(np.abs(md.cpu().numpy() - mdo).max()))
