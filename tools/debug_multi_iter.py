import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import kmeans_oracle as ko
from spark_rapids_ml_b200 import _native
ctx = _native.Context(0)
X, tc = ko.make_blobs(20000, 128, 64, seed=3)
C0 = (tc + 0.25 * np.random.default_rng(0).normal(size=tc.shape)).astype(np.float32)
Xd = torch.from_numpy(X).cuda()
for pair in (1, 0):
    ctx.set_option("pair", pair)
    # oracle trajectory
    Co = C0.copy(); Cs = torch.from_numpy(C0).cuda(); Cm = torch.from_numpy(C0).cuda()
    ctx.kmeans_lloyd(Xd, Cm, 3, -1.0)          # one call, 3 iterations (balance table active from iteration 2)
    for it in range(3):
        Co, w, sh = ko.lloyd_iteration([X], Co)
        ctx.kmeans_lloyd(Xd, Cs, 1, -1.0)      # fresh call per iteration (identity table)
        print(f"pair={pair} iter {it+1}: single-step calls vs oracle {ko.max_center_rel_err(Cs.cpu().numpy(), Co):.2e}")
    print(f"pair={pair} 3-iteration call vs oracle {ko.max_center_rel_err(Cm.cpu().numpy(), Co):.2e}")
    lab, _ = ctx.kmeans_assign(Xd, torch.from_numpy(Co).cuda())
    print("   labels vs oracle:", ko.compare_labels(X, Co, lab.cpu().numpy()))
