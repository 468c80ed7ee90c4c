import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import kmeans_oracle as ko
from spark_rapids_ml_b200 import _native
ctx = _native.Context(0)
g = np.load("tests/golden/lloyd_golden_blobs_d128_k64.npz")
X, C0 = g["X"], g["C0"]
lab, _, _ = ko.assign(X, C0)
def run(Xs, labs, tag):
    Ss, ws = ko.partial_sums(Xs, labs, 64)
    exp = C0.astype(np.float64).copy(); nz = ws > 0; exp[nz] = Ss[nz] / ws[nz][:, None]
    Cg = torch.from_numpy(C0).cuda()
    ctx.kmeans_lloyd(torch.from_numpy(np.ascontiguousarray(Xs)).cuda(), Cg, 1, -1.0)
    Cg = Cg.cpu().numpy().astype(np.float64)
    err = np.linalg.norm(Cg - exp, axis=1) / np.maximum(np.linalg.norm(exp, axis=1), 1e-30)
    bad = np.nonzero(err > 1e-5)[0]
    if len(bad): print(tag, "BAD clusters", bad.tolist(), "sizes", ws[bad].astype(int).tolist(), "hist nonzero", np.count_nonzero(ws), "max", int(ws.max()))
    return len(bad) > 0
for pair in (1, 0):
    ctx.set_option("pair", pair)
    print("pair", pair)
    for t in range(32):
        run(X[t*128:(t+1)*128], lab[t*128:(t+1)*128], f"tile {t}")
    for nt in (32,):
        run(X[:nt*128], lab[:nt*128], f"first {nt} tiles")
