"""[historical: ran against a debug build that had a `probe` option; result: the tensor core TRUNCATES]
Does the tcgen05 tf32 path TRUNCATE or ROUND the low 13 bits of fp32 operands? (GPU probe)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import kmeans_oracle as ko
from spark_rapids_ml_b200 import _native
ctx = _native.Context(0)
ctx.set_option("kernel_path", 2)
n, d, k = 4096, 128, 64
X = ko.make_uniform(n, d, seed=3) + 0.5
C = X[:k].copy()
_, mdo, _ = ko.assign(X, C)
for probe in (0, 1, 2):
    ctx.set_option("probe", probe)
    lab, md = ctx.kmeans_assign(torch.from_numpy(X).cuda(), torch.from_numpy(C).cuda(), want_mindist=True)
    err = np.abs(md.cpu().numpy() - mdo).max() / (X.astype(np.float64) ** 2).sum(1).max()
    cmp = ko.compare_labels(X, C, lab.cpu().numpy())
    print(f"probe={probe}: max |md-md_oracle|/max||x||^2 = {err:.3e}  label mismatches={cmp['n_mismatch']} outside={cmp['n_mismatch_outside_margin']}")
print("probe=1 exact (error like probe=0)  => hardware TRUNCATES fp32 operands to tf32")
