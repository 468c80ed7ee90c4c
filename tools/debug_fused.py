"""GPU bring-up script for the tcgen05 fused kernel: small shapes first, verbose diagnostics.
Run under `timeout` (a protocol bug traps after ~2 s instead of hanging)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import kmeans_oracle as ko
from spark_rapids_ml_b200 import _native

def run(n, d, k, gen="blobs", path=2):
    ctx = _native.Context(0)
    ctx.set_option("kernel_path", path)
    X = ko.make_blobs(n, d, k, seed=7)[0] if gen == "blobs" else ko.make_uniform(n, d, seed=7)
    rng = np.random.default_rng(3)
    C = X[rng.choice(n, size=k, replace=(n < k))].copy()
    Xd = torch.from_numpy(X).cuda(); Cd = torch.from_numpy(C).cuda()
    t0 = time.time()
    labels, md = ctx.kmeans_assign(Xd, Cd, want_mindist=True)
    torch.cuda.synchronize()
    lab = labels.cpu().numpy(); mdg = md.cpu().numpy()
    lo, mdo, margin = ko.assign(X, C)
    mism = np.nonzero(lab != lo)[0]
    out = int((margin[mism] >= 1e-6).sum())
    relerr = np.abs(mdg - mdo).max() / max(1e-30, (X.astype(np.float64) ** 2).sum(1).max())
    print(f"[assign n={n} d={d} k={k} {gen} path={path}] mismatch={mism.size} outside_margin={out} "
          f"md_err_rel_xn={relerr:.2e} t={time.time()-t0:.2f}s", flush=True)
    if mism.size and out:
        for i in mism[:8]:
            print("   row", i, "gpu", lab[i], "oracle", lo[i], "margin", margin[i], "md", mdg[i], mdo[i])
    # one lloyd step
    C1 = Cd.clone()
    n_it, shift = ctx.kmeans_lloyd(Xd, C1, 1, -1.0)
    # expected next centers from the DEVICE labels (tie rows inside the 1e-6 margin may legitimately differ from the oracle)
    S, w = ko.partial_sums(X, lab, k)
    Cn = C.astype(np.float64); Cn[w > 0] = S[w > 0] / w[w > 0][:, None]; Cn = Cn.astype(np.float32)
    sh = float(((Cn.astype(np.float64) - C.astype(np.float64)) ** 2).sum())
    err = ko.max_center_rel_err(C1.cpu().numpy(), Cn)
    print(f"   lloyd step: n_it={n_it} shift={shift:.6e} oracle_shift={sh:.6e} center_rel_err={err:.2e} stats={ctx.stats()}", flush=True)
    ctx.close()
    return out == 0 and err < 1e-4

if __name__ == "__main__":
    shapes = [(128, 128, 64), (256, 128, 64), (1000, 128, 64), (40000, 128, 64), (1000, 32, 8), (130, 20, 5), (3000, 100, 40), (5000, 64, 128)]
    ok = True
    for (n, d, k) in shapes:
        for gen in ("blobs", "uniform"):
            ok &= run(n, d, k, gen)
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)
