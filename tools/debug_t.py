"""Diagnostic driver for the large-shape kernel (b2k_fused_t.cu): prints mismatch details instead of asserting."""
import sys, time
import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import kmeans_oracle as ko
from spark_rapids_ml_b200 import _native


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def run_assign(ctx, n, d, k, gen, force=False):
    X = ko.make_blobs(n, d, k, seed=7)[0] if gen == "blobs" else ko.make_uniform(n, d, seed=7)
    rng = np.random.default_rng(3)
    C = X[rng.choice(n, size=k, replace=(n < k))].copy() + (0.01 if gen == "uniform" else 0.0)
    ctx.set_option("variant_t", 1 if force else 0)
    ctx.set_option("collect_recheck", 1)
    t0 = time.time()
    labels, md = ctx.kmeans_assign(dev(X), dev(C), want_mindist=True)
    torch.cuda.synchronize()
    st = ctx.stats()
    lab = labels.cpu().numpy()
    cmp = ko.compare_labels(X, C, lab, tau=1e-6)
    lo, md_o, _ = ko.assign(X, C)
    mdg = md.cpu().numpy()
    rel = np.abs(mdg - md_o) / np.maximum(md_o, 1e-6)
    print(f"assign n={n} d={d} k={k} {gen}: path={st['last_path']} mismatch={cmp['n_mismatch']} outside={cmp['n_mismatch_outside_margin']} "
          f"recheck_rows={st['recheck_rows']} cand={st['recheck_candidates']} md_relerr_max={rel.max():.2e} t={time.time()-t0:.2f}s", flush=True)
    if cmp["n_mismatch_outside_margin"]:
        bad = np.nonzero(lab != lo)[0][:10]
        print("   first bad rows", bad, "gpu", lab[bad], "oracle", lo[bad], flush=True)
    return cmp["n_mismatch_outside_margin"] == 0


def run_lloyd(ctx, n, d, k, iters, gen="blobs", force=False):
    X, ctr = ko.make_blobs(n, d, k, seed=11)
    if gen == "uniform":
        X = ko.make_uniform(n, d, seed=11)
        C0 = X[:k].copy()
    else:
        C0 = (ctr + 0.25 * np.random.default_rng(0).normal(size=ctr.shape)).astype(np.float32)
    ctx.set_option("variant_t", 1 if force else 0)
    ref = ko.lloyd([X], C0, iters, -1.0)
    C = dev(C0)
    n_it, shift = ctx.kmeans_lloyd(dev(X), C, iters, -1.0)
    st = ctx.stats()
    err = ko.max_center_rel_err(C.cpu().numpy(), ref["centers"])
    print(f"lloyd n={n} d={d} k={k} iters={iters} {gen}: path={st['last_path']} n_it={n_it} center_rel_err={err:.3e} "
          f"recheck_rows={st['recheck_rows']}", flush=True)
    return err <= 1e-4


if __name__ == "__main__":
    ctx = _native.Context(0)
    ctx.set_option("kernel_path", 2)
    ok = True
    stage = sys.argv[1] if len(sys.argv) > 1 else "all"
    if stage in ("assign", "all"):
        for (n, d, k, gen) in [(128, 256, 256, "blobs"), (20000, 256, 256, "blobs"), (5000, 256, 256, "uniform"),
                               (3001, 256, 200, "blobs"), (4096, 192, 130, "uniform"), (1000, 256, 64, "blobs"),
                               (777, 132, 256, "uniform"), (64, 256, 256, "uniform"), (1, 256, 3, "uniform")]:
            ok &= run_assign(ctx, n, d, k, gen)
        # ties: all centres identical -> label 0 everywhere; a duplicated centre never wins
        Xt = ko.make_uniform(512, 256, seed=1)
        Ct = np.repeat(Xt[:1], 256, axis=0).copy()
        lab, _ = ctx.kmeans_assign(dev(Xt), dev(Ct))
        print("tie all-identical: max label", int(lab.max()), "nonzero", int((lab != 0).sum()), flush=True)
        ok &= int(lab.max()) == 0
        C2 = Xt[:256].copy(); C2[200] = C2[7]
        lab, _ = ctx.kmeans_assign(dev(Xt), dev(C2))
        print("tie duplicate: rows labelled 200:", int((lab == 200).sum()), flush=True)
        ok &= int((lab == 200).sum()) == 0
        X1 = ko.make_uniform(1, 256, seed=7); C1 = np.repeat(X1, 3, axis=0) + 0.01
        lab, _ = ctx.kmeans_assign(dev(X1), dev(C1))
        print("n=1 k=3 identical centres: label", lab.cpu().numpy(), flush=True)
        ok &= run_assign(ctx, 5000, 128, 64, "uniform", force=True)
        ok &= run_assign(ctx, 3000, 100, 40, "blobs", force=True)
    if stage in ("lloyd", "all"):
        ok &= run_lloyd(ctx, 20000, 256, 256, 4)
        ok &= run_lloyd(ctx, 6000, 256, 256, 3, gen="uniform")
        ok &= run_lloyd(ctx, 30000, 128, 64, 4, force=True)
        ok &= run_lloyd(ctx, 5000, 160, 200, 3)
    print("ALL OK" if ok else "FAILURES", flush=True)
