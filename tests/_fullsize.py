"""Every-row parity at BASELINE's full per-GPU sizes (the oracle cannot finish these in seconds: the checker is a plain
PyTorch fp64 restatement of the same argmin on the GPU, chunked) + the size-independent Lloyd properties."""
import torch

TAU = 1e-6


def make_blobs(n, d, k, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ctr = torch.rand((k, d), generator=g, device="cuda") * 20 - 10
    X = torch.empty((n, d), device="cuda")
    for s in range(0, n, 1_000_000):
        e = min(n, s + 1_000_000)
        X[s:e] = ctr[torch.randint(0, k, (e - s,), generator=g, device="cuda")] + torch.randn((e - s, d), generator=g, device="cuda")
    C = (ctr + 0.5 * torch.randn((k, d), generator=g, device="cuda")).contiguous()
    return X, C


def check_every_row(ctx, X, C, chunk=500_000):
    """labels / min distances of b2k_kmeans_assign against fp64 torch on EVERY row; a label may differ only where the
    fp64 margin (d2 - d1) / max(d1, ||x||^2) is below TAU (the parity rule of tests/test_gpu_parity.py)."""
    n, d = X.shape
    labels, md = ctx.kmeans_assign(X, C, want_mindist=True)
    C64 = C.double()
    cn = (C64 * C64).sum(1)
    bad = 0
    differ = 0
    worst_md = 0.0
    for s in range(0, n, chunk):
        x = X[s:s + chunk].double()
        xn = (x * x).sum(1)
        D = xn[:, None] + cn[None, :] - 2.0 * (x @ C64.T)
        two = torch.topk(D, 2, dim=1, largest=False)
        ref = two.indices[:, 0]
        got = labels[s:s + chunk].long()
        ne = got != ref
        differ += int(ne.sum())
        if bool(ne.any()):
            margin = (two.values[:, 1] - two.values[:, 0]) / torch.maximum(two.values[:, 0], xn).clamp(min=1e-300)
            # the device's choice must itself be within the margin of the best
            dgot = D.gather(1, got[:, None])[:, 0]
            off = (dgot - two.values[:, 0]) / torch.maximum(two.values[:, 0], xn).clamp(min=1e-300)
            bad += int((ne & ((margin >= TAU) | (off >= TAU))).sum())
        rel = ((md[s:s + chunk].double() - two.values[:, 0]).abs() / torch.maximum(two.values[:, 0], 1e-6 * xn).clamp(min=1e-30))
        worst_md = max(worst_md, float(rel[~ne].max()) if bool((~ne).any()) else 0.0)
    return {"n": n, "labels_differ": differ, "outside_margin": bad, "worst_mindist_rel_err": worst_md, "labels": labels}


def check_one_step(ctx, X, C, labels):
    """one Lloyd step == the fp64 sums implied by the device labels; bitwise determinism."""
    k, d = C.shape
    C1 = C.clone()
    n_it, _ = ctx.kmeans_lloyd(X, C1, 1, 0.0)
    assert n_it == 1
    S = torch.zeros((k, d), dtype=torch.float64, device="cuda")
    for s in range(0, X.shape[0], 1_000_000):
        S.index_add_(0, labels[s:s + 1_000_000].long(), X[s:s + 1_000_000].double())
    w = torch.bincount(labels.long(), minlength=k).double()
    assert int(w.sum()) == X.shape[0]
    exp = torch.where(w[:, None] > 0, S / w.clamp(min=1)[:, None], C.double()).float()
    rel = ((C1 - exp).double().norm(dim=1) / exp.double().norm(dim=1)).max().item()
    C2 = C.clone()
    ctx.kmeans_lloyd(X, C2, 1, 0.0)
    return rel, bool(torch.equal(C1, C2))
