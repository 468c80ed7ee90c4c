"""CPU oracle for the distributed KMeans.fit() Lloyd loop — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (spark_rapids_ml_b200) never does.

What it restates
----------------
The reference (NVIDIA/spark-rapids-ml @ c51743bb) has no arithmetic for this path: it calls
RAPIDS cuML 25.12 ``cuml.cluster.kmeans_mg.KMeansMG`` (call site
``python/src/spark_rapids_ml/clustering.py:381-415``; predict ``clustering.py:584-601``).
cuML's source is NOT under /root/reference and cuML is not installable here, so this file
restates cuML-MG's *published* Lloyd semantics in float64 NumPy:

  * inputs f32 [n,d] row-major, centers f32 [k,d], labels int32
  * assign   l_i = argmin_j ||x_i - c_j||^2, evaluated in fp64 from the f32 values,
             lowest j on ties                              (cuML fusedL2NN argmin)
  * update   c_j = (sum_{l_i=j} x_i) / w_j in fp64, rounded to f32 once;
             w_j == 0 keeps the previous c_j               (cuML keeps old centroid)
  * ranks    all-rank reduction = plain sum (order independent in fp64)
  * stop     sum_j ||c_j_new - c_j_old||^2 < tol, or max_iter; tol == 0 is mapped to
             float32 tiny by the reference (clustering.py:113-123)
  * labels   "labels" = a final assign of X against the FINAL centers (what the reference
             exposes through transform -> predict, clustering.py:598-602)
  * inertia  sum_i ||x_i - c_{l_i}||^2 in fp64 against the final centers
             (python/benchmark/benchmark/bench_kmeans.py:61-112 definition)

Pinning status
--------------
Pinned against every known-answer the reference's own tests hold for this path
(tests/golden/kmeans_known_answers.json, from python/tests/test_kmeans.py:202-249 and
:420-526 and jvm/.../SparkRapidsMLSuite.scala:338-381) and cross-checked against
scikit-learn's Lloyd (same init, n_init=1) in tests/test_oracle.py.  Beyond those toy
known-answers the reference stores NO golden vectors (its seeded test is xfail,
test_kmeans.py:332,355), so seeded-init results are "parity unpinned" by the reference
itself; all parity tests therefore inject identical initial centers.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

F32_TINY = float(np.finfo("float32").tiny)


def map_tol(tol: float) -> float:
    """tol == 0 -> float32 tiny (reference: clustering.py:113-123)."""
    return F32_TINY if tol == 0.0 else float(tol)


def pairwise_sqdist(X: np.ndarray, C: np.ndarray) -> np.ndarray:
    """fp64 ||x-c||^2 via the expanded form on fp64 copies of the f32 inputs.

    Chunked so an [n,k] fp64 matrix is only materialised per block.
    """
    X64 = np.asarray(X, dtype=np.float64)
    C64 = np.asarray(C, dtype=np.float64)
    xn = np.einsum("ij,ij->i", X64, X64)
    cn = np.einsum("ij,ij->i", C64, C64)
    D = xn[:, None] + cn[None, :] - 2.0 * (X64 @ C64.T)
    np.maximum(D, 0.0, out=D)
    return D


def assign(
    X: np.ndarray, C: np.ndarray, block: int = 65536
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Return (labels int32 [n], min sq. distance fp64 [n], margin fp64 [n]).

    margin_i = (d2 - d1) / max(d1, ||x_i||^2): the relative gap between the best and the
    second-best center — the quantity SURVEY.md §8c's parity rule thresholds at 1e-6.
    For k == 1 the margin is +inf.
    """
    n = X.shape[0]
    k = C.shape[0]
    labels = np.empty(n, dtype=np.int32)
    mind = np.empty(n, dtype=np.float64)
    margin = np.empty(n, dtype=np.float64)
    for s in range(0, n, block):
        e = min(n, s + block)
        Xb = np.asarray(X[s:e], dtype=np.float64)
        D = pairwise_sqdist(Xb, C)
        lab = np.argmin(D, axis=1)  # first occurrence == lowest j on ties
        d1 = D[np.arange(e - s), lab]
        labels[s:e] = lab
        mind[s:e] = d1
        if k > 1:
            D[np.arange(e - s), lab] = np.inf
            d2 = D.min(axis=1)
            xn = np.einsum("ij,ij->i", Xb, Xb)
            denom = np.maximum(np.maximum(d1, xn), np.finfo(np.float64).tiny)
            with np.errstate(over="ignore"):
                margin[s:e] = (d2 - d1) / denom
        else:
            margin[s:e] = np.inf
    return labels, mind, margin


def partial_sums(
    X: np.ndarray, labels: np.ndarray, k: int
) -> Tuple[np.ndarray, np.ndarray]:
    """Per-cluster fp64 sums [k,d] and counts [k] (one rank's contribution)."""
    d = X.shape[1]
    S = np.zeros((k, d), dtype=np.float64)
    X64 = np.asarray(X, dtype=np.float64)
    order = np.argsort(labels, kind="stable")
    sl = labels[order]
    bounds = np.searchsorted(sl, np.arange(k + 1))
    w = np.diff(bounds).astype(np.float64)
    for j in range(k):
        if bounds[j + 1] > bounds[j]:
            S[j] = X64[order[bounds[j] : bounds[j + 1]]].sum(axis=0)
    return S, w


def lloyd(
    parts: Sequence[np.ndarray],
    C0: np.ndarray,
    max_iter: int,
    tol: float,
) -> Dict[str, object]:
    """cuML-MG Lloyd loop over `parts` (one f32 [n_r,d] array per rank).

    Returns dict(centers f32 [k,d], n_iter, inertia fp64, labels [per-part int32],
    shifts [per-iteration sum_j||dc_j||^2]).
    """
    tol = map_tol(tol)
    C = np.array(C0, dtype=np.float32, order="C")
    k, d = C.shape
    n_iter = 0
    shifts: List[float] = []
    for it in range(1, max_iter + 1):
        S = np.zeros((k, d), dtype=np.float64)
        w = np.zeros(k, dtype=np.float64)
        for Xp in parts:  # the allreduce: plain sum over ranks
            lab, _, _ = assign(Xp, C)
            Sp, wp = partial_sums(Xp, lab, k)
            S += Sp
            w += wp
        Cn = C.astype(np.float64)
        nz = w > 0
        Cn[nz] = S[nz] / w[nz][:, None]
        Cn32 = Cn.astype(np.float32)
        diff = Cn32.astype(np.float64) - C.astype(np.float64)
        shift = float(np.sum(diff * diff))
        shifts.append(shift)
        C = Cn32
        n_iter = it
        if shift < tol:
            break
    labels = []
    inertia = 0.0
    for Xp in parts:
        lab, md, _ = assign(Xp, C)
        labels.append(lab)
        inertia += float(md.sum())
    return {
        "centers": C,
        "n_iter": n_iter,
        "inertia": inertia,
        "labels": labels,
        "shifts": shifts,
    }


def lloyd_iteration(
    parts: Sequence[np.ndarray], C: np.ndarray
) -> Tuple[np.ndarray, np.ndarray, float]:
    """One Lloyd iteration: returns (new centers f32, counts fp64, shift)."""
    C = np.asarray(C, dtype=np.float32)
    k, d = C.shape
    S = np.zeros((k, d), dtype=np.float64)
    w = np.zeros(k, dtype=np.float64)
    for Xp in parts:
        lab, _, _ = assign(Xp, C)
        Sp, wp = partial_sums(Xp, lab, k)
        S += Sp
        w += wp
    Cn = C.astype(np.float64)
    nz = w > 0
    Cn[nz] = S[nz] / w[nz][:, None]
    Cn32 = Cn.astype(np.float32)
    diff = Cn32.astype(np.float64) - C.astype(np.float64)
    return Cn32, w, float(np.sum(diff * diff))


def inertia(X: np.ndarray, C: np.ndarray) -> float:
    """bench_kmeans.py:61-112 score(): fp64 sum of squared distances to nearest center."""
    _, md, _ = assign(X, C)
    return float(md.sum())


# --------------------------------------------------------------------------------------
# Initialisers (distributional parity only — cuML's RNG stream is unpinnable; see header)
# --------------------------------------------------------------------------------------
def init_random(parts: Sequence[np.ndarray], k: int, seed: int) -> np.ndarray:
    """k distinct rows drawn uniformly from the union of all parts."""
    n_tot = sum(p.shape[0] for p in parts)
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(n_tot, size=k, replace=False))
    X = np.concatenate(parts, axis=0)
    return np.array(X[idx], dtype=np.float32)


def kmeans_plus_plus_weighted(
    P: np.ndarray, wts: np.ndarray, k: int, rng: np.random.Generator
) -> np.ndarray:
    """Weighted greedy k-means++ on a small candidate set (final step of k-means||).

    Greedy = 2+log(k) D^2-sampled trials per step, keep the one with the lowest potential
    (Arthur & Vassilvitskii's suggested variant; also what scikit-learn's k-means++ does).
    """
    m = P.shape[0]
    P64 = P.astype(np.float64)
    trials = 2 + int(np.log(max(k, 2)))
    centers = np.empty((k, P.shape[1]), dtype=np.float64)
    first = rng.choice(m, p=wts / wts.sum())
    centers[0] = P64[first]
    d2 = ((P64 - centers[0]) ** 2).sum(axis=1)
    for j in range(1, k):
        prob = wts * d2
        tot = prob.sum()
        if tot <= 0:
            cands = rng.integers(m, size=1)
        else:
            cands = rng.choice(m, size=trials, p=prob / tot)
        best_pot, best_c, best_d2 = None, None, None
        for c in cands:
            nd2 = np.minimum(d2, ((P64 - P64[c]) ** 2).sum(axis=1))
            pot = float((wts * nd2).sum())
            if best_pot is None or pot < best_pot:
                best_pot, best_c, best_d2 = pot, c, nd2
        centers[j] = P64[best_c]
        d2 = best_d2
    return centers.astype(np.float32)


def init_kmeans_parallel(
    parts: Sequence[np.ndarray],
    k: int,
    seed: int,
    oversampling: float = 2.0,
    rounds: int = 5,
) -> np.ndarray:
    """Scalable k-means++ (k-means||, Bahmani et al.) with oversampling l = oversampling*k.

    The reference forwards init="scalable-k-means++", oversampling_factor=2.0
    (clustering.py:134-136) to cuML.
    """
    rng = np.random.default_rng(seed)
    X = np.concatenate(parts, axis=0)
    n = X.shape[0]
    first = int(rng.integers(n))
    cand = [np.array(X[first], dtype=np.float32)]
    _, d2, _ = assign(X, np.stack(cand))
    ell = oversampling * k
    for _ in range(rounds):
        phi = d2.sum()
        if phi <= 0:
            break
        prob = np.minimum(1.0, ell * d2 / phi)
        pick = np.nonzero(rng.random(n) < prob)[0]
        if pick.size == 0:
            continue
        new = np.array(X[pick], dtype=np.float32)
        cand.extend(list(new))
        _, dn, _ = assign(X, new)
        d2 = np.minimum(d2, dn)
    P = np.stack(cand)
    if P.shape[0] <= k:
        # top up with random rows
        extra = rng.choice(n, size=k - P.shape[0] + 1, replace=False)
        P = np.concatenate([P, X[extra].astype(np.float32)], axis=0)
    lab, _, _ = assign(X, P)
    wts = np.bincount(lab, minlength=P.shape[0]).astype(np.float64)
    wts = np.maximum(wts, 1e-12)
    C = kmeans_plus_plus_weighted(P, wts, k, rng)
    # cuML refines the reduced problem with a few weighted Lloyd steps on the candidates
    for _ in range(10):
        D = pairwise_sqdist(P, C)
        l2 = np.argmin(D, axis=1)
        Cn = C.astype(np.float64)
        for j in range(k):
            m = l2 == j
            if m.any():
                Cn[j] = (P[m].astype(np.float64) * wts[m][:, None]).sum(0) / wts[m].sum()
        C = Cn.astype(np.float32)
    return C


# --------------------------------------------------------------------------------------
# Parity rule — shared by the GPU parity tests and smoke()
# --------------------------------------------------------------------------------------
def compare_labels(
    X: np.ndarray, C: np.ndarray, gpu_labels: np.ndarray, tau: float = 1e-6
) -> Dict[str, int]:
    """Compare device labels with the fp64 oracle on the SAME centers.

    A mismatch is admissible only when the oracle's fp64 relative margin for that row is
    below tau (fp32 ordering noise).  n_mismatch_outside_margin must be 0.
    """
    lab, _, margin = assign(X, C)
    mism = np.nonzero(lab != np.asarray(gpu_labels))[0]
    outside = int(np.count_nonzero(margin[mism] >= tau))
    return {
        "n": int(X.shape[0]),
        "n_mismatch": int(mism.size),
        "n_mismatch_outside_margin": outside,
    }


def max_center_rel_err(Cg: np.ndarray, Co: np.ndarray) -> float:
    """max_j ||c_gpu - c_oracle|| / ||c_oracle|| (north_star: <= 1e-4)."""
    Cg = np.asarray(Cg, dtype=np.float64)
    Co = np.asarray(Co, dtype=np.float64)
    num = np.linalg.norm(Cg - Co, axis=1)
    den = np.maximum(np.linalg.norm(Co, axis=1), np.finfo(np.float64).tiny)
    return float((num / den).max())


# --------------------------------------------------------------------------------------
# Synthetic data (mirrors the reference generators' distributions:
# python/benchmark/gen_data_distributed.py:84-187 blobs, gen_data.py:243-261 uniform)
# --------------------------------------------------------------------------------------
def make_blobs(
    n: int, d: int, k_true: int, seed: int, center_seed: int = 42, std: float = 1.0
) -> Tuple[np.ndarray, np.ndarray]:
    crng = np.random.default_rng(center_seed)
    centers = crng.uniform(-10.0, 10.0, size=(k_true, d)).astype(np.float32)
    rng = np.random.default_rng(seed)
    z = rng.integers(0, k_true, size=n)
    X = centers[z] + rng.normal(0.0, std, size=(n, d)).astype(np.float32)
    return np.ascontiguousarray(X, dtype=np.float32), centers


def make_uniform(n: int, d: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.random((n, d), dtype=np.float32)
