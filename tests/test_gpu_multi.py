"""Two-GPU path (skipped on a 1-GPU box): two barrier-task processes, NCCL uid over the barrier allGather,
row-sharded fit with one fused allreduce per iteration — must equal the single-GPU fit of the same rows."""
import numpy as np
import pytest

from oracle import kmeans_oracle as ko

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_rank_fit_matches_single_rank_and_oracle():
    from spark_rapids_ml_b200.clustering import KMeans
    from spark_rapids_ml_b200.sparkshim import LocalSession

    s = LocalSession({"spark.sql.execution.arrow.maxRecordsPerBatch": "5000", "spark.rapids.ml.num_workers.local": "2"})
    X, _ = ko.make_blobs(40000, 128, 64, seed=4)
    df = s.from_numpy(X, num_partitions=2)
    m2 = KMeans(k=64, maxIter=8, tol=1e-6, seed=3, initMode="random", num_workers=2).fit(df)
    m1 = KMeans(k=64, maxIter=8, tol=1e-6, seed=3, initMode="random", num_workers=1).fit(df)
    C2, C1 = np.array(m2.cluster_centers_), np.array(m1.cluster_centers_)
    assert ko.max_center_rel_err(C2, C1) <= 1e-5
    lab = np.array([r["prediction"] for r in m2.transform(df).collect()])
    cmp = ko.compare_labels(X, C2.astype(np.float32), lab)
    assert cmp["n_mismatch_outside_margin"] == 0
