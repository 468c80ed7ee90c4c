"""End-to-end through the reference-facing surface on the GPU: KMeans(...).fit(df) / KMeansModel.transform(df)
over Arrow batches — the reference's toy integration tests (python/tests/test_kmeans.py:202-282, 420-526)."""
import json
import os

import numpy as np
import pytest

from oracle import kmeans_oracle as ko

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture()
def session():
    from spark_rapids_ml_b200.sparkshim import LocalSession

    return LocalSession({"spark.sql.execution.arrow.maxRecordsPerBatch": "3", "spark.rapids.ml.num_workers.local": "1"})


@pytest.mark.parametrize("arrow_backed", [True, False])
def test_reference_toy_cases_fit_transform_persist(session, tmp_path, arrow_backed):
    from spark_rapids_ml_b200.clustering import KMeans, KMeansModel

    for case in json.load(open(os.path.join(GOLD, "kmeans_known_answers.json")))["cases"]:
        df = session.createDataFrame([(r,) for r in case["data"]], ["features"])
        df.arrow_backed_pandas = arrow_backed
        for init_mode in ("k-means||", "random"):
            hits = 0
            for seed in range(4):
                km = KMeans(k=case["k"], maxIter=case["max_iter"], tol=case["tol"], seed=seed, initMode=init_mode,
                            num_workers=1).setFeaturesCol("features")
                model = km.fit(df)
                assert model.dtype == "float32" and model.n_cols == 2 and model.hasSummary is False
                C = sorted(c.tolist() for c in model.clusterCenters())
                if np.allclose(C, case["expected_sorted_centers"], rtol=max(case["rel_tol"], 1e-7)):
                    hits += 1
                    out = model.setPredictionCol("newPrediction").transform(df)
                    assert sorted(out.columns) == ["features", "newPrediction"]
                    lab = [r["newPrediction"] for r in out.collect()]
                    for a, b in case["same_label_pairs"]:
                        assert lab[a] == lab[b]
                    for a, b in case["diff_label_pairs"]:
                        assert lab[a] != lab[b]
                    p = str(tmp_path / f"m_{case['name']}_{init_mode.strip('|')}_{seed}_{arrow_backed}")
                    model.write().overwrite().save(p)
                    again = KMeansModel.load(p)
                    assert again.cluster_centers_ == model.cluster_centers_
                    assert model.predict(case["data"][0]) == lab[0]
            assert hits >= 2, (case["name"], init_mode)


def test_fit_layouts_agree_with_oracle(session):
    """array<float>, array<double> (cast by float32_inputs) and multi-column layouts; many small Arrow batches."""
    import pandas as pd

    from spark_rapids_ml_b200.clustering import KMeans

    session.conf.set("spark.sql.execution.arrow.maxRecordsPerBatch", "257")
    X, _ = ko.make_blobs(3000, 12, 4, seed=5)
    ref = None
    for layout in ("array_f32", "array_f64", "multi_cols"):
        if layout == "array_f32":
            df = session.from_numpy(X)
            km = KMeans(k=4, maxIter=30, seed=1, initMode="k-means||", num_workers=1)
        elif layout == "array_f64":
            df = session.from_numpy(X.astype(np.float64))
            km = KMeans(k=4, maxIter=30, seed=1, initMode="k-means||", num_workers=1)
        else:
            cols = [f"c{i}" for i in range(12)]
            df = session.createDataFrame(pd.DataFrame(X.astype(np.float64), columns=cols))
            km = KMeans(k=4, maxIter=30, seed=1, initMode="k-means||", num_workers=1).setFeaturesCols(cols)
        model = km.fit(df)
        C = np.array(sorted(c.tolist() for c in model.clusterCenters()))
        if ref is None:
            ref = C
        np.testing.assert_allclose(C, ref, rtol=1e-5, atol=1e-6)
        lab = np.array([r["prediction"] for r in model.transform(df).collect()])
        cmp = ko.compare_labels(X, np.array(model.cluster_centers_, dtype=np.float32), lab)
        assert cmp["n_mismatch_outside_margin"] == 0


def test_empty_partition_raises(session):
    from spark_rapids_ml_b200.clustering import KMeans

    df = session.from_numpy(np.zeros((3, 4), dtype=np.float32), num_partitions=1)
    df._parts = [[]]                      # a partition whose Arrow stream is empty (core.py:959-962)
    with pytest.raises(RuntimeError, match="no data"):
        KMeans(k=2, num_workers=1).fit(df)


@pytest.mark.parametrize("data_type", ["byte", "short", "int", "long"])
def test_integer_feature_columns(session, data_type):
    """reference tests/test_kmeans.py:312-335 (test_kmeans_numeric_type): scalar integer columns are accepted (cast to
    float32 on the way in) — here also checked against the oracle on the same five rows."""
    from spark_rapids_ml_b200.clustering import KMeans

    data = [[1, 4, 4, 4, 0], [2, 2, 2, 2, 1], [3, 3, 3, 2, 2], [3, 3, 3, 2, 3], [5, 2, 1, 3, 4]]
    cols = ["c1", "c2", "c3", "c4", "c5"]
    df = session.createDataFrame(data, schema=", ".join(f"{c} {data_type}" for c in cols))
    model = KMeans(num_workers=1, featuresCols=cols, n_clusters=2, initMode="random", seed=1, maxIter=10).fit(df)
    assert model.n_cols == 5 and model.dtype == "float32"
    X = np.asarray(data, dtype=np.float32)
    C = np.asarray(model.cluster_centers_, dtype=np.float64)
    lab, _, _ = ko.assign(X, C.astype(np.float32))
    for j in range(2):                              # a fixed point of Lloyd on these rows: centres = means of their rows
        assert (lab == j).any() and np.allclose(C[j], X[lab == j].mean(0), rtol=1e-5)
    out = model.transform(df)
    assert [r["prediction"] for r in out.collect()] == lab.tolist()


def test_parameters_validation(session):
    """reference tests/test_kmeans.py:529-546: k = -1 / maxIter = -1 are refused with Spark's wording."""
    from spark_rapids_ml_b200.clustering import KMeans

    df = session.createDataFrame([([1.0, 2.0], 1.0), ([3.0, 1.0], 0.0)], schema="features array<float>, label float")
    with pytest.raises(ValueError, match="k given invalid value -1"):
        KMeans(k=-1, num_workers=1).fit(df)
    with pytest.raises(ValueError, match="maxIter given invalid value -1"):
        KMeans(num_workers=1).setMaxIter(-1).fit(df)


def test_session_confs_reach_the_fit(tmp_path):
    """spark.rapids.ml.float32_inputs = false from the session conf reaches the estimator and its model like the constructor
    argument (reference core.py:1149-1159): double columns are then NOT cast on the DataFrame side and are converted by
    the ingest kernel instead.  The arithmetic of this build is fp32 either way, and the model says so."""
    from spark_rapids_ml_b200.clustering import KMeans
    from spark_rapids_ml_b200.sparkshim import LocalSession

    sess = LocalSession({"spark.rapids.ml.float32_inputs": "false"})
    try:
        X, _ = ko.make_blobs(400, 8, 3, seed=2)
        df = sess.createDataFrame([(r.astype(np.float64).tolist(),) for r in X], schema="features array<double>")
        est = KMeans(k=3, num_workers=1, seed=1, initMode="random")
        assert est._float32_inputs is False
        model = est.fit(df)
        assert model._float32_inputs is False and model.dtype == "float32"
        C = np.asarray(sorted(model.cluster_centers_))
        ref = ko.lloyd([X], np.asarray(model.cluster_centers_, dtype=np.float32), 1, -1.0)["centers"]
        assert ko.max_center_rel_err(np.asarray(sorted(ref.tolist())), C) <= 1e-3      # a converged, sane model
    finally:
        LocalSession()


def test_fit_multiple_and_num_features(session):
    """pyspark.ml.Estimator.fitMultiple over param maps (one fit per map, as the reference does for KMeans) and
    KMeansModel.numFeatures (reference core.py:1961-1967)."""
    from spark_rapids_ml_b200.clustering import KMeans

    X, _ = ko.make_blobs(300, 6, 3, seed=4)
    df = session.createDataFrame([(r.tolist(),) for r in X], schema="features array<float>")
    est = KMeans(k=2, num_workers=1, seed=1, initMode="random", maxIter=5)
    maps = [{est.k: 2}, {est.k: 3, est.maxIter: 7}]
    got = dict(est.fitMultiple(df, maps))
    assert sorted(got) == [0, 1]
    assert len(got[0].cluster_centers_) == 2 and len(got[1].cluster_centers_) == 3
    assert got[1].getMaxIter() == 7 and got[0].getMaxIter() == 5 and est.getK() == 2      # the estimator itself is untouched
    assert got[0].numFeatures == 6 and got[1].numFeatures == 6
