"""Param plumbing of the KMeans surface — mirrors the reference's JVM-free unit tests
(python/tests/test_kmeans.py:66-199 test_params/test_kmeans_params/test_kmeans_copy, and the persistence and
alias-conflict checks of test_common_estimator.py:406-483).  CPU only."""
import numpy as np
import pytest

from spark_rapids_ml_b200.clustering import KMeans, KMeansModel
from spark_rapids_ml_b200.sparkshim import LocalSession


@pytest.fixture(autouse=True)
def session():
    LocalSession({"spark.rapids.ml.num_workers.local": "2"})
    yield


def test_default_cuml_params_and_spark_defaults():
    km = KMeans()
    # Spark-side defaults (pyspark.ml.clustering.KMeans): k=2, maxIter=20, tol=1e-4, initMode=k-means||
    assert km.getK() == 2 and km.getMaxIter() == 20 and km.getTol() == 1e-4 and km.getInitMode() == "k-means||"
    cp = km.cuml_params
    assert cp["n_clusters"] == 2 and cp["max_iter"] == 20 and cp["tol"] == 1e-4
    assert cp["init"] == "scalable-k-means++" and cp["n_init"] == 1
    assert cp["oversampling_factor"] == 2.0 and cp["max_samples_per_batch"] == 32768
    assert cp["random_state"] == km.getSeed() and 0 <= km.getSeed() <= 0x7FFFFFFF


def test_spark_and_backend_aliases():
    km = KMeans(k=5, maxIter=7, tol=0.5, seed=42, initMode="random")
    cp = km.cuml_params
    assert (cp["n_clusters"], cp["max_iter"], cp["tol"], cp["random_state"], cp["init"]) == (5, 7, 0.5, 42, "random")
    km2 = KMeans(n_clusters=6, max_iter=9, init="random", random_state=3, n_init=1)
    assert km2.getK() == 6 and km2.getMaxIter() == 9 and km2.getSeed() == 3
    with pytest.raises(ValueError, match="alias"):
        KMeans(k=2, n_clusters=3)
    with pytest.raises(ValueError, match="Unsupported param"):
        KMeans(not_a_param=1)
    with pytest.raises(ValueError):
        KMeans(initMode="bogus")


def test_tol_zero_maps_to_float32_tiny(caplog):
    km = KMeans()
    with caplog.at_level("WARNING"):
        km.setTol(0.0)
    assert km.cuml_params["tol"] == np.finfo("float32").tiny.item()
    assert "tol=0 is not supported in cuml yet" in caplog.text
    assert km.getTol() == 0.0


def test_setters_clear_copy_and_unsupported_params():
    km = KMeans().setK(4).setMaxIter(30).setSeed(7).setInitMode("random").setFeaturesCol("f")
    assert km.cuml_params["n_clusters"] == 4 and km.cuml_params["max_iter"] == 30
    km.clear(km.maxIter)
    assert km.getMaxIter() == 20 and km.cuml_params["max_iter"] == 20
    with pytest.raises(ValueError):
        km.setSeed(0x80000000)
    with pytest.raises(ValueError, match="weightCol"):
        km.setWeightCol("w")
    with pytest.raises(ValueError):
        KMeans(distanceMeasure="cosine")        # mapped to None: unsupported on GPU, no CPU fallback here
    KMeans(initSteps=5, solver="auto")          # mapped to "": accepted and ignored
    c = km.copy({km.getParam("k"): 9})
    assert c.getK() == 9 and c.cuml_params["n_clusters"] == 9 and km.getK() == 4
    km.setFeaturesCol(["a", "b"])
    assert km.getFeaturesCol() == ["a", "b"] and km.getFeaturesCols() == ["a", "b"]


def test_num_workers():
    assert KMeans().num_workers == 2            # inferred from the (shim) cluster
    assert KMeans(num_workers=1).num_workers == 1
    with pytest.raises(ValueError):
        _ = KMeans(num_workers=8).num_workers


def test_estimator_and_model_persistence(tmp_path):
    km = KMeans(k=3, maxIter=11, seed=5, num_workers=1, float32_inputs=False)
    p = str(tmp_path / "est")
    km.write().overwrite().save(p)
    km2 = KMeans.load(p)
    assert km2.cuml_params == km.cuml_params and km2.getK() == 3 and km2.num_workers == 1
    assert km2._float32_inputs is False and km2.uid == km.uid
    with pytest.raises(IOError):
        km.write().save(p)
    m = KMeansModel([[0.5, 0.5], [8.5, 8.5]], 2, "float32")
    m.setPredictionCol("newPrediction")
    mp = str(tmp_path / "model")
    m.write().overwrite().save(mp)
    m2 = KMeansModel.load(mp)
    assert m2.cluster_centers_ == m.cluster_centers_ and m2.n_cols == 2 and m2.dtype == "float32"
    assert m2.getPredictionCol() == "newPrediction" and m2.hasSummary is False
    assert [c.tolist() for c in m2.clusterCenters()] == [[0.5, 0.5], [8.5, 8.5]]


def test_merge_model_chunks_orders_by_chunk_id():
    from spark_rapids_ml_b200.sparkshim import Row

    rows = [Row(chunk_id=1, cluster_centers_=[[3.0]], n_cols=1, dtype="float32"),
            Row(chunk_id=0, cluster_centers_=[[1.0], [2.0]], n_cols=1, dtype="float32")]
    merged = KMeans()._merge_model_chunks(rows)
    assert merged[0]["cluster_centers_"] == [[1.0], [2.0], [3.0]]
    with pytest.raises(ValueError):
        KMeans()._merge_model_chunks([])


def test_constructor_defaults_from_session_confs():
    """reference core.py:1124-1170 / tests/test_kmeans.py:590-660: spark.rapids.ml.{verbose,float32_inputs,num_workers}
    fill in constructor arguments the user did not pass; explicit arguments win; bad values are loud."""
    import pytest
    from spark_rapids_ml_b200.sparkshim import LocalSession
    from spark_rapids_ml_b200.sparkshim.sql import LocalSession as LS

    saved = LS._active
    try:
        LocalSession(conf={"spark.rapids.ml.verbose": "5", "spark.rapids.ml.float32_inputs": "false",
                           "spark.rapids.ml.num_workers": "3"})
        est = KMeans()
        assert est._input_kwargs["verbose"] == 5 and est._input_kwargs["float32_inputs"] is False
        assert est._input_kwargs["num_workers"] == 3
        assert est._num_workers == 3 and est._float32_inputs is False and est.cuml_params["verbose"] == 5
        est = KMeans(verbose=False, float32_inputs=True, num_workers=1)          # explicit arguments win
        assert est._input_kwargs["verbose"] is False and est._float32_inputs is True and est._num_workers == 1
        LocalSession(conf={"spark.rapids.ml.verbose": "TRUE"})
        assert KMeans()._input_kwargs["verbose"] is True
        for key, bad in (("spark.rapids.ml.verbose", "7"), ("spark.rapids.ml.verbose", "loud"),
                         ("spark.rapids.ml.float32_inputs", "1"), ("spark.rapids.ml.num_workers", "0"),
                         ("spark.rapids.ml.num_workers", "two")):
            LocalSession(conf={key: bad})
            with pytest.raises(ValueError, match="Invalid value for " + key.replace(".", r"\.")):
                KMeans()
        LocalSession()
        assert "float32_inputs" not in KMeans()._input_kwargs                     # nothing set: nothing injected
    finally:
        LS._active = saved


def test_stage_level_scheduling_plan():
    """reference core.py:637-740 (_skip_stage_level_scheduling / _try_stage_level_scheduling) as a pure decision function."""
    from spark_rapids_ml_b200.spark_binding import stage_level_scheduling_plan as plan

    base = {"spark.master": "spark://h:7077", "spark.executor.cores": "12", "spark.executor.resource.gpu.amount": "1",
            "spark.task.resource.gpu.amount": "0.08"}

    def p(version="3.5.1", local=False, plugins="", sql="true", **over):
        conf = dict(base)
        for k, v in over.items():
            key = k.replace("__", ".")
            if v is None:
                conf.pop(key, None)
            else:
                conf[key] = v
        return plan(version, conf.get, local, plugins, sql)[0]

    assert p() == (7, 1.0)                                              # cores // 2 + 1: two tasks never share an executor
    assert p(plugins="com.nvidia.spark.SQLPlugin") == (12, 1.0)         # SQL plugin on: the whole executor
    assert p(plugins="com.nvidia.spark.SQLPlugin", sql="false") == (7, 1.0)
    assert p(spark__task__resource__gpu__amount=None) == (7, 1.0)       # ETL tasks take no GPU: training still must
    assert p(local=True) is None
    assert p(version="3.3.2") is None
    assert p(version="3.4.1", spark__master="yarn") is None             # 3.4.x: standalone / local-cluster only
    assert p(version="3.4.1") == (7, 1.0)
    assert p(version="3.5.1", spark__master="yarn") == (7, 1.0)
    assert p(version="3.10.0", spark__master="yarn") == (7, 1.0)        # numeric, not lexicographic, version order
    assert p(spark__executor__cores=None) is None and p(spark__executor__resource__gpu__amount=None) is None
    assert p(spark__executor__cores="1") is None
    assert p(spark__executor__resource__gpu__amount="2") is None
    assert p(spark__task__resource__gpu__amount="1") is None            # already one task per GPU
