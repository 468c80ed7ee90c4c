"""On-disk format of estimator / model persistence (reference core.py:268-355): the directory layout
and metadata fields of pyspark's DefaultParamsWriter, so that what the reference wrote loads here and vice versa.
CPU only, no pyspark: the local writer / reader.  The pyspark branch (MLWriter + DefaultParamsWriter over a
SparkContext) and install.py's proxy are covered in test_pyspark_binding.py against the recording fake pyspark."""
import json
import os

import numpy as np
import pytest

from spark_rapids_ml_b200.clustering import KMeans, KMeansModel


def test_saved_directories_look_like_sparks(tmp_path):
    km = KMeans(k=5, maxIter=7, tol=1e-3, seed=3, num_workers=2).setFeaturesCol("f")
    p = str(tmp_path / "est")
    km.write().overwrite().save(p)
    assert sorted(os.listdir(p)) == ["metadata"]
    assert sorted(os.listdir(os.path.join(p, "metadata"))) == ["_SUCCESS", "part-00000"]
    lines = open(os.path.join(p, "metadata", "part-00000")).read().splitlines()
    assert len(lines) == 1                                   # one JSON object on one line, as saveAsTextFile writes it
    meta = json.loads(lines[0])
    for key in ("class", "timestamp", "sparkVersion", "uid", "paramMap", "defaultParamMap",
                "_cuml_params", "_num_workers", "_float32_inputs"):
        assert key in meta, key
    assert meta["class"].endswith("clustering.KMeans") and meta["uid"] == km.uid
    assert isinstance(meta["timestamp"], int) and int(meta["sparkVersion"].split(".")[0]) >= 3
    assert meta["paramMap"]["k"] == 5 and meta["paramMap"]["featuresCol"] == "f" and meta["_num_workers"] == 2
    assert meta["_cuml_params"]["n_clusters"] == 5 and meta["_cuml_params"]["max_iter"] == 7

    m = KMeansModel(cluster_centers_=[[0.0, 1.0], [2.0, 3.0]], n_cols=2, dtype="float32")
    mp = str(tmp_path / "model")
    m.write().overwrite().save(mp)
    assert sorted(os.listdir(mp)) == ["data", "metadata"]
    assert sorted(os.listdir(os.path.join(mp, "data"))) == ["_SUCCESS", "part-00000"]
    attrs = json.loads(open(os.path.join(mp, "data", "part-00000")).read())
    assert attrs == {"cluster_centers_": [[0.0, 1.0], [2.0, 3.0]], "n_cols": 2, "dtype": "float32"}
    m2 = KMeansModel.load(mp)
    assert m2.uid == m.uid and m2.cluster_centers_ == m.cluster_centers_
    assert all(p.parent == m2.uid for p in m2.params)        # the Params were re-parented (_resetUid)
    with pytest.raises(IOError):
        m.write().save(mp)                                   # exists, no overwrite()


def _write_as_the_reference_would(root, cls, uid, param_map, default_map, cuml_params, attrs=None):
    """What spark_rapids_ml's _CumlEstimatorWriter / _CumlModelWriter leave behind on a real cluster
    (DefaultParamsWriter.saveMetadata + sc.parallelize([json]).saveAsTextFile): part files, _SUCCESS and Hadoop's .crc."""
    meta = {"class": cls, "timestamp": 1718000000000, "sparkVersion": "3.5.1", "uid": uid, "paramMap": param_map,
            "defaultParamMap": default_map, "_cuml_params": cuml_params, "_num_workers": 4, "_float32_inputs": True}
    for sub, text in (("metadata", json.dumps(meta)),) + ((("data", json.dumps(attrs)),) if attrs is not None else ()):
        d = os.path.join(root, sub)
        os.makedirs(d)
        with open(os.path.join(d, "part-00000"), "w") as f:
            f.write(text + "\n")
        open(os.path.join(d, "_SUCCESS"), "w").close()
        open(os.path.join(d, ".part-00000.crc"), "wb").write(b"crc")
        open(os.path.join(d, "._SUCCESS.crc"), "wb").write(b"crc")


def test_directories_written_by_the_reference_load_here(tmp_path):
    defaults = {"k": 2, "maxIter": 20, "tol": 0.0001, "seed": 1, "initMode": "k-means||", "featuresCol": "features",
                "predictionCol": "prediction", "distanceMeasure": "euclidean", "initSteps": 2,
                "solver": "auto", "maxBlockSizeInMB": 0.0}     # the last two are Spark >= 3.4 params the GPU path ignores
    cuml = {"n_clusters": 3, "max_iter": 11, "tol": 0.0001, "verbose": False, "random_state": 1, "init": "scalable-k-means++",
            "n_init": 1, "oversampling_factor": 2.0, "max_samples_per_batch": 32768}
    ep = str(tmp_path / "ref_est")
    _write_as_the_reference_would(ep, "spark_rapids_ml.clustering.KMeans", "KMeans_4c2f0a1b9d3e",
                                  {"k": 3, "maxIter": 11, "featuresCol": "feats"}, defaults, cuml)
    est = KMeans.load(ep)
    assert est.uid == "KMeans_4c2f0a1b9d3e" and est.getK() == 3 and est.getMaxIter() == 11
    assert est.getFeaturesCol() == "feats" and est._num_workers == 4
    assert est.cuml_params["n_clusters"] == 3 and est.cuml_params["max_iter"] == 11

    centers = np.arange(6, dtype=np.float64).reshape(3, 2).tolist()
    mp = str(tmp_path / "ref_model")
    _write_as_the_reference_would(mp, "spark_rapids_ml.clustering.KMeansModel", "KMeans_4c2f0a1b9d3e",
                                  {"k": 3, "maxIter": 11, "featuresCol": "feats"}, defaults, cuml,
                                  attrs={"cluster_centers_": centers, "n_cols": 2, "dtype": "float64"})
    model = KMeansModel.load(mp)
    assert model.uid == "KMeans_4c2f0a1b9d3e" and model.n_cols == 2 and model.dtype == "float64"
    assert np.allclose(np.array(model.clusterCenters()), np.array(centers)) and model.getFeaturesCol() == "feats"
    # and back: what is saved here is what the reference's _CumlModelReader reads (same keys, same files)
    back = str(tmp_path / "back")
    model.write().overwrite().save(back)
    meta = json.loads(open(os.path.join(back, "metadata", "part-00000")).read())
    assert meta["uid"] == model.uid and meta["paramMap"]["featuresCol"] == "feats" and meta["_num_workers"] == 4
    assert json.loads(open(os.path.join(back, "data", "part-00000")).read())["cluster_centers_"] == centers


def test_save_is_write_save_and_refuses_an_existing_path(tmp_path):
    """pyspark.ml.util.MLWritable.save(path) == write().save(path): no silent overwrite (reference tests use
    estimator.save(path) / Model.load(path), tests/test_kmeans.py:505-512)."""
    km = KMeans(k=2)
    p = str(tmp_path / "kmeans")
    km.save(p)
    assert KMeans.load(p).getK() == 2
    with pytest.raises(IOError):
        km.save(p)
    km.write().overwrite().save(p)
    m = KMeansModel(cluster_centers_=[[0.5, 0.5], [8.5, 8.5]], n_cols=2, dtype="float32")
    mp = str(tmp_path / "kmeans_model")
    m.save(mp)
    m2 = KMeansModel.load(mp)
    assert m2.hasSummary is False and all(np.array_equal(a, b) for a, b in zip(m.clusterCenters(), m2.clusterCenters()))
    with pytest.raises(IOError):
        m.save(mp)
