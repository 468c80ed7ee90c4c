"""The pyspark branch of the Spark<->worker boundary (SURVEY.md 8b "B1"), checked without a pyspark / JVM: a fresh
interpreter imports the library with tests/fake_pyspark on sys.path (HAVE_PYSPARK = True) and runs fit / transform on a
recording fake pyspark DataFrame.  What must hold (reference call sites in parentheses):

  * KMeans IS a pyspark.ml.Estimator with pyspark Params, KMeansModel a pyspark.ml.Model
  * fit: select/cast of the feature column, VectorUDT -> vector_to_array (core.py:523-525), repartition(num_workers),
    mapInPandas(_train_udf, schema).rdd.barrier().mapPartitions(...).collect() (core.py:1005-1013), the worker takes its
    context from pyspark.BarrierTaskContext.get(), local mode derived from the session's master URL
  * transform: pandas_udf("int") over struct(features) appended with withColumn (core.py:1846-1878)

CPU version: the device pieces of the worker (GPU selection, CumlContext, the device row appender, the fit function)
are replaced by host stand-ins — the WIRING is what is under test.  GPU version: nothing is replaced."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAKE = os.path.join(ROOT, "tests", "fake_pyspark")

_COMMON = '''
import sys, json
import numpy as np
import pyspark
from pyspark import CALLS
from pyspark.sql import DataFrame
import spark_rapids_ml_b200.sparkshim as shim
assert shim.HAVE_PYSPARK
from spark_rapids_ml_b200.clustering import KMeans, KMeansModel
import pyspark.ml, pyspark.ml.param
assert issubclass(KMeans, pyspark.ml.Estimator) and issubclass(KMeansModel, pyspark.ml.Model)
assert isinstance(KMeans().getParam("k"), pyspark.ml.param.Param)
sess = shim.LocalSession()
rng = np.random.default_rng(0)
ctr = rng.uniform(-10, 10, size=(4, 8))
X = (ctr[rng.integers(0, 4, size=600)] + 0.05 * rng.normal(size=(600, 8)))
'''

_CPU_STUBS = '''
import pandas as pd
import spark_rapids_ml_b200.core as core
import spark_rapids_ml_b200.common.cuml_context as cc

class HostAppender:
    def __init__(self, ctx, d, first_capacity=0):
        self.d, self.rows_ = d, []
    def append_values(self, values, offsets, n_b):
        lo = int(offsets[0]) if offsets is not None else 0
        self.rows_.append(np.asarray(values[lo:lo + n_b * self.d], dtype=np.float32).reshape(n_b, self.d))
    def append_columns(self, cols):
        self.rows_.append(np.stack(cols, 1).astype(np.float32))
    def finish(self):
        return np.concatenate(self.rows_)

class HostContext:
    def __init__(self, *a, **k): self.handle, self._loop = object(), None
    def __enter__(self): return self
    def __exit__(self, *a): return None

core.DeviceRowAppender = HostAppender
cc.CumlContext = HostContext
core._CumlCommon._set_gpu_device = staticmethod(lambda context, is_local, is_transform=False: 0)

def host_fit_func(self, dataset, extra_params=None):
    def fit(inputs, params):
        Xh = inputs[0][0]
        k = params[core.param_alias.cuml_init]["n_clusters"]
        C = Xh[:k].astype(np.float64)
        for _ in range(5):
            lab = ((Xh[:, None, :] - C[None]) ** 2).sum(-1).argmin(1)
            C = np.stack([Xh[lab == j].mean(0) if (lab == j).any() else C[j] for j in range(k)])
        return {"chunk_id": [0], "cluster_centers_": [C.tolist()], "n_cols": [Xh.shape[1]], "dtype": ["float32"]}
    return fit
KMeans._get_cuml_fit_func = host_fit_func

def host_transform_func(self, dataset, eval_metric_info=None):
    C = np.asarray(self.cluster_centers_)
    def construct(gpu=0): return C
    def transform(model, df):
        col = df[core.alias.data] if hasattr(df, "columns") and core.alias.data in df.columns else df
        A = np.array([np.asarray(r, dtype=np.float64) for r in col])
        return pd.Series(((A[:, None, :] - model[None]) ** 2).sum(-1).argmin(1).astype("int32"))
    return construct, transform, None
KMeansModel._get_cuml_transform_func = host_transform_func
'''

_BODY = '''
results = {}
for kind in ("array<float>", "array<double>", "vector"):
    del CALLS[:]
    Xs = X.astype(np.float32) if kind == "array<float>" else X
    local = sess.from_numpy(Xs, col="features", num_partitions=2)
    df = DataFrame(local, vector_cols=("features",) if kind == "vector" else ())
    model = KMeans(k=4, maxIter=5, initMode="random", seed=1, num_workers=1).setFeaturesCol("features").fit(df)
    fit_calls = [c[0] for c in CALLS]
    C = np.array(model.cluster_centers_)
    del CALLS[:]
    out = model.transform(df)
    tr_calls = [c[0] for c in CALLS]
    pred = np.array([r["prediction"] for r in out.collect()])
    truth = ((X[:, None, :] - C[None]) ** 2).sum(-1).argmin(1)
    results[kind] = {"fit_calls": fit_calls, "transform_calls": tr_calls, "pred_ok": bool((pred == truth).all()),
                     "n_centers": int(C.shape[0]), "pred_col": "prediction" in out.columns}
print("RESULT " + json.dumps(results))
'''


def _run(script: str) -> dict:
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([FAKE, ROOT, env.get("PYTHONPATH", "")])
    res = subprocess.run([sys.executable, "-c", textwrap.dedent(script)], env=env, capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def _check(results: dict) -> None:
    for kind, r in results.items():
        fc = r["fit_calls"]
        # the reference's fit plan, in order
        i_map, i_bar, i_mp, i_col = (fc.index(x) for x in ("mapInPandas", "rdd.barrier", "rdd.mapPartitions", "rdd.collect"))
        assert fc.index("select") < fc.index("repartition") < i_map < i_bar < i_mp < i_col, (kind, fc)
        assert "BarrierTaskContext.get" in fc, (kind, fc)          # the worker asked pyspark for its barrier context
        assert ("vector_to_array" in fc) == (kind == "vector"), (kind, fc)
        tc = r["transform_calls"]
        assert tc.index("pandas_udf") < tc.index("withColumn"), (kind, tc)
        assert ("vector_to_array" in tc) == (kind == "vector"), (kind, tc)
        assert r["pred_ok"] and r["pred_col"] and r["n_centers"] == 4, (kind, r)


def test_pyspark_branch_wiring_with_host_stand_ins():
    _check(_run(_COMMON + _CPU_STUBS + _BODY))


@pytest.mark.gpu
def test_pyspark_branch_end_to_end_on_gpu():
    """Same plan, nothing replaced: the worker function ingests the Arrow batches on the device and calls libb2kmeans."""
    _check(_run(_COMMON + _BODY))
