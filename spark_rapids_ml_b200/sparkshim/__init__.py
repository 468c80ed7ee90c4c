"""Minimal stand-ins for the pyspark pieces the KMeans path touches, used ONLY when pyspark is not importable
(this image has no pyspark/JVM — SURVEY.md §8c).  When pyspark is present the real classes are re-exported
instead, so `spark_rapids_ml_b200.clustering.KMeans` is a pyspark.ml Estimator there.

  params   : Param / Params / TypeConverters / keyword_only       (pyspark.ml.param, pyspark.keyword_only)
  sql      : LocalSession / LocalDataFrame / Row                    (SparkSession / DataFrame / Row subset:
             createDataFrame, repartition, select, first, count, collect, mapInPandas[barrier])
  barrier  : BarrierTaskContext (partitionId, allGather, barrier)   (pyspark.BarrierTaskContext)
"""
try:  # pragma: no cover - not reachable in this image
    import pyspark  # noqa: F401

    HAVE_PYSPARK = True
except Exception:
    HAVE_PYSPARK = False

from .barrier import BarrierTaskContext  # noqa: E402,F401
from .params import Param, Params, TypeConverters, keyword_only  # noqa: E402,F401
from .sql import LocalDataFrame, LocalSession, Row, get_session  # noqa: E402,F401
