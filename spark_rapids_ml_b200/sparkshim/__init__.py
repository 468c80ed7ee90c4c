"""Stand-ins for the pyspark pieces the KMeans path touches, and the switch between them and real pyspark.

When `import pyspark` succeeds, Param / Params / TypeConverters / keyword_only / Row are the REAL pyspark classes (so
`spark_rapids_ml_b200.clustering.KMeans` is a `pyspark.ml.Estimator` with pyspark Params, and the fitted model a
`pyspark.ml.Model`), and a pyspark DataFrame handed to fit()/transform() is driven through
`spark_rapids_ml_b200/spark_binding.py` (mapInPandas + barrier RDD, pandas_udf).  This image has no pyspark / JVM
(SURVEY.md 8c); there the minimal local versions below are used, and tests/test_pyspark_binding.py checks the pyspark
branch against a recording fake `pyspark` package.

  params   : Param / Params / TypeConverters / keyword_only       (pyspark.ml.param, pyspark.keyword_only)
  sql      : LocalSession / LocalDataFrame / Row                    (SparkSession / DataFrame / Row subset:
             createDataFrame, repartition, select, first, count, collect, mapInPandas[barrier]) — always available:
             a LocalDataFrame is accepted by fit()/transform() with or without pyspark
  barrier  : BarrierTaskContext (partitionId, allGather, barrier)   (pyspark.BarrierTaskContext) for the local tasks
"""
try:
    import pyspark  # noqa: F401

    HAVE_PYSPARK = True
except Exception:
    HAVE_PYSPARK = False

from .barrier import BarrierTaskContext  # noqa: E402,F401
from .sql import LocalDataFrame, LocalSession, get_session  # noqa: E402,F401

if HAVE_PYSPARK:
    from pyspark import keyword_only  # noqa: E402,F401
    from pyspark.ml import Estimator as EstimatorBase, Model as ModelBase  # noqa: E402,F401
    from pyspark.ml.param import Param, Params, TypeConverters  # noqa: E402,F401
    from pyspark.sql import Row  # noqa: E402,F401
else:
    from .params import Param, Params, TypeConverters, keyword_only  # noqa: E402,F401
    from .sql import Row  # noqa: E402,F401

    class EstimatorBase:  # pyspark.ml.Estimator's place in the MRO
        pass

    class ModelBase:      # pyspark.ml.Model's
        pass
