"""Fake pyspark.sql: DataFrame / RDD wrappers that record the calls made on them and execute them on the library's
local Arrow-batch frame."""
import pandas as pd
import pyarrow as pa

from .. import CALLS
from .types import ArrayType, DoubleType, FloatType, IntegerType, LongType, StringType, StructField, StructType



class Row(dict):
    """pyspark.sql.Row subset: attribute + key access, asDict()."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def asDict(self):
        return dict(self)


def _to_spark_type(t):
    if pa.types.is_list(t) or pa.types.is_fixed_size_list(t) or pa.types.is_large_list(t):
        return ArrayType(_to_spark_type(t.value_type))
    return {pa.float32(): FloatType(), pa.float64(): DoubleType(), pa.int32(): IntegerType(), pa.int64(): LongType(),
            pa.string(): StringType()}.get(t, StringType())


CLUSTER_CONF = None   # a dict of Spark confs makes the fake session look like a cluster (tests of stage-level scheduling)


class _Conf:
    def get(self, key, default=None):
        return (CLUSTER_CONF or {}).get(key, default)


class _SC:
    @property
    def master(self):
        return (CLUSTER_CONF or {}).get("spark.master", "local[1]")

    def getConf(self):
        return _Conf()


class SparkSession:
    sparkContext = _SC()
    version = "3.5.1"
    conf = _Conf()

    @classmethod
    def getActiveSession(cls):
        return None


class RDD:
    def __init__(self, df, barrier=False):
        self._df, self._barrier = df, barrier

    def getNumPartitions(self):
        return self._df._local.getNumPartitions()

    def barrier(self):
        CALLS.append(("rdd.barrier", None))
        return RDD(self._df, True)

    def mapPartitions(self, f):
        CALLS.append(("rdd.mapPartitions", None))
        return self

    def withResources(self, profile):
        CALLS.append(("rdd.withResources", profile))
        return self

    def collect(self):
        CALLS.append(("rdd.collect", None))
        fn, schema = self._df._pending
        assert self._barrier, "the fit stage must be a barrier stage"
        if self._df._coalesced:
            raise RuntimeError("org.apache.spark.scheduler.BarrierJobUnsupportedRDDChainException: [SPARK-24820][SPARK-24821]: "
                               "Barrier execution mode does not allow the following pattern of RDD chain within a barrier stage")
        return self._df._local.mapInPandas(fn, schema=schema, barrier=True).collect()


class DataFrame:
    """Wraps a LocalDataFrame; `vector_cols`: columns presented as VectorUDT (stored as array<double>)."""

    def __init__(self, local, vector_cols=(), coalesced=False):
        self._local = local
        self._vector_cols = set(vector_cols)
        self._pending = None
        self._coalesced = coalesced   # Spark refuses a barrier stage on a coalesced RDD chain
        self.sparkSession = SparkSession()

    def coalesce(self, n):
        CALLS.append(("coalesce", n))
        return DataFrame(self._local.repartition(n), self._vector_cols, coalesced=True)

    @property
    def schema(self):
        from ..ml.linalg import VectorUDT

        return StructType([StructField(f.name, VectorUDT() if f.name in self._vector_cols else _to_spark_type(f.type))
                           for f in self._local.schema])

    @property
    def columns(self):
        return self._local.columns

    @property
    def rdd(self):
        return RDD(self)

    def select(self, *cols):
        CALLS.append(("select", [(c.name, c.out_name, c.cast_to) for c in cols]))
        df = self._local.select(*[c.name for c in cols])
        for c in cols:
            if c.cast_to is not None:
                t = pa.list_(pa.float32() if c.cast_to[1] == "float" else pa.float64()) if isinstance(c.cast_to, tuple) \
                    else (pa.float32() if c.cast_to == "float" else pa.float64())
                df = df.cast_column(c.name, t)
            if c.out_name != c.name:
                df = df.withColumnRenamed(c.name, c.out_name)
        return DataFrame(df, coalesced=self._coalesced)

    def first(self):
        return self._local.first()

    def repartition(self, n):
        CALLS.append(("repartition", n))
        return DataFrame(self._local.repartition(n), self._vector_cols)

    def mapInPandas(self, fn, schema=None):
        CALLS.append(("mapInPandas", str(schema)))
        out = DataFrame(self._local, self._vector_cols, coalesced=self._coalesced)
        out._pending = (fn, schema)
        return out

    def withColumn(self, name, column):
        CALLS.append(("withColumn", name))
        fn, cols = column.udf
        parts = []
        for part in self._local._parts:
            def frames():
                for b in part:
                    sel = b.select([c.name for c in cols])
                    pdf = sel.to_pandas(types_mapper=pd.ArrowDtype)
                    ren = {c.name: c.out_name for c in cols if c.out_name != c.name}
                    yield pdf.rename(columns=ren) if ren else pdf
            parts.append([pa.array(list(s), type=pa.int32()) for s in fn(frames())])
        return DataFrame(self._local.with_appended_column(name, parts), self._vector_cols)

    def collect(self):
        return self._local.collect()
