"""A RECORDING FAKE of the slice of pyspark that spark_rapids_ml_b200 touches — test infrastructure only
(tests/test_pyspark_binding.py).  There is no pyspark / JVM in this image; this package lets the tests import the
library the way it is imported next to a real pyspark (HAVE_PYSPARK = True) and checks WHICH pyspark calls the fit /
transform paths make (mapInPandas -> rdd.barrier().mapPartitions, pandas_udf + withColumn, vector_to_array), executing
them on the local Arrow-batch frame underneath.  Param / Params are the library's own local implementations re-exported
under the pyspark names (same API as pyspark.ml.param)."""
__version__ = "0.0-fake"
CALLS = []   # (what, detail) in call order


def keyword_only(func):
    import functools

    @functools.wraps(func)
    def wrapper(self, *args, **kwargs):
        if len(args) > 0:
            raise TypeError("Method %s forces keyword arguments." % func.__name__)
        self._input_kwargs = kwargs
        return func(self, **kwargs)

    return wrapper


class _TextRDD:
    def __init__(self, lines, path=None):
        self._lines, self._path = lines, path

    def saveAsTextFile(self, path):
        import os

        CALLS.append(("rdd.saveAsTextFile", path))
        os.makedirs(path)
        with open(os.path.join(path, "part-00000"), "w") as f:
            f.write("".join(l + "\n" for l in self._lines))
        open(os.path.join(path, "_SUCCESS"), "w").close()

    def collect(self):
        return list(self._lines)

    def first(self):
        return self._lines[0]


class SparkContext:
    _active_spark_context = None   # tests that exercise persistence through Spark install one
    master = "local[1]"

    def parallelize(self, data, numSlices=None):
        CALLS.append(("sc.parallelize", numSlices))
        return _TextRDD(list(data))

    def textFile(self, path):
        import os

        CALLS.append(("sc.textFile", path))
        lines = []
        for name in sorted(os.listdir(path)):
            if name.startswith("part-"):
                with open(os.path.join(path, name)) as f:
                    lines += [l.rstrip("\n") for l in f if l.strip()]
        return _TextRDD(lines, path)


class TaskContext:
    _pid = 0

    @classmethod
    def get(cls):
        return cls()

    def partitionId(self):
        return TaskContext._pid

    def resources(self):
        return {}


class BarrierTaskContext(TaskContext):
    @classmethod
    def get(cls):
        CALLS.append(("BarrierTaskContext.get", None))
        from spark_rapids_ml_b200.sparkshim.barrier import BarrierTaskContext as Local

        return Local.get()   # the local barrier task this fake stage runs in
