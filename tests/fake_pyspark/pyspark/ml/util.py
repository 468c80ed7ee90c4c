"""Recording fake of pyspark.ml.util (persistence): MLWriter / MLReader over the fake SparkContext, DefaultParamsWriter /
DefaultParamsReader writing Spark's metadata layout (path/metadata/part-00000, one JSON line)."""
import json
import os
import shutil
import time

from pyspark import CALLS, SparkContext


class _HasSC:
    @property
    def sc(self):
        return SparkContext._active_spark_context


class MLWriter(_HasSC):
    def __init__(self):
        self.shouldOverwrite = False

    def overwrite(self):
        self.shouldOverwrite = True
        return self

    def save(self, path):
        CALLS.append(("MLWriter.save", path))
        if os.path.exists(path):
            if not self.shouldOverwrite:
                raise IOError("Path %s already exists. To overwrite it, please use write.overwrite().save(path)" % path)
            shutil.rmtree(path)
        self.saveImpl(path)

    def saveImpl(self, path):
        raise NotImplementedError


class MLReader(_HasSC):
    def load(self, path):
        raise NotImplementedError


class DefaultParamsWriter(MLWriter):
    @staticmethod
    def saveMetadata(instance, path, sc, extraMetadata=None, paramMap=None):
        CALLS.append(("DefaultParamsWriter.saveMetadata", path))
        cls = instance.__module__ + "." + instance.__class__.__name__
        meta = {"class": cls, "timestamp": int(time.time() * 1000), "sparkVersion": "3.5.1", "uid": instance.uid,
                "paramMap": {p.name: v for p, v in instance._paramMap.items()},
                "defaultParamMap": {p.name: v for p, v in instance._defaultParamMap.items()}}
        if extraMetadata:
            meta.update(extraMetadata)
        sc.parallelize([json.dumps(meta, separators=(",", ":"))], 1).saveAsTextFile(os.path.join(path, "metadata"))


class DefaultParamsReader(MLReader):
    @staticmethod
    def loadMetadata(path, sc, expectedClassName=""):
        CALLS.append(("DefaultParamsReader.loadMetadata", path))
        meta = json.loads(sc.textFile(os.path.join(path, "metadata")).first())
        if expectedClassName:
            assert meta["class"] == expectedClassName
        return meta

    @staticmethod
    def getAndSetParams(instance, metadata, skipParams=None):
        CALLS.append(("DefaultParamsReader.getAndSetParams", None))
        for name, v in metadata.get("defaultParamMap", {}).items():
            if instance.hasParam(name):
                instance._setDefault(**{name: v})
        for name, v in metadata["paramMap"].items():
            if instance.hasParam(name):
                instance._set(**{name: v})
