"""Stand-ins for the STOCK pyspark.ml.clustering classes (what install.py's proxy must keep returning to pyspark.ml
itself and for names it does not accelerate)."""


class KMeans:
    stock = True


class KMeansModel:
    stock = True


class BisectingKMeans:
    stock = True


def _sibling_lookup():
    """pyspark.ml code importing its own sibling must keep seeing the stock class (the file path contains /pyspark/ml/)."""
    import sys

    return sys.modules["pyspark.ml.clustering"].KMeans
