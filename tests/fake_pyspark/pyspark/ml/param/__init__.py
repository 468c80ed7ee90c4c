from spark_rapids_ml_b200.sparkshim.params import Param, Params, TypeConverters  # noqa: F401
