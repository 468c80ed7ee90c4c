"""pyspark.ml.param of the recording fake: the library's own local Param / Params implementation (same API as
pyspark.ml.param), loaded straight from its source file so that importing the fake never imports the library package
(the library imports pyspark, not the other way round)."""
import importlib.util
import os

_src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "..", "..", "spark_rapids_ml_b200",
                    "sparkshim", "params.py")
_spec = importlib.util.spec_from_file_location("_fake_pyspark_params_impl", os.path.normpath(_src))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
Param, Params, TypeConverters = _mod.Param, _mod.Params, _mod.TypeConverters
