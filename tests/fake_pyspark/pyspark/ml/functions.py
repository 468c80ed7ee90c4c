from .. import CALLS
from ..sql.functions import Column


def vector_to_array(col, dtype="float64"):
    CALLS.append(("vector_to_array", (col.name, dtype)))
    out = Column(col.name)
    out.cast_to = ("array", "float" if dtype == "float32" else "double")
    return out


def array_to_vector(col):
    CALLS.append(("array_to_vector", col.name))
    return col
