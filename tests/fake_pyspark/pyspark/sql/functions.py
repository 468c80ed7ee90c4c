from .. import CALLS


class Column:
    def __init__(self, name):
        self.name = name
        self.out_name = name
        self.cast_to = None      # ("array", "float"|"double") or "float"/"double"
        self.udf = None          # (fn, [Column]) for an applied pandas_udf

    def alias(self, name):
        self.out_name = name
        return self

    def cast(self, t):
        from .types import ArrayType, FloatType

        if isinstance(t, ArrayType):
            self.cast_to = ("array", "float" if isinstance(t.elementType, FloatType) else "double")
        else:
            self.cast_to = "float" if isinstance(t, FloatType) else "double"
        return self


def col(name):
    return Column(name)


def struct(*cols):
    return list(cols)


def pandas_udf(return_type):
    CALLS.append(("pandas_udf", str(return_type)))

    def deco(fn):
        def apply(cols):
            c = Column("<udf>")
            c.udf = (fn, list(cols))
            return c

        return apply

    return deco
