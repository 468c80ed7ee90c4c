class DataType:
    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __hash__(self):
        return hash(type(self).__name__)


class FloatType(DataType):
    pass


class DoubleType(DataType):
    pass


class IntegerType(DataType):
    pass


class LongType(DataType):
    pass


class StringType(DataType):
    pass


class ArrayType(DataType):
    def __init__(self, elementType, containsNull=True):
        self.elementType = elementType


class StructField:
    def __init__(self, name, dataType):
        self.name, self.dataType = name, dataType


class StructType:
    def __init__(self, fields):
        self.fields = list(fields)

    def __getitem__(self, name):
        for f in self.fields:
            if f.name == name:
                return f
        raise KeyError(name)

    @property
    def names(self):
        return [f.name for f in self.fields]
