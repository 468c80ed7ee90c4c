"""Fake pyspark.resource.profile."""


class ResourceProfileBuilder:
    def __init__(self):
        self._t = None

    def require(self, treqs):
        self._t = treqs
        return self

    @property
    def build(self):
        return dict(self._t.req)
