from .param import Params


class Estimator(Params):
    def fit(self, dataset, params=None):
        est = self.copy(params) if params else self
        return est._fit(dataset)


class Transformer(Params):
    def transform(self, dataset, params=None):
        return self._transform(dataset)


class Model(Transformer):
    pass
