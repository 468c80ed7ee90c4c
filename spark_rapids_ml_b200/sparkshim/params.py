"""pyspark.ml.param look-alike (Param, Params, TypeConverters, keyword_only) — just enough for the
reference's _CumlParams semantics (params.py:260-707) and its tests to run without a JVM."""
from __future__ import annotations

import copy as _copy
import functools
import uuid
from typing import Any, Callable, Dict, List, Optional


class TypeConverters:
    @staticmethod
    def identity(v: Any) -> Any:
        return v

    @staticmethod
    def toInt(v: Any) -> int:
        if isinstance(v, bool) or not isinstance(v, (int, float)) or int(v) != v:
            raise TypeError(f"Could not convert {v!r} to int")
        return int(v)

    @staticmethod
    def toFloat(v: Any) -> float:
        if isinstance(v, bool) or not isinstance(v, (int, float)):
            raise TypeError(f"Could not convert {v!r} to float")
        return float(v)

    @staticmethod
    def toString(v: Any) -> str:
        if not isinstance(v, str):
            raise TypeError(f"Could not convert {type(v)} to string type")
        return v

    @staticmethod
    def toListString(v: Any) -> List[str]:
        if not isinstance(v, (list, tuple)) or not all(isinstance(x, str) for x in v):
            raise TypeError(f"Could not convert {v!r} to list of strings")
        return list(v)


class Param:
    def __init__(self, parent: Any, name: str, doc: str, typeConverter: Optional[Callable] = None):
        self.parent = parent.uid if hasattr(parent, "uid") else str(parent)
        self.name = name
        self.doc = doc
        self.typeConverter = typeConverter or TypeConverters.identity

    def __repr__(self) -> str:
        return f"Param(parent={self.parent!r}, name={self.name!r})"

    def __hash__(self) -> int:
        return hash((self.parent, self.name))

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, Param) and self.parent == other.parent and self.name == other.name


def keyword_only(func: Callable) -> Callable:
    """Stores the keyword arguments of the call in self._input_kwargs (pyspark.keyword_only)."""

    @functools.wraps(func)
    def wrapper(self: Any, *args: Any, **kwargs: Any) -> Any:
        if args:
            raise TypeError(f"Method {func.__name__} forces keyword arguments.")
        self._input_kwargs = kwargs
        return func(self, **kwargs)

    return wrapper


class Params:
    """Instance-level param maps with class-level Param declarations (pyspark.ml.param.Params subset)."""

    def __init__(self) -> None:
        if not hasattr(self, "uid"):
            self.uid = f"{type(self).__name__}_{uuid.uuid4().hex[:12]}"
        self._paramMap: Dict[Param, Any] = getattr(self, "_paramMap", {})
        self._defaultParamMap: Dict[Param, Any] = getattr(self, "_defaultParamMap", {})
        # re-parent class-level Param declarations onto this instance (as pyspark does)
        for klass in type(self).__mro__:
            for name, val in vars(klass).items():
                if isinstance(val, Param) and not isinstance(self.__dict__.get(name), Param):
                    p = Param(self, val.name, val.doc, val.typeConverter)
                    setattr(self, name, p)

    # --- queries ---
    @property
    def params(self) -> List[Param]:
        return sorted([v for v in self.__dict__.values() if isinstance(v, Param)], key=lambda p: p.name)

    def hasParam(self, name: str) -> bool:
        return isinstance(name, str) and isinstance(self.__dict__.get(name), Param)

    def getParam(self, name: str) -> Param:
        p = self.__dict__.get(name)
        if not isinstance(p, Param):
            raise ValueError(f"Cannot find param with name {name}.")
        return p

    def _resolve(self, param: Any) -> Param:
        return self.getParam(param) if isinstance(param, str) else self.getParam(param.name)

    def isSet(self, param: Any) -> bool:
        return self._resolve(param) in self._paramMap

    def hasDefault(self, param: Any) -> bool:
        try:
            return self._resolve(param) in self._defaultParamMap
        except ValueError:
            return False

    def isDefined(self, param: Any) -> bool:
        return self.isSet(param) or self.hasDefault(param)

    def getOrDefault(self, param: Any) -> Any:
        p = self._resolve(param)
        if p in self._paramMap:
            return self._paramMap[p]
        if p in self._defaultParamMap:
            return self._defaultParamMap[p]
        raise KeyError(f"Param {p.name} is not set and has no default")

    def extractParamMap(self, extra: Optional[Dict[Param, Any]] = None) -> Dict[Param, Any]:
        m = dict(self._defaultParamMap)
        m.update(self._paramMap)
        if extra:
            m.update(extra)
        return m

    def _resetUid(self, newUid: str) -> "Params":
        """pyspark.ml.param.Params._resetUid: change the uid and re-parent every Param (and the two maps)."""
        newUid = str(newUid)
        self.uid = newUid
        remap: Dict[Param, Param] = {}
        for name, val in list(self.__dict__.items()):
            if isinstance(val, Param):
                q = Param(self, val.name, val.doc, val.typeConverter)
                remap[val] = q
                setattr(self, name, q)
        self._paramMap = {remap.get(p, p): v for p, v in self._paramMap.items()}
        self._defaultParamMap = {remap.get(p, p): v for p, v in self._defaultParamMap.items()}
        return self

    # --- mutation ---
    def _set(self, **kwargs: Any) -> "Params":
        for name, value in kwargs.items():
            p = self.getParam(name)
            if value is not None:
                try:
                    value = p.typeConverter(value)
                except (TypeError, ValueError) as e:
                    raise TypeError(f'Invalid param value given for param "{name}". {e}')
            self._paramMap[p] = value
        return self

    def set(self, param: Param, value: Any) -> None:
        self._set(**{param.name: value})

    def _setDefault(self, **kwargs: Any) -> "Params":
        for name, value in kwargs.items():
            self._defaultParamMap[self.getParam(name)] = value
        return self

    def clear(self, param: Param) -> None:
        self._paramMap.pop(self._resolve(param), None)

    def copy(self, extra: Optional[Dict[Param, Any]] = None) -> "Params":
        that = _copy.copy(self)
        that._paramMap = {}
        that._defaultParamMap = {}
        for p in list(that.__dict__.values()):
            pass
        # rebuild Param objects so that they belong to the copy (same uid, as pyspark's copy keeps uid)
        for name, val in list(self.__dict__.items()):
            if isinstance(val, Param):
                setattr(that, name, Param(that, val.name, val.doc, val.typeConverter))
        for p, v in self._defaultParamMap.items():
            that._defaultParamMap[that.getParam(p.name)] = v
        for p, v in self._paramMap.items():
            that._paramMap[that.getParam(p.name)] = v
        if extra:
            for p, v in extra.items():
                that._paramMap[that.getParam(p.name)] = v
        return that

    def _copyValues(self, to: "Params", extra: Optional[Dict[Param, Any]] = None) -> "Params":
        pm = self.extractParamMap(extra)
        for p, v in pm.items():
            if to.hasParam(p.name):
                if p in self._defaultParamMap and p not in self._paramMap and not (extra and p in extra):
                    to._defaultParamMap[to.getParam(p.name)] = v
                else:
                    to._paramMap[to.getParam(p.name)] = v
        return to

    def explainParams(self) -> str:
        return "\n".join(f"{p.name}: {p.doc}" for p in self.params)
