"""ctypes binding for oracle/kmeans_oracle.c — TEST INFRASTRUCTURE ONLY (see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkmeans_oracle.so")
_lib: Optional[ctypes.CDLL] = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kmeans_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_assign.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_assign.restype = None
        L.oracle_lloyd.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
                                   ctypes.c_void_p]
        L.oracle_lloyd.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_set_threads.restype = None
        L.oracle_parallel_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]
        L.oracle_parallel_copy.restype = None
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def first_touch_copy(X: np.ndarray) -> np.ndarray:
    """A copy of X whose pages are first touched by the OpenMP threads that will read them."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    out = np.empty_like(X)
    lib().oracle_parallel_copy(out.ctypes.data, X.ctypes.data, X.shape[0], X.shape[1])
    return out


def assign(X: np.ndarray, C: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    X = np.ascontiguousarray(X, dtype=np.float32)
    C = np.ascontiguousarray(C, dtype=np.float32)
    n, d = X.shape
    labels = np.empty(n, dtype=np.int32)
    mind = np.empty(n, dtype=np.float64)
    lib().oracle_assign(X.ctypes.data, n, d, C.ctypes.data, C.shape[0], labels.ctypes.data,
                        mind.ctypes.data)
    return labels, mind


def lloyd(X: np.ndarray, C0: np.ndarray, max_iter: int, tol: float, want_labels: bool = True):
    X = np.ascontiguousarray(X, dtype=np.float32)
    C = np.array(C0, dtype=np.float32, order="C")
    n, d = X.shape
    tol = float(np.finfo("float32").tiny) if tol == 0.0 else float(tol)
    inertia = ctypes.c_double(0.0)
    labels = np.empty(n, dtype=np.int32) if want_labels else None
    n_iter = lib().oracle_lloyd(X.ctypes.data, n, d, C.ctypes.data, C.shape[0], int(max_iter), tol,
                                ctypes.byref(inertia),
                                labels.ctypes.data if labels is not None else None)
    return {"centers": C, "n_iter": int(n_iter), "inertia": float(inertia.value), "labels": labels}
