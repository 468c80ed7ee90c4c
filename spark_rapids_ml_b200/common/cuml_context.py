"""CumlContext: NCCL communicator lifecycle for one barrier task (reference: common/cuml_context.py:35-175).

Same protocol as the reference — rank 0 creates the NCCL unique id, it travels base64-encoded through
BarrierTaskContext.allGather (cuml_context.py:75-81), every rank calls ncclCommInitRank (:123-131); on exit the
communicator is destroyed, or ABORTED when an exception is in flight so peers do not hang (:158-175).  The
"handle" given to the fit function is the libb2kmeans context (the raft Handle's stand-in).  The UCX branch of the
reference is not built: KMeans is collectives-only (core.py:564-571)."""
from __future__ import annotations

import base64
from typing import Any, Optional

from .. import _native
from ..utils import get_logger


class CumlContext:
    def __init__(self, rank: int, nranks: int, context: Any, enable: bool, require_ucx: bool = False,
                 device: Optional[int] = None) -> None:
        if require_ucx:
            raise NotImplementedError("UCX p2p endpoints are out of scope for the KMeans path")
        self._rank, self._nranks, self._context = rank, nranks, context
        self.enable = enable
        self._device = device if device is not None else 0
        self._handle: Optional[_native.Context] = None
        self._loop = None  # reference exposes an asyncio loop for UCX; unused here
        self._logger = get_logger(type(self))
        self._uid: Optional[bytes] = None
        if enable and nranks > 1:
            msg = base64.b64encode(_native.comm_unique_id()).decode() if rank == 0 else ""
            msgs = context.allGather(msg)
            self._uid = base64.b64decode(msgs[0])

    @property
    def handle(self) -> _native.Context:
        assert self._handle is not None
        return self._handle

    def __enter__(self) -> "CumlContext":
        self._handle = _native.Context(self._device)
        if self.enable and self._nranks > 1:
            assert self._uid is not None
            self._handle.comm_init(self._nranks, self._rank, self._uid)
        return self

    def __exit__(self, exc_type: Any, *args: Any) -> None:
        if self._handle is None:
            return
        try:
            if self.enable and self._nranks > 1:
                if exc_type is None:
                    self._handle.comm_destroy()
                else:
                    self._handle.comm_abort()   # do not block on peers that may already be dead
        finally:
            self._handle.close()
            self._handle = None
