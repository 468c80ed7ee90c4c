"""BarrierTaskContext stand-in: what a Spark barrier task offers the worker (core.py:856-861, cuml_context.py:75-81,
utils.py:345-349): partitionId(), allGather(str) -> List[str], barrier(), resources().

Backed by a torch.distributed TCPStore (no GPU needed), so world_size-2 CPU tests exercise the same rendezvous
the GPU workers use."""
from __future__ import annotations

import os
import threading
from typing import Any, Dict, List, Optional


class _Resource:
    def __init__(self, addresses: List[str]):
        self.addresses = addresses


class BarrierTaskContext:
    _current = threading.local()

    def __init__(self, partition_id: int, num_tasks: int, store: Any = None, gpu_addresses: Optional[List[str]] = None):
        self._pid = int(partition_id)
        self._n = int(num_tasks)
        self._store = store
        self._epoch = 0
        self._gpus = gpu_addresses

    # -- Spark API surface --
    @classmethod
    def get(cls) -> "BarrierTaskContext":
        ctx = getattr(cls._current, "ctx", None)
        if ctx is None:
            raise RuntimeError("It is not in a barrier stage")
        return ctx

    def partitionId(self) -> int:
        return self._pid

    def resources(self) -> Dict[str, _Resource]:
        return {"gpu": _Resource(self._gpus)} if self._gpus else {}

    def getTaskInfos(self) -> List[Any]:
        return [None] * self._n

    def allGather(self, message: str = "") -> List[str]:
        if self._n == 1 or self._store is None:
            return [message]
        ep = self._epoch
        self._epoch += 1
        self._store.set(f"ag/{ep}/{self._pid}", message.encode())
        out = []
        for r in range(self._n):
            out.append(self._store.get(f"ag/{ep}/{r}").decode())  # blocks until the key exists
        return out

    def barrier(self) -> None:
        self.allGather("")

    # -- harness --
    @classmethod
    def _install(cls, ctx: Optional["BarrierTaskContext"]) -> None:
        cls._current.ctx = ctx


def make_store(rank: int, world: int, port: int, host: str = "127.0.0.1") -> Any:
    """TCPStore shared by the barrier tasks of one stage (rank 0 hosts it)."""
    from datetime import timedelta

    from torch.distributed import TCPStore

    return TCPStore(host, port, world, is_master=(rank == 0), timeout=timedelta(seconds=300),
                    wait_for_workers=True)


def free_port() -> int:
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]
