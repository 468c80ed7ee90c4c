"""Worker utilities on the KMeans path (reference: python/src/spark_rapids_ml/utils.py:138-170 GPU id from task
resources, :358-400 _concat_and_free, :403-522 reserved-buffer ingest, :555-576 logger) — re-designed around the
device-resident ingest of libb2kmeans (no host-side stacking, no second host copy)."""
from __future__ import annotations

import logging
import sys
from typing import Any, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd
import pyarrow as pa

_ArrayOrder = str  # "C" | "F"


def get_logger(cls_or_callable: Any, level: str = "INFO") -> logging.Logger:
    """reference: utils.py:555-576 — one stderr logger per class name."""
    name = cls_or_callable if isinstance(cls_or_callable, str) else getattr(cls_or_callable, "__name__", str(cls_or_callable))
    logger = logging.getLogger(f"spark_rapids_ml_b200.{name}")
    logger.setLevel(level)
    if not logger.handlers:
        h = logging.StreamHandler(sys.stderr)
        h.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
        logger.addHandler(h)
    return logger


def _get_gpu_id(task_context: Any) -> int:
    """reference: utils.py:138-170 — GPU address from the barrier task's resources, else CUDA_VISIBLE_DEVICES[0]."""
    import os

    res = task_context.resources() if task_context is not None else {}
    if "gpu" in res and res["gpu"].addresses:
        return int(res["gpu"].addresses[0].strip())
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        return 0  # first visible device
    raise RuntimeError("Couldn't get gpu id, please check the GPU resource configuration")


def _is_local(session: Any) -> bool:
    return True  # the shim session is always local mode; real Spark: sc._jsc.sc().isLocal() (utils.py:128-135)


class PartitionDescriptor:
    """reference: utils.py:300-355 — (m, n, rank, parts_rank_size) built from every rank's part sizes."""

    def __init__(self, m: int, n: int, rank: int, parts_rank_size: List[Tuple[int, int]]):
        self.m, self.n, self.rank, self.parts_rank_size = m, n, rank, parts_rank_size

    @classmethod
    def build(cls, partition_rows: List[int], total_cols: int) -> "PartitionDescriptor":
        import json

        from .sparkshim import BarrierTaskContext

        context = BarrierTaskContext.get()
        rank = context.partitionId()
        msgs = context.allGather(json.dumps((rank, partition_rows)))
        parts: List[Tuple[int, int]] = []
        total = 0
        for m_ in msgs:
            r, rows = json.loads(m_)
            for sz in rows:
                parts.append((r, sz))
                total += sz
        return cls(total, total_cols, rank, parts)


# ------------------------------------------------------------------------------------------------
# Arrow-buffer access for the ingest fast path
# ------------------------------------------------------------------------------------------------
def arrow_list_column_buffers(col: pd.Series, d: Optional[int] = None) -> Optional[Tuple[np.ndarray, np.ndarray, int]]:
    """If `col` is an Arrow-backed list<T> column, return (flat values ndarray view, int32 offsets, n_rows)
    WITHOUT copying; else None (object column of ndarrays -> the caller stacks on the host like the reference).
    `d`: the expected row width; a fixed_size_list of another width is an error (list<T> rows are validated against
    their offsets by the library)."""
    dt = col.dtype
    if not isinstance(dt, pd.ArrowDtype):
        return None
    arr = col.array._pa_array if hasattr(col.array, "_pa_array") else pa.chunked_array(col.array)
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.chunk(0) if arr.num_chunks == 1 else arr.combine_chunks()   # single chunk: stay zero-copy
    t = arr.type
    if arr.null_count:
        raise ValueError("null feature rows are not supported")
    if pa.types.is_fixed_size_list(t):
        width = t.list_size
        if d is not None and width != d:
            raise ValueError(f"feature rows are {width} wide, expected {d}")
        vals = arr.flatten().to_numpy(zero_copy_only=True)
        return vals, None, len(arr) if width else 0
    if pa.types.is_list(t):
        offsets = arr.offsets.to_numpy(zero_copy_only=True)
        vals = arr.values.to_numpy(zero_copy_only=True)  # full child buffer; offsets[0] locates the first row
        return vals, offsets, len(arr)
    return None


class DeviceRowAppender:
    """Growing device matrix [n, d] f32 fed batch by batch through b2k_ingest_append (replaces core.py:907-941
    + clustering.py:388-393).  Capacity grows geometrically by segments; segments are concatenated on the device
    once at the end (HBM copy, not a host copy)."""

    def __init__(self, ctx: Any, d: int, first_capacity: int = 1 << 20):
        import torch

        self._torch = torch
        self.ctx, self.d = ctx, d
        self.segments: List[Any] = []
        self.fill: List[int] = []
        self.next_cap = max(1024, first_capacity)

    def _room(self, n_b: int) -> Tuple[Any, int]:
        t = self._torch
        if not self.segments or self.fill[-1] + n_b > self.segments[-1].shape[0]:
            cap = max(self.next_cap, n_b)
            self.segments.append(t.empty((cap, self.d), dtype=t.float32, device=self.ctx.device))
            self.fill.append(0)
            self.next_cap = cap * 2
        return self.segments[-1], self.fill[-1]

    def append_values(self, values: np.ndarray, offsets: Optional[np.ndarray], n_b: int) -> None:
        seg, r0 = self._room(n_b)
        self.ctx.ingest_rows(seg, r0, values, self.d, offsets=offsets, n_rows=n_b)
        self.fill[-1] += n_b

    def append_columns(self, cols: Sequence[np.ndarray]) -> None:
        n_b = int(cols[0].shape[0])
        seg, r0 = self._room(n_b)
        self.ctx.ingest_columns(seg, r0, cols)
        self.fill[-1] += n_b

    @property
    def rows(self) -> int:
        return sum(self.fill)

    def finish(self) -> Any:
        t = self._torch
        if not self.segments:
            return t.empty((0, self.d), dtype=t.float32, device=self.ctx.device)
        if len(self.segments) == 1:
            return self.segments[0][: self.fill[0]]
        out = t.cat([s[:f] for s, f in zip(self.segments, self.fill)], dim=0)
        self.segments, self.fill = [], []
        return out
