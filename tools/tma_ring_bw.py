"""Bandwidth a TMA ring of N x 16 KB slots sustains per SM count, vs how long each slot is held (GPU diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spark_rapids_ml_b200 import _native
n, d = 10_000_000, 128
ctx = _native.Context(0)
X = torch.randn((n, d), device="cuda")
gb = n * d * 4 / 1e9
print("ring_slots hold_cycles  ms   GB/s")
for hold in (0, -4, -8, -16, -24):     # negative: that many extra warps polling an mbarrier with try_wait
    for nslot in (4, 9):
        ms = ctx.debug_tma_stream(X, nslot, hold)
        print(f"{nslot:6d} {hold:8d} {ms:8.3f} {gb / (ms / 1e3):8.0f}", flush=True)
