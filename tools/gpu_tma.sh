#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/tma_bw.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from spark_rapids_ml_b200 import _native
ctx = _native.Context(0)
for d in (128, 256):
    n = 12_500_000 if d == 256 else 10_000_000
    X = torch.randn((n, d), device="cuda")
    for nslot in (6, 12):
        for hold in (0, 4000):
            ms = min(ctx.debug_tma_stream(X, nslot, hold) for _ in range(3))
            print(f"d={d} nslot={nslot} hold={hold}: {ms:.3f} ms  {n*d*4/ms/1e6:.0f} GB/s", flush=True)
    del X
PY
cat gpurun_out/tma_bw.log
