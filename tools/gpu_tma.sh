#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/tma_bw.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from spark_rapids_ml_b200 import _native
ctx = _native.Context(0)
d = 256; n = 12_500_000
X = torch.randn((n, d), device="cuda")
for box in (128, 64, 32):
    ctx.set_option("tma_box_rows", box)
    for nslot in (8, 16, 24):
        if nslot * box > 13 * 128: continue
        for hold in (0, 2000, 5000):
            ms = min(ctx.debug_tma_stream(X, nslot, hold) for _ in range(3))
            print(f"box_rows={box} nslot={nslot} ({nslot*box*128//1024} KB) hold={hold}: {ms:.3f} ms  {n*d*4/ms/1e6:.0f} GB/s", flush=True)
PY
cat gpurun_out/tma_bw.log
