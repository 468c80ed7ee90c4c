#!/bin/bash
mkdir -p gpurun_out
export B2K_LIB=$PWD/spark_rapids_ml_b200/libb2kmeans_probe.so
for p in 0 10 11; do
timeout 300 python bench.py --config cfg3 --init near_true --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --probe $p > gpurun_out/probe_$p.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/probe_$p.log"):
    if l.startswith("{"):
        j=json.loads(l); r=j["roofline"]; print("probe $p kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"],3), j["clocks"]["sm_mhz"])
    elif "rror" in l: print(l.strip()[:200])
PY
done
