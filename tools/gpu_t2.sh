#!/bin/bash
mkdir -p gpurun_out
for init in near_true first_k; do
timeout 600 python bench.py --config cfg3 --init $init --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg3_$init.log 2>&1; echo "rc=$?" >> gpurun_out/bench_cfg3_$init.log
python - <<PY
import json
for l in open("gpurun_out/bench_cfg3_$init.log"):
    if l.startswith("{"):
        j=json.loads(l); r=j["roofline"]; print("$init", "ms/step", round(j["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"],3), "recheck/iter", r.get("recheck_rows_per_iter"), r.get("recheck_candidates_per_iter"), j["clocks"])
    elif "rc=" in l or "rror" in l: print(l.strip())
PY
done
