#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -f -o gpurun_out/fused_full \
   python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1; echo "rc=$?"
