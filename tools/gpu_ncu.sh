#!/bin/bash
# ncu launch list + one full capture of the fused kernel (inputs of tools/make_profiles.py)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -f -o gpurun_out/fused_full \
   python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1; echo "rc=$?"
