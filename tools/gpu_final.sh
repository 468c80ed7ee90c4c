#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "=== full capture"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -f -o gpurun_out/fused_full \
   python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1; echo "rc=$?"
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_full.log | cut -c1-2600
