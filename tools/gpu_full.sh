#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_full.log | cut -c1-3000
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-900
