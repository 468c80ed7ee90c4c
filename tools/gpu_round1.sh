#!/bin/bash
# first GPU bring-up: generic-path parity, fused-kernel debug, fused parity, short bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "=== generic parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "generic or ingest or init_modes" > gpurun_out/t_generic.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_generic.log
echo "=== fused debug"; timeout 300 python tools/debug_fused.py > gpurun_out/debug_fused.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/debug_fused.log
echo "=== fused parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tcgen05 or errors" > gpurun_out/t_fused.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_fused.log
echo "=== bench generic"; timeout 600 python bench.py --steps 10 --warmup 3 --kernel-path generic --no-e2e --no-cpu-baseline > gpurun_out/bench_generic.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_generic.log
echo "=== bench auto"; timeout 900 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_auto.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_auto.log
