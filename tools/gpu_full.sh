#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== debug"; timeout 300 python tools/debug_fused.py 2>&1 | tail -2
