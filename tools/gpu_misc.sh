#!/bin/bash
mkdir -p gpurun_out
echo "== tests Q"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_estimator.py -x -q 2>&1 | tail -2
timeout 300 python tools/trace_tiles.py > gpurun_out/trace_Q.log 2>&1; tail -1 gpurun_out/trace_Q.log | cut -c1-400
timeout 1200 python tools/ab2.py tools/_ab/N.so tools/_ab/Q.so
