#!/bin/bash
# first contact of the large-shape kernel with hardware: assign first, then lloyd; each under its own timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader > gpurun_out/gpu.txt
timeout 300 python tools/debug_t.py assign > gpurun_out/t_assign.log 2>&1; echo "assign rc=$?" >> gpurun_out/t_assign.log
tail -20 gpurun_out/t_assign.log
timeout 300 python tools/debug_t.py lloyd > gpurun_out/t_lloyd.log 2>&1; echo "lloyd rc=$?" >> gpurun_out/t_lloyd.log
tail -12 gpurun_out/t_lloyd.log
