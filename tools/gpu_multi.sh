#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_n$N.log | cut -c1-1500
if [ "$N" == "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5; fi
