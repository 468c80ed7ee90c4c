#!/bin/bash
# usage: gpu_ab.sh [-t] <new.so> [more .so ...]
#   -t: run the parity tests on the first build; always: role profile per build (gpurun_out/prof_<name>.log), then a
#   same-box alternating A/B of tools/_ab/base.so and all the builds
mkdir -p gpurun_out
TEST=0; if [ "$1" == "-t" ]; then TEST=1; shift; fi
cp spark_rapids_ml_b200/libb2kmeans.so /tmp/orig.so
if [ $TEST == 1 ]; then
  cp $1 spark_rapids_ml_b200/libb2kmeans.so
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
fi
for L in "$@"; do
  cp $L spark_rapids_ml_b200/libb2kmeans.so
  timeout 300 python tools/profile_roles.py 10000000 > gpurun_out/prof_$(basename $L .so).log 2>&1
done
timeout 1200 python tools/ab2.py tools/_ab/base.so "$@" 2>&1 | tail -16
cp /tmp/orig.so spark_rapids_ml_b200/libb2kmeans.so
