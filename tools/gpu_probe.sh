#!/bin/bash
mkdir -p gpurun_out
export B2K_LIB=$PWD/spark_rapids_ml_b200/libb2kmeans_probe.so
timeout 300 python tools/trace_t.py > gpurun_out/trace_t.log 2>&1; head -10 gpurun_out/trace_t.log | cut -c1-330; grep period gpurun_out/trace_t.log
