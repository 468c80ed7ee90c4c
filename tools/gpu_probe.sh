#!/bin/bash
mkdir -p gpurun_out
export B2K_LIB=$PWD/spark_rapids_ml_b200/libb2kmeans_probe.so
B2K_PROBE_N=12 timeout 300 python tools/trace_t.py > gpurun_out/trace_t_p12.log 2>&1; head -8 gpurun_out/trace_t_p12.log | cut -c1-330; grep period gpurun_out/trace_t_p12.log
