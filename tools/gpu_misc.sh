#!/bin/bash
mkdir -p gpurun_out
echo "== tests S12d (installed)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
cp spark_rapids_ml_b200/libb2kmeans.so /tmp/orig.so
for L in XB XC1; do
cp tools/_ab/$L.so spark_rapids_ml_b200/libb2kmeans.so
echo "== tests $L"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1
done
cp /tmp/orig.so spark_rapids_ml_b200/libb2kmeans.so
timeout 1500 python tools/ab2.py tools/_ab/V8.so tools/_ab/S12d.so tools/_ab/XA.so tools/_ab/XB.so tools/_ab/XC1.so tools/_ab/XC2.so
