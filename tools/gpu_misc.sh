#!/bin/bash
mkdir -p gpurun_out
echo "== tests V7 (installed)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 1200 python tools/ab2.py tools/_ab/V2.so tools/_ab/V6.so tools/_ab/V7.so
