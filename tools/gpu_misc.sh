#!/bin/bash
mkdir -p gpurun_out
echo "== tests V8 (installed)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 1200 python tools/ab2.py tools/_ab/N.so tools/_ab/V7.so tools/_ab/V8.so
