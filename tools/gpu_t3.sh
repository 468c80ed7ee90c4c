#!/bin/bash
bash tools/gpu_t1.sh 2>&1 | grep -v "^lloyd n="
bash tools/gpu_t2.sh
