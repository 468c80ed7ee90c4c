#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_fused.py > gpurun_out/debug_fused.log 2>&1; echo "debug rc=$?"; grep -c "outside_margin=0" gpurun_out/debug_fused.log; grep -E "outside_margin=[1-9]|Error|error|timeout|center_rel_err=[0-9.]*e-0[0-3]" gpurun_out/debug_fused.log | head
timeout 300 python tools/profile_roles.py 10000000 2>&1 | tail -7
timeout 600 python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
l=open('gpurun_out/bench_quick.log').read().strip().split('\n')[-1]
try:
    j=json.loads(l); print('ms_per_step',j['ms_per_step'],'kernel_ms',j['roofline'].get('kernel_ms'),'frac',j['roofline']['frac'])
except Exception as e: print(l[-600:])
PY
