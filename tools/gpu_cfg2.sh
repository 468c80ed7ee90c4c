#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
python bench.py --config cfg2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-cfg3 --long-steps 200 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']; print('cfg2 kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), 'long frac', round(r['long_run']['frac'],4), r['long_run']['clocks']['sm_mhz'], [round(x,4) for x in j['repeat_ms_per_step']])
"
done
