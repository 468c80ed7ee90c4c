#!/bin/bash
# r02 ncu evidence (inputs of tools/make_profiles.py): launch lists + one --set full capture of the dominant kernel,
# for BASELINE cfg3's shape (k_fused_t) and for cfg2 (k_fused_assign_update).  One GPU; numbers printed under ncu are
# never bench values.
mkdir -p gpurun_out
WHICH=${1:-both}   # cfg2 | cfg3 | both
B3="python bench.py --config cfg3 --init near_true --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-cfg3 --long-steps 0"
B2="python bench.py --config cfg2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-cfg3 --long-steps 0"
[ "$WHICH" != cfg2 ] && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches_cfg3.csv $B3 > gpurun_out/ncu_l3.log 2>&1; echo "launches cfg3 rc=$?"
[ "$WHICH" != cfg2 ] && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_t -s 4 -c 1 -f -o gpurun_out/fused_t_full $B3 > gpurun_out/ncu_f3.log 2>&1; echo "full cfg3 rc=$?"
[ "$WHICH" != cfg3 ] && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches_cfg2.csv $B2 > gpurun_out/ncu_l2.log 2>&1; echo "launches cfg2 rc=$?"
[ "$WHICH" != cfg3 ] && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_assign -s 4 -c 1 -f -o gpurun_out/fused_full $B2 > gpurun_out/ncu_f2.log 2>&1; echo "full cfg2 rc=$?"
ls -la gpurun_out/*.ncu-rep
