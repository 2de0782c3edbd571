#!/bin/bash
# Round evidence: ncu launch list of a bench step, one --set full capture (with source) of the three field-kernel launches
# and one of the composite.  usage: tools/profile_round.sh <tag>      (outputs under gpurun_out/<tag>_*)
tag=${1:-r02}
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sub-records"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_forward_kernel -s 3 -c 3 \
  -o gpurun_out/${tag}_field_full -f $B > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 1 -c 1 \
  -o gpurun_out/${tag}_composite_full -f $B > gpurun_out/${tag}_ncu_comp.log 2>&1; echo "ncu composite rc=$?"
