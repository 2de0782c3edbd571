#!/bin/bash
# One ncu --set full capture (with source-level stall sampling) of the first field_forward_kernel launch of a bench step.
# usage: tools/ncu_field.sh <tag>       -> gpurun_out/<tag>_field_full.ncu-rep
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_forward_kernel -s 3 -c 1 \
  -o gpurun_out/${tag}_field_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_ncu_full.log
