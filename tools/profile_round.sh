#!/bin/bash
# Round evidence: ncu launch list of a bench step, --set full captures (with source) of the three field-kernel launches of
# a frame, of the staged composite, and of the training step's backward chain and weight-gradient kernels.
# usage: tools/profile_round.sh <tag>      (outputs under gpurun_out/<tag>_*)
tag=${1:-r02}
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sub-records"
T="python bench.py --workload train --steps 1 --warmup 3 --no-train-graph"
mkdir -p gpurun_out
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:field_forward_kernel -s 3 -c 3 \
  -o gpurun_out/${tag}_field_full -f $B > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:composite_staged_kernel -s 1 -c 1 \
  -o gpurun_out/${tag}_composite_full -f $B > gpurun_out/${tag}_ncu_comp.log 2>&1; echo "ncu composite rc=$?"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k "regex:field_forward_kernel<0, 1, 1>" -s 9 -c 3 \
  -o gpurun_out/${tag}_chain_full -f $T > gpurun_out/${tag}_ncu_chain.log 2>&1; echo "ncu chain rc=$?"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:img_wgrad_kernel -s 26 -c 2 \
  -o gpurun_out/${tag}_wgrad_full -f $T > gpurun_out/${tag}_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
ls -la gpurun_out/${tag}_*.ncu-rep
