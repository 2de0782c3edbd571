#!/bin/bash
# end-of-round evidence with the final library: GPU tests, smoke, default bench, ncu launch list + --set full of the field kernel
tag=${1:-final}
bash tools/final_check.sh $tag
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:field_forward_kernel -s 3 -c 3 \
  -o gpurun_out/${tag}_field_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sub-records > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
