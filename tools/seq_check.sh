#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload sequence --steps 4 --warmup 3 > gpurun_out/seq_bench.json 2> gpurun_out/seq_bench.err; echo "rc=$?"; cat gpurun_out/seq_bench.json; tail -3 gpurun_out/seq_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/seq_launches.csv \
  python bench.py --workload sequence --steps 1 --warmup 3 > gpurun_out/seq_ncu.log 2>&1; echo "ncu rc=$?"
