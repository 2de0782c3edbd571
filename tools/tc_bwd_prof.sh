#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:img_gemm_kernel -s 150 -c 12 \
  -o gpurun_out/bwd_gemm_full -f python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/bwd_gemm_ncu.log 2>&1; echo "ncu gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:img_wgrad_kernel -s 60 -c 4 \
  -o gpurun_out/bwd_wgrad_full -f python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/bwd_wgrad_ncu.log 2>&1; echo "ncu wgrad rc=$?"
