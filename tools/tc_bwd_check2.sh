#!/bin/bash
mkdir -p gpurun_out
echo "=== backward tests"
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_tile_gemm.py -q --timeout=300 2>&1 | tail -25
echo "=== train bench (tensor-core backward)"
timeout 300 python bench.py --workload train --steps 5 --warmup 3 2>&1 | tail -2
echo "=== train bench (fp32 backward)"
timeout 300 python bench.py --workload train --steps 3 --warmup 3 --grad-precision fp32 2>&1 | tail -2
echo "=== launch list of one training step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/train_launches.csv \
  python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/train_ncu.log 2>&1; echo "ncu rc=$?"
