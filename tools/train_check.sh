#!/bin/bash
mkdir -p gpurun_out
if [ "$1" != "notest" ]; then
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_tile_gemm.py -q --timeout=300 > gpurun_out/bwd_tests.log 2>&1
grep -E "^E  .*Error|^FAILED|passed|failed" gpurun_out/bwd_tests.log | head -40
fi
for flag in "" "--bwd-recompute"; do
echo "=== train bench $flag"
timeout 300 python bench.py --workload train --steps 10 --warmup 3 $flag 2>/dev/null | python tools/print_train.py
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/train_launches.csv \
  python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/train_ncu.log 2>&1; echo "ncu rc=$?"
