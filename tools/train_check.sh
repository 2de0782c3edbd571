#!/bin/bash
# training-step iteration loop: backward / tile-GEMM / caller / loss tests, then bench.py --workload train with the chained
# backward (default) and the layer-wise one (NSFF_BWD_CHAIN=0), then an ncu launch list of one step.
# usage: tools/train_check.sh [notest] [tag]
mkdir -p gpurun_out
tag=${2:-train}
if [ "$1" != "notest" ]; then
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_tile_gemm.py tests/test_gpu_callers.py tests/test_gpu_loss.py -q -x --timeout=300 > gpurun_out/bwd_tests.log 2>&1
grep -E "^E  .*Error|^E  .*assert|^FAILED|passed|failed" gpurun_out/bwd_tests.log | head -40
fi
for chain in 1 0; do
echo "=== train bench NSFF_BWD_CHAIN=$chain"
NSFF_BWD_CHAIN=$chain timeout 300 python bench.py --workload train --steps 10 --warmup 3 2>gpurun_out/${tag}_bench_$chain.err | tee gpurun_out/${tag}_bench_$chain.json | python tools/print_train.py
tail -3 gpurun_out/${tag}_bench_$chain.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
python tools/step_breakdown.py gpurun_out/${tag}_launches.csv 16
