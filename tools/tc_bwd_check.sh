#!/bin/bash
# tensor-core backward bring-up: kernel unit checks (both MN-major descriptor variants), then the gradient tests
mkdir -p gpurun_out
for v in 0 1; do
  echo "=== NSFF_WG_DESC=$v"
  NSFF_WG_DESC=$v timeout 180 python tests/gpu_diag_tc_bwd.py 2>&1 | tail -16
done
echo "=== unit tests"
timeout 300 python -m pytest tests/test_gpu_tile_gemm.py -x -q --timeout=120 2>&1 | tail -8
echo "=== full backward diag"
timeout 300 python tests/gpu_diag_tc_bwd.py --full 2>&1 | tail -50
