#!/bin/bash
# compute-sanitizer memcheck over the kernels' unit tests and a representative slice of the parity tests (GPU box)
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_tile_gemm.py -q -k "300 or 128-4" --timeout=200 2>&1 | tail -4
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_backward.py tests/test_gpu_parity.py -q \
  -k "saved or bf16_network or appearance or train_warp_bf16 or empty or frame_rays or interpolate or hierarchical_bf16 or static_only" --timeout=400 2>&1 | tail -4
