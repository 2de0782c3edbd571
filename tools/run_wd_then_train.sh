export NSFF_LIB=$PWD/nsff_pl_b200/libnsff_b200_wd.so
timeout -k 3 400 python -m pytest tests/test_gpu_backward.py tests/test_gpu_tile_gemm.py tests/test_gpu_callers.py tests/test_gpu_loss.py -x -q --timeout=120 > gpurun_out/wd_tests.log 2>&1
grep -E "WATCHDOG block" gpurun_out/wd_tests.log | head -5
grep -E "^E  .*Error|^E  .*assert|^FAILED|passed|failed" gpurun_out/wd_tests.log | head -20
if grep -q "failed\|WATCHDOG block" gpurun_out/wd_tests.log; then echo "STOP: failures under the watchdog build"; exit 0; fi
unset NSFF_LIB
timeout -k 3 900 bash tools/train_check.sh notest r02w
