#!/bin/bash
# merged N = 256 "pre" ops (NSFF_TC_MERGE, default 1) vs two N = 128 ops: parity under the watchdog build, then A/B bench lines
export NSFF_LIB=$PWD/nsff_pl_b200/libnsff_b200_wd.so
timeout -k 3 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_callers.py -x -q --timeout=120 > gpurun_out/wd_merge.log 2>&1
grep -E "WATCHDOG block" gpurun_out/wd_merge.log | head -5
grep -E "^E  .*Error|^E  .*assert|^FAILED|passed|failed" gpurun_out/wd_merge.log | head -10
if grep -q "failed\|WATCHDOG block" gpurun_out/wd_merge.log; then echo "STOP"; exit 0; fi
unset NSFF_LIB
line() { timeout -k 5 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sub-records 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'cycles(G)', round(d['ms_per_step']*d['clocks']['sm_mhz']/1e3,1), 'frac', round(d['roofline']['frac'],3))
"; }
for rep in 1 2; do
for v in 0 1; do echo "== NSFF_TC_MERGE=$v"; NSFF_TC_MERGE=$v line; done
done
echo "== train"; timeout -k 5 200 python bench.py --workload train --steps 20 --warmup 5 2>/dev/null | python tools/print_train.py
