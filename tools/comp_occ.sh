#!/bin/bash
# composite kernel A/B: direct kernel vs the staged kernel at several (stages, warps per block) points.
# Rebuilds on the box; the shipped library is restored by the default build at the end.
line() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sub-records 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d['roofline_composite']; print('step ms', round(d['ms_per_step'],2), 'composite ms', round(c['avg_launch_ms'],3), 'GB/s', round(c['achieved']), 'frac', round(c['frac'],3))
"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=200 --tb=short -k "staged_composite" 2>&1 | tail -4
echo "== direct"; NSFF_COMP_DIRECT=1 line
echo "== staged default"; line
for cfg in "2 18" "2 20" "2 22" "3 14"; do
  set -- $cfg
  NSFF_NVCC_EXTRA="-DNSFF_COMP_STAGES=$1 -DNSFF_COMP_WARPS=$2" python -m nsff_pl_b200.build > /dev/null 2>&1
  echo "== staged stages=$1 warps=$2"; line
done
python -m nsff_pl_b200.build > /dev/null 2>&1
