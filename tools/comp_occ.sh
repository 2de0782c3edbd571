#!/bin/bash
# composite kernel: occupancy (min blocks per SM -> register cap) A/B
for mb in 3 4 5 6; do
  NSFF_NVCC_EXTRA=-DNSFF_COMP_MINB=$mb python -m nsff_pl_b200.build > /dev/null 2>&1
  echo "== min blocks/SM $mb"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d['roofline_composite']; print('step ms', round(d['ms_per_step'],2), 'composite ms', round(c['avg_launch_ms'],3), 'GB/s', round(c['achieved']), 'frac', round(c['frac'],3))
"
done
