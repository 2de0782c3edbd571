#!/bin/bash
# needs a build with NSFF_NVCC_EXTRA=-DNSFF_EXPERIMENT; results are WRONG by design, only timing matters
for mode in 1 2; do for f in 0 1 3 7; do
  echo "== CTAS=$mode FLAGS=$f"
  NSFF_TC_CTAS=$mode NSFF_EXP_FLAGS=$f timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step',round(d['ms_per_step'],2),'frac',round(d['roofline']['frac'],3),d['clocks'])
"
done; done
