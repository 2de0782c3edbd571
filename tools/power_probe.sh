#!/bin/bash
for mode in 1 2; do for steps in 3 30; do
  echo "== CTAS=$mode steps=$steps"
  NSFF_TC_CTAS=$mode timeout 300 python bench.py --steps $steps --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step',round(d['ms_per_step'],2),'frac',round(d['roofline']['frac'],3),d['clocks'])
"
done; done
