#!/bin/bash
# A/B: single-CTA kernel vs CTA-pair kernel (parity subset, then bench)
for mode in 1 2; do
  echo "=== NSFF_TC_CTAS=$mode parity"
  NSFF_TC_CTAS=$mode timeout 200 python -m pytest tests/test_gpu_parity.py -q -x --timeout=120 --tb=short -k "bf16" 2>&1 | tail -6
  echo "=== NSFF_TC_CTAS=$mode bench"
  NSFF_TC_CTAS=$mode timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step',d['ms_per_step'],'value',d['value'],'frac',d['roofline']['frac'],d['clocks'])
    elif 'rror' in l: print(l.strip())
"
done
