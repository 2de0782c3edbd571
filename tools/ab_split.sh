#!/bin/bash
# A/B of the field kernel's two accumulator modes (NSFF_TC_SPLIT=1: N=128 halves, default; 0: whole-layer accumulators):
# parity tests first, then a short bench line each.  usage: tools/ab_split.sh [tag]
tag=${1:-ab}
mkdir -p gpurun_out
for mode in 1 0; do
  echo "=== NSFF_TC_SPLIT=$mode parity"
  NSFF_TC_SPLIT=$mode timeout 400 python -m pytest tests/test_gpu_parity.py -q -x --timeout=120 --tb=short 2>&1 | tail -12
  echo "=== NSFF_TC_SPLIT=$mode bench"
  NSFF_TC_SPLIT=$mode timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${tag}_split${mode}.err | tee gpurun_out/${tag}_split${mode}.json | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms/step', round(d['ms_per_step'],2), 'M rs/s', round(d['value']/1e6,1), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'W', d['clocks']['power_w_median'])
"
  tail -3 gpurun_out/${tag}_split${mode}.err
done
