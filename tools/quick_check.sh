#!/bin/bash
# Field-kernel iteration loop: forward parity, (optionally) backward/save-path parity, a short bench line and a timeline.
# usage: tools/quick_check.sh <tag> [full]
tag=${1:-q}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -q -x --timeout=120 --tb=short 2>&1 | tail -6
if [ "$2" = "full" ]; then timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_tile_gemm.py -q -x --timeout=200 --tb=short 2>&1 | tail -6; fi
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms/step', round(d['ms_per_step'],2), 'M rs/s', round(d['value']/1e6,1), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'W', d['clocks']['power_w_median'])
"
tail -3 gpurun_out/${tag}_bench.err
if [ -f nsff_pl_b200/libnsff_b200_tl.so ]; then NSFF_LIB=$PWD/nsff_pl_b200/libnsff_b200_tl.so timeout 120 python tools/timeline.py run gpurun_out/${tag}_tl; fi
