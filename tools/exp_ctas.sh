for c in 148 111 74 37; do echo "== NSFF_MAX_CTAS=$c"; NSFF_MAX_CTAS=$c timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step',d['ms_per_step'],'value',d['value'],'frac',d['roofline']['frac'])
"; done
