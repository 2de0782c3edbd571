#!/bin/bash
# Round-end check on one GPU: the whole GPU test suite, smoke(), the default bench line.  usage: tools/final_check.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -q -m gpu --timeout=300 -x > gpurun_out/${tag}_gpu_tests.log 2>&1
grep -E "^E  .*Error|^E  .*assert|^FAILED|passed|failed" gpurun_out/${tag}_gpu_tests.log | head -20
timeout -k 5 300 python __graft_entry__.py smoke 2>&1 | grep -E "smoke|Error|error" | tail -6
timeout -k 5 900 python bench.py 2> gpurun_out/${tag}_bench.err > gpurun_out/${tag}_bench.json; echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/${tag}_bench.json"):
    if not l.startswith("{"): continue
    d = json.loads(l)
    r, c = d["roofline"], d["roofline_composite"]
    print("ms/step", round(d["ms_per_step"], 2), "M rs/s", round(d["value"] / 1e6, 1), "e2e", round(d["e2e"]["value"] / 1e6, 1), "frac", round(r["frac"], 3),
          "frac_exec", round(r.get("frac_executed", 0), 3), "composite", round(c["avg_launch_ms"], 3), "ms frac", round(c["frac"], 3),
          "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "parity", d.get("parity"))
    for k, v in (d.get("sub") or {}).items():
        if v: print(" sub", k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("ms_per_step", "value", "cuda_graph", "frac_of_bf16_peak", "ms_per_frame", "full_dict", "selected")})
PY
tail -3 gpurun_out/${tag}_bench.err
