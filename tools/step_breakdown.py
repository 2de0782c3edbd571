#!/usr/bin/env python
"""Per-kernel time of the last complete training step in an ncu launch list (gpu__time_duration.sum CSV)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for r in rows[1:]:
    v, u = float(r[vi].replace(",", "")), r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    seq.append((r[ki].split("(")[0][:80], ms))
idx = [i for i, (n, _) in enumerate(seq) if "sample_points" in n]
step = seq[idx[-2]:idx[-1]]
agg, cnt = collections.OrderedDict(), collections.Counter()
for n, ms in step:
    agg[n] = agg.get(n, 0) + ms
    cnt[n] += 1
tot = sum(agg.values())
print(f"kernels in one step: {len(step)}, total {tot:.3f} ms\n")
print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"| `{k}` | {cnt[k]} | {v:.3f} | {v / cnt[k] * 1e3:.1f} | {100 * v / tot:.1f}% |")
