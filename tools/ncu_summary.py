#!/usr/bin/env python
"""Summarise ncu artefacts into profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py launches <launches.csv> <out.md>     # per-kernel time shares of one bench run
  python tools/ncu_summary.py full <prof.ncu-rep> <out.md>         # key metrics of the --set full capture
"""
import csv
import collections
import subprocess
import sys


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg, cnt = collections.OrderedDict(), collections.Counter()
    seq = []
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        agg[name] = agg.get(name, 0.0) + ms
        cnt[name] += 1
        seq.append((name, ms))
    tot = sum(agg.values())
    with open(out, "w") as fh:
        fh.write(f"# ncu launch list ({path})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over one bench.py run "
                 f"(cold-cache, serialised: compare SHARES).\n\ntotal kernel time {tot:.2f} ms over {len(seq)} launches\n\n"
                 "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            fh.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / tot:.1f}% |\n")
        fh.write("\n## one step, in launch order (last occurrence of the step pattern)\n\n| # | kernel | ms |\n|---|---|---:|\n")
        last = max(i for i, (n, _) in enumerate(seq) if "composite" in n)
        first = max(i for i, (n, _) in enumerate(seq[:last]) if "sample_points" in n)
        for i, (n, ms) in enumerate(seq[first:last + 1]):
            fh.write(f"| {i} | `{n}` | {ms:.3f} |\n")
    print(open(out).read())


KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full summary ({path})\n\n`ncu --set full --clock-control none --import-source on -k regex:<kernel>` "
                 "(see tools/profile_round.sh for the exact command)\n\n")
        names = [r[hdr.index("Kernel Name")].split("(")[0] for r in data]
        fh.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|---|" + "---:|" * len(data) + "\n")
        for k in KEYS:
            cols = [i for i, h in enumerate(hdr) if h == k or h.endswith("." + k)]
            if not cols:
                continue
            i = cols[0]
            fh.write(f"| `{k}` | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")
        fh.write(f"\nkernels: {names}\n")
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
