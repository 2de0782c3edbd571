#!/usr/bin/env python
"""In-kernel timeline of field_forward_kernel (CTA 0), for finding what the tensor pipe waits on.

  here:   NSFF_NVCC_EXTRA=-DNSFF_TIMELINE python -m nsff_pl_b200.build     (then rebuild normally afterwards!)
  gpu:    python tools/timeline.py run gpurun_out/tl          -> gpurun_out/tl.<launch>.bin
  here:   python tools/timeline.py show gpurun_out/tl.0.bin [first_step last_step]
"""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PER = 8192
EV = {1: "prod: stage free, copy issued", 2: "mma: chunk top", 3: "mma: A chunk ready", 4: "mma: weights ready, issue",
      8: "mma: elected", 9: "mma: 1st issued", 10: "mma: all issued", 11: "mma: committed", 5: "epi: wait acc", 6: "epi: acc complete", 7: "epi: hidden chunk published", 12: "epi: prologue published",
      13: "epi: tile start", 14: "mma: acc slot free", 15: "mma: prologue ready"}


def run(prefix):
    import torch
    os.environ["NSFF_TIMELINE_OUT"] = prefix
    from parity_util import build_fields, synthetic_rays
    from nsff_pl_b200.models.rendering import render_rays
    models, emb, _, _ = build_fields(W=256, seed=0, n_frames=30)
    rays, ts = synthetic_rays(148 * 4, n_frames=30, seed=1)        # 4 tiles per SM at S=128
    with torch.no_grad():
        render_rays(models, emb, rays.cuda(), ts.cuda(), 29, N_samples=128, test_time=False,
                    output_transient_flow=["fw", "bw", "disocc"])
    torch.cuda.synchronize()
    print("wrote", prefix + ".*.bin")


def load(path):
    raw = open(path, "rb").read()
    vals = struct.unpack(f"<{len(raw) // 8}Q", raw)
    ev = []
    for role in range(3):
        base = role * 2 * PER
        for i in range(PER):
            tag, clk = vals[base + 2 * i], vals[base + 2 * i + 1]
            if clk == 0:
                break
            ev.append((clk, role, tag >> 8, (tag >> 4) & 15, tag & 15))
    ev.sort()
    return ev


def show(path, s0=0, s1=10**9):
    ev = load(path)
    t0 = ev[0][0]
    last = {}
    for clk, role, step, kc, e in ev:
        if not (s0 <= step <= s1):
            continue
        d = clk - last.get(role, clk)
        last[role] = clk
        print(f"{clk - t0:9d}  (+{d:6d})  step {step:3d} kc {kc}  {EV.get(e, e)}")
    # per-step summary: mma first issue -> next step first issue
    issues = [(clk, step, kc) for clk, role, step, kc, e in ev if e == 4]
    per_step = {}
    for clk, step, kc in issues:
        per_step.setdefault(step, []).append(clk)
    steps = sorted(per_step)
    print("\nstep: first-issue -> last-issue span, gap to next step's first issue")
    for a, b in zip(steps, steps[1:]):
        print(f"  step {a:3d}: chunks {len(per_step[a])}  span {per_step[a][-1] - per_step[a][0]:6d}  next-first - this-first {per_step[b][0] - per_step[a][0]:6d}")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2])
    else:
        show(sys.argv[2], *(int(x) for x in sys.argv[3:5]))
