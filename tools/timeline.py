#!/usr/bin/env python
"""In-kernel timeline of field_forward_kernel (CTA 0), for finding what the tensor pipe waits on.

  here:   NSFF_LIB_OUT=$PWD/nsff_pl_b200/libnsff_b200_tl.so NSFF_NVCC_EXTRA=-DNSFF_TIMELINE python -m nsff_pl_b200.build
  gpu:    NSFF_LIB=$PWD/nsff_pl_b200/libnsff_b200_tl.so python tools/timeline.py run gpurun_out/tl    -> gpurun_out/tl.<launch>.bin
  here:   python tools/timeline.py show gpurun_out/tl.0.bin [first_sub last_sub]

Events are logged to shared memory (8 bytes: tag << 40 | clock) and dumped when each role ends; 512 events per role
(0 = producer, 1 = MMA issuer, 2 = epilogue warp 0).  tag = global sub-step index << 8 | slice << 4 | event.
"""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PER = 512
EV = {1: "prod: stage free, copy issued", 3: "mma: A ready", 4: "mma: weights ready, issue", 11: "mma: sub-step committed",
      5: "epi: wait acc", 6: "epi: acc complete", 7: "epi: hidden half published", 13: "epi: tile start",
      14: "mma: accumulators free", 15: "mma: prologue ready", 8: "epi: staging buffer free", 9: "epi: rows staged"}


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


def run_train(prefix):
    """one training forward + backward (1024 rays x 128): launches 0-2 = saving forwards (main, fw, bw), 3-5 = backward chains
    (bw, fw, main)"""
    import torch
    os.environ["NSFF_TIMELINE_OUT"] = prefix
    from parity_util import build_fields, synthetic_rays
    from nsff_pl_b200.models.rendering import render_rays
    models, emb, _, _ = build_fields(W=256, seed=0, n_frames=30)
    for m in models.values():
        m.requires_grad_(True)
    emb["t"].weight.requires_grad_(True)
    rays, ts = synthetic_rays(1024, n_frames=30, seed=1)
    res = render_rays(models, emb, rays.cuda(), ts.cuda(), 29, N_samples=128, test_time=False,
                      output_transient_flow=["fw", "bw", "disocc"])
    loss = sum(res[k].square().mean() for k in ("rgb_fine", "rgb_fw", "rgb_bw", "transient_flow_fw", "transient_flow_bw",
                                                 "xyzs_fw_bw", "xyzs_bw_fw"))
    loss.backward()
    torch.cuda.synchronize()
    print("wrote", prefix + ".*.bin")


def load(path):
    raw = open(path, "rb").read()
    vals = struct.unpack(f"<{len(raw) // 8}Q", raw)
    ev = []
    for role in range(3):
        for i in range(PER):
            v = vals[role * PER + i]
            if v == 0:
                break
            tag, clk = v >> 40, v & 0xFFFFFFFFFF
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
        print(f"{clk - t0:9d}  (+{d:6d})  sub {step:3d} sl {kc}  {EV.get(e, e)}")
    commits = [(clk, step) for clk, role, step, kc, e in ev if e == 11]
    print("\nsub-step: commit -> next commit")
    for (c0, a), (c1, b) in zip(commits, commits[1:]):
        print(f"  sub {a:3d}: {c1 - c0:6d}")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2])
    elif sys.argv[1] == "run_train":
        run_train(sys.argv[2])
    else:
        show(sys.argv[2], *(int(x) for x in sys.argv[3:5]))
