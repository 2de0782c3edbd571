#!/usr/bin/env python
"""Which part of the training step survives CUDA-graph capture?  Captures growing prefixes of the step and reports."""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import build_fields, synthetic_rays  # noqa: E402
from nsff_pl_b200.models.rendering import render_rays  # noqa: E402
import ref_harness as H  # noqa: E402

dev = torch.device("cuda", 0)
N_FRAMES, IMG_WH, R, S = 30, (512, 288), 1024, 128
models, emb, _, _ = build_fields(W=256, seed=0, n_frames=N_FRAMES, device=dev, precision="bf16")
for m in models.values():
    m.requires_grad_(True)
emb["t"].weight.requires_grad_(True)
params = [p for m in models.values() for p in m.parameters()] + [emb["t"].weight]
opt = torch.optim.Adam(params, lr=5e-4, fused=True, capturable=True)
rays, ts = synthetic_rays(R, n_frames=N_FRAMES, seed=100)
rays, ts = rays.to(dev), ts.to(dev)
batch = H.to_device(H.make_batch(rays.cpu(), ts.cpu(), img_wh=IMG_WH, seed=7), dev)
Ks, Ps = H.make_cameras(N_FRAMES, IMG_WH, 0)
from nsff_pl_b200.losses import NeRFWLoss  # noqa: E402
loss_fn = NeRFWLoss()
loss_fn.register_buffer("Ks", Ks)
loss_fn.register_buffer("Ps", Ps)
loss_fn.max_t = N_FRAMES - 1
loss_fn = loss_fn.to(dev)
kw = dict(N_samples=S, perturb=1.0, noise_std=1.0, N_importance=0, test_time=False, output_transient_flow=["fw", "bw", "disocc"])


def step(upto):
    opt.zero_grad(set_to_none=True)
    res = render_rays(models, emb, rays, ts, N_FRAMES - 1, **kw)
    if upto == "forward":
        return res["rgb_fine"].detach().sum()
    if upto == "simple_backward":
        loss = sum(res[k].square().mean() for k in ("rgb_fine", "rgb_fw", "rgb_bw", "transient_flow_fw", "xyzs_fw_bw"))
        loss.backward()
        return loss.detach()
    loss = sum(loss_fn(res, batch, epoch=5, output_transient=True, output_transient_flow=["fw", "bw", "disocc"]).values())
    if upto == "loss":
        return loss.detach()
    loss.backward()
    if upto == "backward":
        return loss.detach()
    opt.step()
    return loss.detach()


for upto in sys.argv[1:] or ["forward", "simple_backward", "loss", "backward", "all"]:
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            step(upto)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            with torch.cuda.graph(g):
                out = step(upto)
        g.replay()
        torch.cuda.synchronize()
        print(f"capture up to {upto:16s}: OK  value {float(out):.5f}  warnings: {[str(x.message)[:80] for x in w][:2]}", flush=True)
    except Exception as e:
        print(f"capture up to {upto:16s}: FAILED {type(e).__name__}: {str(e)[:300]}", flush=True)
        break
