"""Diagnostic (not a test): bf16 error of the tensor-core path vs the INTRINSIC bf16 envelope of the reference
(the unmodified reference under torch.autocast('cpu', bfloat16) against itself in fp32), per result key and per
parameter gradient, on energised weights.   python tests/gpu_diag_envelope.py"""
import sys, os
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path[:0] = [HERE, os.path.dirname(HERE)]
import torch
from oracle import ref_loader
from parity_util import build_fields, synthetic_rays, max_abs
import ref_harness as H

ref = ref_loader.load()
torch.manual_seed(0)
KW = dict(N_samples=64, perturb=0, noise_std=0, N_importance=0, test_time=False, output_transient_flow=["fw", "bw", "disocc"])
for seed in (6, 16):
    models, emb, _, _ = build_fields(W=256, seed=seed, precision="bf16", n_frames=8)
    rmodels, remb = ref_loader.build_reference_fields(models, emb)
    rays, ts = synthetic_rays(96, n_frames=8, seed=seed + 1)
    from nsff_pl_b200.models.rendering import render_rays
    with torch.no_grad():
        want = ref.rendering.render_rays(rmodels, remb, rays, ts, 7, **KW)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            env = ref.rendering.render_rays(rmodels, remb, rays, ts, 7, **KW)
        got = render_rays(models, emb, rays.cuda(), ts.cuda(), 7, **KW)
    print(f"--- forward, seed {seed}: key, |ours-ref|, |autocast-ref|, ratio, scale")
    for k in want:
        e, v = max_abs(got[k], want[k]), max_abs(env[k].float(), want[k])
        print(f"{k:26s} {e:10.3e} {v:10.3e} {e / max(v, 1e-12):8.2f}  {float(want[k].abs().max()):8.3f}")

# gradients through the reference loss
for energise in (True, False):
    import parity_util
    real = parity_util.energise_
    if not energise:
        parity_util.energise_ = lambda m, **k: m
    models, emb, _, _ = build_fields(W=256, seed=31, precision="bf16", n_frames=8)
    parity_util.energise_ = real
    R, S = 40, 24
    rays, ts = synthetic_rays(R, n_frames=8, seed=32)
    batch = H.make_batch(rays, ts, img_wh=(64, 36), seed=33)
    kwargs = dict(epoch=4, output_transient=True, output_transient_flow=["fw", "bw", "disocc"])

    def ref_grads(autocast):
        rm, re = ref_loader.build_reference_fields(models, emb)
        for m in rm.values():
            m.requires_grad_(True)
        re["t"].weight.requires_grad_(True)
        ctx = torch.autocast("cpu", dtype=torch.bfloat16) if autocast else torch.autocast("cpu", enabled=False)
        with ctx:
            res = H.system_forward(ref.rendering.render_rays, rm, re, rays, ts, 8, S, 0, 0, 0, 64, **kwargs)
        res = {k: v.float() for k, v in res.items()}
        sum(H.make_loss(ref, 8, img_wh=(64, 36))(res, batch, **kwargs).values()).backward()
        return {f"{key}.{n}": p.grad.double() for key, m in rm.items() for n, p in m.named_parameters() if p.grad is not None}

    g32, g16 = ref_grads(False), ref_grads(True)
    for m in models.values():
        m.requires_grad_(True)
    emb["t"].weight.requires_grad_(True)
    res = H.system_forward(render_rays, models, emb, rays.cuda(), ts.cuda(), 8, S, 0, 0, 0, 64, **kwargs)
    sum(H.make_loss(ref, 8, img_wh=(64, 36), device="cuda")(res, H.to_device(batch, "cuda"), **kwargs).values()).backward()
    print(f"--- gradients (energised={energise}): param, rel err ours, rel err autocast, ratio")
    for key, m in models.items():
        for n, p in m.named_parameters():
            r = g32.get(f"{key}.{n}")
            if r is None or p.grad is None:
                continue
            eo = float((p.grad.cpu().double() - r).norm() / (r.norm() + 1e-12))
            ea = float((g16[f"{key}.{n}"] - r).norm() / (r.norm() + 1e-12))
            print(f"{n:36s} {eo:10.3e} {ea:10.3e} {eo / max(ea, 1e-12):8.2f}")
