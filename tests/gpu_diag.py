"""Diagnostic (not a test): prints per-key max-abs error tables of the CUDA paths vs the oracle.
   python tests/gpu_diag.py [case ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from parity_util import build_fields, max_abs, psnr, synthetic_rays
from oracle import nsff_oracle as O
from nsff_pl_b200.models.rendering import render_rays


def table(tag, got, want):
    print(f"--- {tag}")
    for k, w in want.items():
        g = got[k]
        sc = float(w.abs().max()) if w.numel() else 0
        print(f"  {k:26s} shape {str(tuple(w.shape)):16s} |ref|max {sc:9.3e}  err {max_abs(g, w):9.3e}  finite {bool(torch.isfinite(g).all())}")


def run(models, emb, specs, enc, rays, ts, max_t, tag, **kw):
    for prec in ("fp32", "bf16"):
        for m in models.values():
            m.precision = prec
        kw_gpu = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
        got = render_rays(models, emb, rays.cuda(), None if ts is None else ts.cuda(), max_t, **kw_gpu)
        torch.cuda.synchronize()
        want = O.render_rays(specs, enc, rays, ts, max_t, **kw)
        table(f"{tag} [{prec}]", got, want)
        if "rgb_fine" in want:
            print(f"  PSNR rgb_fine {psnr(got['rgb_fine'], want['rgb_fine']):.1f} dB")


if __name__ == "__main__":
    which = sys.argv[1:] or ["warp", "hier", "default_init"]
    if "warp" in which:
        models, emb, specs, enc = build_fields(W=256, seed=5, n_frames=8)
        rays, ts = synthetic_rays(64, n_frames=8, seed=64)
        run(models, emb, specs, enc, rays, ts, 7, "train warp 64x128 energised", N_samples=128, test_time=False,
            output_transient_flow=["fw", "bw", "disocc"])
    if "hier" in which:
        models, emb, specs, enc = build_fields(W=256, coarse=True, seed=6, n_frames=8)
        rays, ts = synthetic_rays(24, n_frames=8, seed=3)
        run(models, emb, specs, enc, rays, ts, 7, "hier test 16+16", N_samples=16, N_importance=16, test_time=True,
            output_transient_flow=["fw", "bw"])
    if "default_init" in which:
        import scenes
        real = scenes.energise_
        import parity_util
        parity_util.energise_ = lambda m, **k: m
        models, emb, specs, enc = build_fields(W=256, seed=5, n_frames=8)
        parity_util.energise_ = real
        rays, ts = synthetic_rays(64, n_frames=8, seed=64)
        run(models, emb, specs, enc, rays, ts, 7, "train warp 64x128 DEFAULT INIT", N_samples=128, test_time=False,
            output_transient_flow=["fw", "bw", "disocc"])
