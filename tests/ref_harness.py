"""Caller harness: drives a `render_rays` (ours or the reference's) the way the reference's two entry points do,
and builds the reference's NeRFWLoss with the buffers train.py gives it.  Test infrastructure.

pytorch_lightning / test_tube / imageio are not installed, so train.py and eval.py cannot be imported; these are
minimal re-statements of the two call loops (same argument order, same chunking, same concatenation and .cpu()
handling), written against the call sites:

  system_forward  <- NSFFSystem.forward        /root/reference/train.py:99-123
  eval_f          <- eval.f                    /root/reference/eval.py:81-110
  make_loss       <- NSFFSystem.__init__/setup /root/reference/train.py:41-44,135-138 (Ks, Ps buffers, max_t)
  make_batch      <- the training batch dict   /root/reference/datasets/monocular.py:240-249
"""
import copy
from collections import defaultdict

import numpy as np
import torch


def system_forward(render_rays, models, embeddings, rays, ts, n_frames, N_samples, perturb, noise_std, N_importance,
                   chunk, test_time=False, **kwargs):
    """Positional call of render_rays per `chunk` rays, per-key lists, torch.cat (values .cpu()'d at test time)."""
    out = defaultdict(list)
    for i in range(0, rays.shape[0], chunk):
        part = render_rays(models, embeddings, rays[i:i + chunk], None if ts is None else ts[i:i + chunk],
                           n_frames - 1, N_samples, 0 if test_time else perturb, 0 if test_time else noise_std,
                           N_importance, chunk // 4 if test_time else chunk, **dict(kwargs))
        for k, v in part.items():
            out[k].append(v.cpu() if test_time else v)
    return {k: torch.cat(v, 0) for k, v in out.items()}


@torch.no_grad()
def eval_f(render_rays, models, embeddings, rays, ts, max_t, N_samples, N_importance, chunk, **kwargs):
    """test_time=True, perturb = noise = 0, per-chunk view_dir slices, EVERY value moved to the host."""
    out = defaultdict(list)
    kw = copy.copy(kwargs)
    for i in range(0, rays.shape[0], chunk):
        if "view_dir" in kwargs:
            kw["view_dir"] = kwargs["view_dir"][i:i + chunk]
        part = render_rays(models, embeddings, rays[i:i + chunk], None if ts is None else ts[i:i + chunk], max_t,
                           N_samples, 0, 0, N_importance, chunk, test_time=True, **kw)
        for k, v in part.items():
            out[k].append(v.cpu())
    return {k: torch.cat(v, 0) for k, v in out.items()}


def make_cameras(n_frames, img_wh, seed=0):
    """Ks (1,3,3) and world-to-image projections Ps (1,N_frames,3,4) of a slowly moving pinhole camera."""
    rng = np.random.RandomState(seed)
    w, h = img_wh
    f = 0.9 * w
    K = np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1]], dtype=np.float32)
    Ps = []
    for t in range(n_frames):
        a = 0.02 * t + rng.uniform(-0.01, 0.01)
        Rm = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], dtype=np.float32)
        tr = np.array([[0.05 * t], [0.01 * rng.randn()], [0.02 * rng.randn()]], dtype=np.float32)
        # the reference's camera looks down -z (ray_utils.py:127-181 flips y,z when projecting): P = K [R | t] with
        # the sign flip folded in
        flip = np.diag([1, -1, -1]).astype(np.float32)
        Ps.append(K @ flip @ np.concatenate([Rm, tr], 1))
    return torch.tensor(K)[None], torch.tensor(np.stack(Ps))[None]


def make_loss(ref, n_frames, img_wh=(512, 288), device="cpu", seed=0, **loss_kwargs):
    """The reference's NeRFWLoss (losses.py:31-171), unmodified, with the buffers train.py:135-138 registers."""
    loss = ref.losses.NeRFWLoss(**loss_kwargs)
    Ks, Ps = make_cameras(n_frames, img_wh, seed)
    loss.register_buffer("Ks", Ks)
    loss.register_buffer("Ps", Ps)
    loss.max_t = n_frames - 1
    return loss.to(device)


def make_batch(rays, ts, img_wh=(512, 288), seed=0):
    """Targets of one training batch (monocular.py:240-249), random but fixed."""
    g = torch.Generator().manual_seed(seed)
    R = rays.shape[0]
    w, h = img_wh
    uv = torch.stack([torch.rand(R, generator=g) * w, torch.rand(R, generator=g) * h], -1)
    return {"rays": rays, "ts": ts, "rgbs": torch.rand(R, 3, generator=g), "cam_ids": torch.zeros(R, dtype=torch.long),
            "disps": torch.rand(R, generator=g) + 0.1, "rays_mask": torch.zeros(R),
            "uv_fw": uv + torch.randn(R, 2, generator=g), "uv_bw": uv - torch.randn(R, 2, generator=g)}


def to_device(batch, device):
    return {k: v.to(device) for k, v in batch.items()}
