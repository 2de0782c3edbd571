"""GPU: error behaviour and drop-in details of the Python boundary (round-1 advisor findings)."""
import pytest
import torch

from parity_util import build_fields, synthetic_rays

pytestmark = pytest.mark.gpu


def test_interpolate_accepts_the_host_dicts_eval_py_builds():
    """eval.py:106-107 .cpu()'s every value before interpolate() and calls .numpy() on the outputs (eval.py:201-225)."""
    from nsff_pl_b200.models.rendering import interpolate, render_rays
    from scenes import FakeDataset
    w, h = 16, 9
    models, emb, _, _ = build_fields(W=256, seed=5, precision="fp32", n_frames=8)
    ds = FakeDataset(n_frames=8, n_cams=1, img_wh=(w, h), seed=1)
    rays, _ = synthetic_rays(w * h, n_frames=8, seed=6, img_wh=(w, h), frame=2)
    kw = dict(N_samples=12, perturb=0, noise_std=0, test_time=True, output_transient_flow=["fw", "bw"])
    with torch.no_grad():
        r0 = render_rays(models, emb, rays.cuda(), torch.full((w * h,), 2).cuda(), 7, **kw)
        r1 = render_rays(models, emb, rays.cuda(), torch.full((w * h,), 3).cuda(), 7, **kw)
    K, c2w = ds.Ks[0], torch.tensor(ds.poses[0])
    rgb_d, depth_d = interpolate(r0, r1, 0.3, K, c2w, (w, h))
    rgb_h, depth_h = interpolate({k: v.cpu() for k, v in r0.items()}, {k: v.cpu() for k, v in r1.items()}, 0.3, K, c2w, (w, h))
    assert rgb_h.device.type == "cpu" and depth_h.device.type == "cpu" and rgb_d.is_cuda
    rgb_h.numpy(), depth_h.view(h, w).numpy()                      # what eval.py does next
    assert torch.allclose(rgb_h, rgb_d.cpu(), atol=1e-5) and torch.allclose(depth_h, depth_d.cpu(), atol=1e-5)


def test_fp32_path_chunks_on_ray_boundaries():
    """B > 65536 points with S not dividing 65536 (static hierarchical pass: S = 128 + 64 = 192)."""
    from nsff_pl_b200.models.rendering import render_rays
    models, emb, _, _ = build_fields(W=256, transient=False, coarse=True, seed=7, precision="fp32")
    rays, _ = synthetic_rays(400, seed=8)
    with torch.no_grad():
        res = render_rays(models, emb, rays.cuda(), None, 0, N_samples=128, N_importance=64, test_time=True)
    assert res["zs_fine"].shape == (400, 192) and torch.isfinite(res["rgb_fine"]).all()


def test_time_code_table_is_validated_like_an_embedding():
    from nsff_pl_b200.models.rendering import render_rays
    models, emb, _, _ = build_fields(W=256, seed=9, precision="bf16", n_frames=8)
    rays, ts = synthetic_rays(8, n_frames=8, seed=10)
    kw = dict(N_samples=8, test_time=True, output_transient_flow=["fw", "bw"])
    with torch.no_grad():
        with pytest.raises(IndexError):
            render_rays(models, emb, rays.cuda(), ts.cuda(), 8, **kw)           # max_t beyond the table
        emb["t"].cpu()
        with pytest.raises(NotImplementedError):
            render_rays(models, emb, rays.cuda(), ts.cuda(), 7, **kw)           # table left on the host


def test_training_composite_rejects_more_than_256_samples_before_the_forward():
    from nsff_pl_b200.models.rendering import render_rays
    models, emb, _, _ = build_fields(W=256, seed=11, precision="bf16", n_frames=8)
    for m in models.values():
        m.requires_grad_(True)
    rays, ts = synthetic_rays(4, n_frames=8, seed=12)
    with pytest.raises(NotImplementedError, match="at most 256 samples"):
        render_rays(models, emb, rays.cuda(), ts.cuda(), 7, N_samples=264, test_time=False,
                    output_transient_flow=["fw", "bw", "disocc"])


def test_pos_embedding_refuses_to_drop_a_graph():
    from nsff_pl_b200.models.nerf import PosEmbedding
    x = torch.rand(4, 3, device="cuda", requires_grad=True)
    with pytest.raises(NotImplementedError):
        PosEmbedding(9, 10)(x)
    with torch.no_grad():
        assert PosEmbedding(9, 10)(x).shape == (4, 63)


def test_keys_selects_outputs_in_reference_order():
    """render_rays(keys=...) drops the per-sample tensors eval.f would otherwise .cpu()."""
    from nsff_pl_b200.models.rendering import PER_RAY_KEYS, render_rays
    models, emb, _, _ = build_fields(W=256, seed=13, precision="bf16", n_frames=8)
    rays, ts = synthetic_rays(16, n_frames=8, seed=14)
    kw = dict(N_samples=16, test_time=True, output_transient_flow=["fw", "bw"])
    with torch.no_grad():
        full = render_rays(models, emb, rays.cuda(), ts.cuda(), 7, **kw)
        some = render_rays(models, emb, rays.cuda(), ts.cuda(), 7, keys=PER_RAY_KEYS, **kw)
        with pytest.raises(KeyError):
            render_rays(models, emb, rays.cuda(), ts.cuda(), 7, keys=["rgb_fine", "rgb_fw"], **kw)   # no warp at test time
    assert list(some) == [k for k in full if k in PER_RAY_KEYS] and len(some) == len(PER_RAY_KEYS)
    assert all(torch.equal(some[k], full[k]) for k in some)
    assert sum(v.numel() for v in some.values()) * 8 < sum(v.numel() for v in full.values())     # (x 60 at 128 samples)
