#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference's models/rendering.py imports kornia (absent) and cupy (absent,
via models/softsplat.py) at module import; neither is used by render_rays, so
two in-memory shims satisfy the import.  Nothing from the
reference is copied into this repo: only tensors it *computed* are stored.

Every fixture stores: the network state_dicts (reference parameter names), the
constructor flags, the inputs, the call kwargs, the torch seed used right
before the call, and every tensor of the returned dict.
"""
import os
import sys
import types
import json

import numpy as np
import torch

REF = os.environ.get("NSFF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from scenes import energise_, synthetic_rays, FakeDataset  # noqa: E402  (tests/scenes.py)


def import_reference():
    kornia = types.ModuleType("kornia")

    def create_meshgrid(h, w, normalized_coordinates=True, device="cpu"):
        xs = torch.linspace(0, w - 1, w, device=device)
        ys = torch.linspace(0, h - 1, h, device=device)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack([gx, gy], -1)[None]

    kornia.create_meshgrid = create_meshgrid
    cupy = types.ModuleType("cupy")
    cupy.memoize = lambda **kw: (lambda f: f)
    cupy.ndarray = type("ndarray", (), {})
    cupy.cuda = types.SimpleNamespace(compile_with_cache=None)
    sys.modules.setdefault("kornia", kornia)
    sys.modules.setdefault("cupy", cupy)
    sys.path.insert(0, REF)
    from models import nerf as ref_nerf          # noqa
    from models import rendering as ref_rendering  # noqa
    return ref_nerf, ref_rendering


def sd_np(module):
    return {k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def flags(m):
    return dict(typ=m.typ, D=m.D, W=m.W, skips=list(m.skips), in_channels_xyz=m.in_channels_xyz,
                use_viewdir=m.use_viewdir, in_channels_dir=m.in_channels_dir,
                encode_appearance=bool(m.encode_appearance), in_channels_a=m.in_channels_a,
                encode_transient=bool(m.encode_transient), in_channels_t=m.in_channels_t,
                output_flow=bool(m.output_flow), flow_scale=float(getattr(m, "flow_scale", 0.2)))


def save(name, meta, arrays):
    path = os.path.join(HERE, name + ".npz")
    arrays = {k: np.ascontiguousarray(v) for k, v in arrays.items()}
    arrays["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB, {len(arrays)} arrays)")


def pack_models(models):
    arrays, meta = {}, {}
    for key, m in models.items():
        meta[key] = flags(m)
        for k, v in sd_np(m).items():
            arrays[f"model.{key}.{k}"] = v
    return arrays, meta


def main():
    ref_nerf, ref_rendering = import_reference()
    torch.set_num_threads(4)
    N_FRAMES, N_TAU, N_A = 6, 48, 48

    def build(W, D=8, coarse=False, transient=True, appearance=False, viewdir=True, seed=0):
        torch.manual_seed(seed)
        models = {"fine": ref_nerf.NeRF("fine", D=D, W=W, skips=[4], in_channels_xyz=63, in_channels_dir=27,
                                        use_viewdir=viewdir,
                                        encode_appearance=appearance, in_channels_a=N_A,
                                        encode_transient=transient, in_channels_t=N_TAU,
                                        output_flow=transient, flow_scale=0.2)}
        if coarse:
            models["coarse"] = ref_nerf.NeRF("coarse", D=D, W=W, skips=[4], in_channels_xyz=63, in_channels_dir=27,
                                             use_viewdir=viewdir, encode_transient=transient, in_channels_t=N_TAU)
        emb = {"xyz": ref_nerf.PosEmbedding(9, 10), "dir": ref_nerf.PosEmbedding(3, 4)}
        if transient:
            emb["t"] = torch.nn.Embedding(N_FRAMES, N_TAU)
        if appearance:
            emb["a"] = torch.nn.Embedding(N_FRAMES, N_A)
        for m in models.values():
            energise_(m)
        return models, emb

    def run_case(name, models, emb, rays, ts, call_kwargs, seed, extra_arrays=None, dataset=None):
        arrays, meta_models = pack_models(models)
        for k in ("t", "a"):
            if k in emb:
                arrays[f"emb.{k}"] = emb[k].weight.detach().numpy().copy()
        arrays["rays"] = rays.numpy()
        if ts is not None:
            arrays["ts"] = ts.numpy()
        kw = dict(call_kwargs)
        tensor_kw = {}
        for k in ("view_dir", "t_embedded", "a_embedded"):
            if k in kw:
                tensor_kw[k] = kw[k]
                arrays[f"kw.{k}"] = kw[k].numpy()
                kw.pop(k)
        meta = dict(models=meta_models, call=kw, seed=seed, n_frames=N_FRAMES,
                    tensor_kwargs=sorted(tensor_kw), has_dataset=dataset is not None)
        torch.manual_seed(seed)
        with torch.no_grad():
            full_kw = dict(kw)
            full_kw.update(tensor_kw)
            if dataset is not None:
                full_kw["dataset"] = dataset
            max_t = full_kw.pop("max_t")
            out = ref_rendering.render_rays(models, emb, rays, ts, max_t, **full_kw)
        meta["result_keys"] = list(out.keys())
        for k, v in out.items():
            arrays[f"out.{k}"] = v.detach().numpy()
        if extra_arrays:
            arrays.update(extra_arrays)
        save(name, meta, arrays)

    # ---- 1. op-level: PosEmbedding, NeRF.forward modes, sample_pdf -----------
    models, emb = build(W=64, coarse=True)
    torch.manual_seed(3)
    x3 = torch.randn(37, 3)
    arrays, meta_models = pack_models(models)
    arrays["pe.x"] = x3.numpy()
    arrays["pe.xyz"] = emb["xyz"](x3).numpy()
    arrays["pe.dir"] = emb["dir"](x3).numpy()
    xin = torch.cat([emb["xyz"](x3), emb["dir"](x3 * 0.3), torch.randn(37, N_TAU)], 1)
    arrays["fwd.x"] = xin.numpy()
    with torch.no_grad():
        f = models["fine"]
        c = models["coarse"]
        arrays["fwd.fine.full_flow"] = f(xin, output_transient_flow=["fw", "bw", "disocc"]).numpy()
        arrays["fwd.fine.full_noflow"] = f(xin).numpy()
        arrays["fwd.fine.static"] = f(xin[:, :90], output_transient=False).numpy()
        arrays["fwd.fine.warp_fw"] = f(xin, output_static=False, output_transient_flow=["fw"]).numpy()
        arrays["fwd.fine.warp_bw"] = f(xin, output_static=False, output_transient_flow=["bw"]).numpy()
        xs = torch.cat([xin[:, :63], xin[:, 90:]], 1)
        arrays["fwd.coarse.sigma_both"] = c(xs, sigma_only=True).numpy()
        arrays["fwd.coarse.sigma_static"] = c(xin[:, :63], sigma_only=True, output_transient=False).numpy()
        arrays["fwd.coarse.full"] = c(xin).numpy()
    torch.manual_seed(5)
    bins = torch.sort(torch.rand(9, 15), -1)[0]
    wts = torch.rand(9, 14)
    wts[2] = 0
    wts[3, 4:9] = 0
    arrays["pdf.bins"], arrays["pdf.weights"] = bins.numpy(), wts.numpy()
    arrays["pdf.det"] = ref_rendering.sample_pdf(bins, wts, 11, det=True).numpy()
    torch.manual_seed(77)
    arrays["pdf.rand"] = ref_rendering.sample_pdf(bins, wts, 11, det=False).numpy()
    save("ops_w64", dict(models=meta_models, pdf_seed=77, n_importance=11), arrays)

    # ---- 2. render_rays mode matrix at W=64 -----------------------------------
    rays, ts = synthetic_rays(24, n_frames=N_FRAMES, seed=11)
    models, emb = build(W=64, seed=1)
    run_case("train_warp_w64", models, emb, rays, ts,
             dict(max_t=N_FRAMES - 1, N_samples=32, perturb=0, noise_std=0, N_importance=0, test_time=False,
                  output_transient_flow=["fw", "bw", "disocc"]), seed=100)
    run_case("train_warp_noisy_w64", models, emb, rays, ts,
             dict(max_t=N_FRAMES - 1, N_samples=32, perturb=1.0, noise_std=1.0, N_importance=0, test_time=False,
                  output_transient_flow=["fw", "bw", "disocc"]), seed=101)
    run_case("test_flow_w64", models, emb, rays, ts,
             dict(max_t=N_FRAMES - 1, N_samples=32, perturb=0, noise_std=0, N_importance=0, test_time=True,
                  output_transient_flow=["fw", "bw"]), seed=102)
    run_case("test_noflow_w64", models, emb, rays, ts,
             dict(max_t=N_FRAMES - 1, N_samples=32, perturb=0, noise_std=0, N_importance=0, test_time=True),
             seed=103)
    run_case("static_only_kw_w64", models, emb, rays, ts,
             dict(max_t=N_FRAMES - 1, N_samples=32, perturb=0, noise_std=0, N_importance=0, test_time=True,
                  output_transient=False), seed=104)
    torch.manual_seed(9)
    run_case("test_tensor_kwargs_w64", models, emb, rays, ts,
             dict(max_t=N_FRAMES - 1, N_samples=16, perturb=0, noise_std=0, N_importance=0, test_time=True,
                  output_transient_flow=["fw", "bw"], view_dir=torch.randn(24, 3) * 0.2,
                  t_embedded=torch.randn(24, N_TAU)), seed=105)

    # config 1: static-only NeRF (no transient branch at all), ts=None
    models_s, emb_s = build(W=64, transient=False, seed=2)
    run_case("cfg1_static_w64", models_s, emb_s, rays, None,
             dict(max_t=0, N_samples=64, perturb=0, noise_std=0, N_importance=0, test_time=True), seed=106)
    models_nv, emb_nv = build(W=64, transient=False, viewdir=False, seed=12)
    # use_viewdir=False still receives the dir columns (they are split off and dropped)
    run_case("static_noviewdir_w64", models_nv, emb_nv, rays, None,
             dict(max_t=0, N_samples=16, perturb=0, noise_std=0, N_importance=0, test_time=False), seed=113)

    # hierarchical
    models_h, emb_h = build(W=64, coarse=True, seed=3)
    run_case("hier_test_w64", models_h, emb_h, rays, ts,
             dict(max_t=N_FRAMES - 1, N_samples=16, perturb=0, noise_std=0, N_importance=16, test_time=True,
                  output_transient_flow=["fw", "bw"]), seed=107)
    run_case("hier_train_w64", models_h, emb_h, rays, ts,
             dict(max_t=N_FRAMES - 1, N_samples=16, perturb=1.0, noise_std=0.5, N_importance=8, test_time=False,
                  output_transient_flow=["fw", "bw", "disocc"]), seed=108)
    models_hs, emb_hs = build(W=64, coarse=True, transient=False, seed=4)
    run_case("hier_static_w64", models_hs, emb_hs, rays, None,
             dict(max_t=0, N_samples=16, perturb=0, noise_std=0, N_importance=16, test_time=True), seed=109)

    # appearance code
    models_a, emb_a = build(W=64, appearance=True, seed=5)
    run_case("appearance_train_w64", models_a, emb_a, rays, ts,
             dict(max_t=N_FRAMES - 1, N_samples=16, perturb=0, noise_std=0, N_importance=0, test_time=False,
                  output_transient_flow=["fw", "bw", "disocc"]), seed=110)

    # eval-only visibility mask (rendering.py:190-200)
    ds = FakeDataset(n_frames=N_FRAMES, n_cams=2, img_wh=(32, 18), seed=6)
    ts_same = torch.full_like(ts, 2)
    run_case("visibility_test_w64", models, emb, rays, ts_same,
             dict(max_t=N_FRAMES - 1, N_samples=32, perturb=0, noise_std=0, N_importance=0, test_time=True,
                  output_transient_flow=["fw", "bw"]), seed=111, dataset=ds,
             extra_arrays={"ds.K": ds.Ks[0].numpy(), "ds.poses": np.asarray(ds.poses, dtype=np.float32)})

    # ---- 2b. gradients: autograd of the reference through a fixed scalar loss over EVERY differentiable output ----
    def loss_of(res, seed):
        g = torch.Generator().manual_seed(seed)
        tot = 0.0
        for k, v in res.items():
            w = torch.randn(v.shape, generator=g)
            if v.requires_grad:
                tot = tot + (v * w).sum()
        return tot

    for name, kw, coarse in (
            ("grads_train_warp_w64", dict(N_samples=16, perturb=0, noise_std=0, N_importance=0, test_time=False,
                                          output_transient_flow=["fw", "bw", "disocc"]), False),
            ("grads_hier_train_w64", dict(N_samples=12, perturb=0, noise_std=0, N_importance=6, test_time=False,
                                          output_transient_flow=["fw", "bw", "disocc"]), True)):
        models_g, emb_g = build(W=64, coarse=coarse, seed=21)
        rays_g, ts_g = synthetic_rays(10, n_frames=N_FRAMES, seed=13)
        res = ref_rendering.render_rays(models_g, emb_g, rays_g, ts_g, N_FRAMES - 1, **kw)
        loss = loss_of(res, 77)
        loss.backward()
        arrays, meta_models = pack_models(models_g)
        arrays["emb.t"] = emb_g["t"].weight.detach().numpy().copy()
        arrays["rays"], arrays["ts"] = rays_g.numpy(), ts_g.numpy()
        for key, m in models_g.items():
            for n, p_ in m.named_parameters():
                if p_.grad is not None:
                    arrays[f"grad.{key}.{n}"] = p_.grad.numpy().copy()
        arrays["grad.emb.t"] = emb_g["t"].weight.grad.numpy().copy()
        save(name, dict(models=meta_models, call=dict(kw, max_t=N_FRAMES - 1), loss_seed=77, loss=float(loss),
                        n_frames=N_FRAMES, grad_keys=[k for k, v in res.items() if v.requires_grad],
                        result_keys=list(res.keys())), arrays)

    # ---- 2c. test-time interpolation (rendering.py:365-460).  The splat op is a cupy kernel that cannot run here: the
    # reference's interpolate() is executed with FunctionSoftsplat replaced by the oracle's CPU restatement of
    # softsplat.py:6-44 ('average') and Tensor.cuda() made a no-op, so the projection / blend / MPI composite are the
    # reference's own code and only the splat itself is restated.
    from oracle import nsff_oracle as O
    models_i, emb_i = build(W=64, seed=31)
    wi, hi = 12, 7
    rays_i, _ = synthetic_rays(wi * hi, n_frames=N_FRAMES, seed=14, img_wh=(wi, hi), frame=2)
    ts_i = torch.full((wi * hi,), 2, dtype=torch.long)
    kw_i = dict(N_samples=10, N_importance=0, test_time=True, output_transient_flow=["fw", "bw"])
    with torch.no_grad():
        rt = ref_rendering.render_rays(models_i, emb_i, rays_i, ts_i, N_FRAMES - 1, **kw_i)
        rp = ref_rendering.render_rays(models_i, emb_i, rays_i, ts_i + 1, N_FRAMES - 1, **kw_i)
    ds_i = FakeDataset(n_frames=N_FRAMES, n_cams=1, img_wh=(wi, hi), seed=8)
    K_i, c2w_i = ds_i.Ks[0], torch.tensor(ds_i.poses[1])
    saved = ref_rendering.FunctionSoftsplat, torch.Tensor.cuda
    ref_rendering.FunctionSoftsplat = lambda tenInput, tenFlow, tenMetric, strType: O.splat_average(tenInput, tenFlow)
    torch.Tensor.cuda = lambda self, *a_, **k_: self
    try:
        arrays = {}
        for dt in (0.0, 0.3, 0.75):
            rgb_i, depth_i = ref_rendering.interpolate(rt, rp, dt, K_i, c2w_i, (wi, hi))
            arrays[f"out.rgb.{dt}"], arrays[f"out.depth.{dt}"] = rgb_i.numpy(), depth_i.numpy()
    finally:
        ref_rendering.FunctionSoftsplat, torch.Tensor.cuda = saved
    keys = ("xyzs_fine", "zs_fine", "static_rgbs_fine", "static_alphas_fine", "transient_rgbs_fine",
            "transient_alphas_fine", "transient_flows_fw", "transient_flows_bw")
    for k in keys:
        arrays[f"t.{k}"], arrays[f"tp1.{k}"] = rt[k].numpy().copy(), rp[k].numpy().copy()
    arrays["K"], arrays["c2w"] = K_i.numpy(), c2w_i.numpy()
    save("interpolate_w64", dict(img_wh=[wi, hi], dts=[0.0, 0.3, 0.75], keys=list(keys)), arrays)

    # ---- 2d. test-frame ray generation (ray_utils.py:7-106 composed as monocular.py:268-276 does) ----------
    from datasets import ray_utils as ref_ru
    wr, hr = 16, 9
    ds_r = FakeDataset(n_frames=N_FRAMES, n_cams=1, img_wh=(wr, hr), seed=5)
    arrays = {"K": ds_r.Ks[0].numpy()}
    for pi in range(3):
        c2w_r = torch.tensor(ds_r.poses[pi]).clone()
        if pi == 2:
            c2w_r[2, 3] = -1.7                                  # camera behind z=-1: shift_near = 1.7 branch
        dirs = ref_ru.get_ray_directions(hr, wr, ds_r.Ks[0])
        ro, rd = ref_ru.get_rays(dirs, c2w_r)
        shift_near = -min(-1.0, float(c2w_r[2, 3]))
        ro, rd = ref_ru.get_ndc_rays(ds_r.Ks[0], 1.0, shift_near, ro, rd)
        arrays[f"c2w.{pi}"], arrays[f"rays.{pi}"] = c2w_r.numpy(), torch.cat([ro, rd], 1).numpy()
    save("frame_rays", dict(img_wh=[wr, hr], n=3, near=1.0), arrays)

    # ---- 3. full-width network (the shape the CUDA path is built for) ----------
    # weights are NOT stored (4.9 MB each); they are re-created from the seed and
    # guarded by a checksum so a different torch RNG is detected, not mis-compared.
    models_f, emb_f = build(W=256, seed=7)
    rays8, ts8 = synthetic_rays(8, n_frames=N_FRAMES, seed=12)
    torch.manual_seed(112)
    with torch.no_grad():
        out = ref_rendering.render_rays(models_f, emb_f, rays8, ts8, N_FRAMES - 1, N_samples=16, perturb=0,
                                        noise_std=0, N_importance=0, test_time=False,
                                        output_transient_flow=["fw", "bw", "disocc"])
    arrays = {"rays": rays8.numpy(), "ts": ts8.numpy()}
    csum = float(sum(v.double().abs().sum() for v in models_f["fine"].state_dict().values()))
    for k, v in out.items():
        arrays[f"out.{k}"] = v.numpy()
    save("train_warp_w256_seeded", dict(build_seed=7, weight_abs_sum=csum, n_frames=N_FRAMES,
                                        emb_abs_sum=float(emb_f["t"].weight.double().abs().sum()),
                                        call=dict(max_t=N_FRAMES - 1, N_samples=16, perturb=0, noise_std=0,
                                                  N_importance=0, test_time=False,
                                                  output_transient_flow=["fw", "bw", "disocc"]),
                                        result_keys=list(out.keys())), arrays)


if __name__ == "__main__":
    main()
