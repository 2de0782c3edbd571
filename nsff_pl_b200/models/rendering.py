"""Drop-in `models/rendering.py`: render_rays() / sample_pdf() with the reference's signatures and result
dict (reference: models/rendering.py:10-49 and 52-362), executed by the sm_100a kernels behind
include/nsff_b200.h.  This file is host orchestration only: it validates inputs, allocates outputs,
draws the random tensors with torch (same draw order as the reference) and sequences kernel launches:

    nsff_sample_points -> nsff_field_forward (main pass) -> [nsff_field_forward x2 at the flow-warped
    points, t+1 / t-1] -> nsff_composite                      (+ nsff_sample_pdf / nsff_merge_sorted
                                                                 between the coarse and the fine pass)

`chunk` is accepted for signature compatibility and ignored: the kernels stream over all points.
"""
import ctypes as C

import torch

from .. import _cabi
from .nerf import NeRF

__all__ = ["render_rays", "sample_pdf", "interpolate"]

# Optional profiling hook (bench.py): when a list, every inference composite launch appends
# (start_event, end_event, algorithmic_hbm_bytes) recorded on the launching stream.
COMPOSITE_EVENTS = None
# per-ray outputs of the fine pass (what eval.py's image / depth / flow writers read)
PER_RAY_KEYS = ("rgb_fine", "depth_fine", "transient_alpha_fine", "transient_rgb_fine", "_static_rgb_fine",
                "_static_depth_fine", "xyz_fine", "transient_flow_fw", "transient_flow_bw")
_PER_SAMPLE_OUT = {"static_sigmas": 1, "transient_sigmas": 1, "static_weights": 1, "transient_weights": 1, "weights": 1,
                   "static_alphas": 1, "transient_alphas": 1, "xyzs_fw_bw": 3, "xyzs_bw_fw": 3, "disoccs_fw": 1,
                   "disoccs_bw": 1}


def _composite_bytes(a, B):
    """Algorithmic HBM bytes of one nsff_composite launch: every per-sample input read once, every per-sample
    output written once (per-ray outputs are ~0.2 % and ignored); SURVEY.md section 8d."""
    floats = a.raw_ld + 1                                   # raw row + z
    if a.has_flow:
        floats += 3                                         # xyz (for xyz_fine)
    for name in ("noise_static", "noise_trans", "noise_fw", "noise_bw", "vis"):
        floats += 1 if getattr(a, name) else 0
    if a.raw_fw:
        floats += 14 + (6 if (a.xyzs_fw_bw or a.xyzs_bw_fw) else 0)
    floats += sum(n for k, n in _PER_SAMPLE_OUT.items() if getattr(a, k))
    return 4 * B * floats


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5):
    """reference rendering.py:10-49.  bins (R, n+1), weights (R, n) -> (R, N_importance)."""
    if not (isinstance(bins, torch.Tensor) and isinstance(weights, torch.Tensor)):
        raise TypeError("bins and weights must be torch.Tensors")
    if weights.dim() != 2:
        raise ValueError(f"weights must be (N_rays, N_samples_), got {tuple(weights.shape)}")
    if weights.stride(1) != 1:
        weights = weights.contiguous()
    if bins.dim() == 2 and bins.stride(0) == 0:
        bins = bins[0]                                   # one row broadcast to every ray
    if bins.stride(-1) != 1:
        bins = bins.contiguous()
    for t, nm in ((bins, "bins"), (weights, "weights")):
        if not t.is_cuda:
            raise NotImplementedError(f"{nm} is a CPU tensor: nsff_pl_b200 kernels are CUDA-only (no CPU fallback)")
        if t.dtype != torch.float32:
            raise TypeError(f"{nm} must be float32")
    R, n = weights.shape
    if bins.shape[-1] != n + 1:
        raise ValueError(f"bins must have {n + 1} columns, got {tuple(bins.shape)}")
    u = None if det else torch.rand(R, N_importance, device=weights.device)
    out = torch.empty(R, N_importance, device=weights.device, dtype=torch.float32)
    _cabi.select_device(weights)
    bins_rs = 0 if bins.dim() == 1 else bins.stride(0)
    _cabi.check(_cabi.lib().nsff_sample_pdf(_cabi.ptr(bins), bins_rs, _cabi.ptr(weights), weights.stride(0), R, n,
                                            N_importance, _cabi.ptr(u), float(eps), _cabi.ptr(out),
                                            _cabi.stream_ptr(weights.device)), "nsff_sample_pdf")
    return out


def _require_octaves(embeddings):
    for key, n in (("xyz", 10), ("dir", 4)):
        emb = embeddings[key]
        freqs = torch.as_tensor(emb.freqs, dtype=torch.float32).cpu()
        if len(freqs) != n or not bool((freqs == 2.0 ** torch.arange(n)).all()):
            raise NotImplementedError(
                f"render_rays kernels are built for PosEmbedding({n - 1}, {n}) on '{key}' (reference defaults, "
                f"opt.py:25-28); got freqs={freqs.tolist()}")


def _field_pass(results, model, embeddings, xyz, zs, dirs, ts, max_t, want_transient, flow_heads, test_time,
                noise_std, t_embedded, a_embedded, kwargs):
    """The reference's `inference` closure (rendering.py:83-300) as kernel launches."""
    L = _cabi.lib()
    typ = model.typ
    R, S = zs.shape
    B = R * S
    dev = zs.device
    st = _cabi.stream_ptr(dev)
    f32 = dict(device=dev, dtype=torch.float32)
    results[f"zs_{typ}"] = zs
    results[f"xyzs_{typ}"] = xyz

    sigma_only = (typ == "coarse" and test_time)
    mask = NeRF.out_mask(sigma_only, True, want_transient, flow_heads)
    Cn = NeRF.out_channels(mask)
    raw = torch.empty(R, S, Cn, **f32)
    t_weight = None
    if want_transient and (t_embedded is None or (not test_time and bool(flow_heads))):   # (warp passes always gather)
        # the kernel gathers rows clamp(ts + shift, 0, max_t) of the time-code table itself: check what nn.Embedding
        # would have checked (device, dtype, layout, row count) instead of reading out of bounds on the device
        t_weight = _cabi.dev_tensor(embeddings["t"].weight.data, "embeddings['t'].weight")
        if t_weight.device != dev:
            raise ValueError(f"embeddings['t'].weight is on {t_weight.device}, the rays are on {dev}")
        if t_weight.dim() != 2 or t_weight.shape[1] != model.in_channels_t or t_weight.shape[0] <= int(max_t):
            raise IndexError(f"embeddings['t'].weight has shape {tuple(t_weight.shape)}; render_rays needs "
                             f"(> max_t = {int(max_t)}, {model.in_channels_t})")

    def points(xyz_in, offset=None, offset_ld=0, xyz_out=None, t_shift=0, use_t_embedded=True):
        p = _cabi.Points()
        p.xyz = xyz_in.data_ptr()
        p.offset = offset
        p.offset_ld = offset_ld
        p.xyz_out = None if xyz_out is None else xyz_out.data_ptr()
        p.dirs = dirs.data_ptr()
        p.zs = zs.data_ptr()
        p.B, p.S = B, S
        p.t_shift, p.max_t = t_shift, int(max_t)
        if want_transient:
            if use_t_embedded and t_embedded is not None:     # kwargs['t_embedded'] (R, N_tau): row r = ray r
                p.t_table, p.t_index = t_embedded.data_ptr(), None
            else:
                p.t_table, p.t_index = t_weight.data_ptr(), ts.data_ptr()
        if model.encode_appearance:
            p.a_table, p.a_index = a_embedded.data_ptr(), None
        return p

    # training-mode flow passes: the main pass also writes the warped points xyz + flow_fw / xyz + flow_bw
    # (rendering.py:219,225), which the two re-query passes then read as plain positions
    pre_warp = (want_transient and (not test_time) and bool(flow_heads) and "fw" in flow_heads and "bw" in flow_heads
                and model.precision == "bf16")
    main_pts = points(xyz)
    if pre_warp:
        xyz_fw, xyz_bw = torch.empty(R, S, 3, **f32), torch.empty(R, S, 3, **f32)
        main_pts.xyz_fw_out, main_pts.xyz_bw_out = xyz_fw.data_ptr(), xyz_bw.data_ptr()
    model.run_field(mask, raw.view(B, Cn), pts=main_pts)

    if not sigma_only:
        results[f"static_rgbs_{typ}"] = raw[..., :3]
        if want_transient:
            results[f"transient_rgbs_{typ}"] = raw[..., 4:7]
            if flow_heads:
                results["transient_flows_fw"] = raw[..., 8:11]      # already zeroed where zs > 0.95
                results["transient_flows_bw"] = raw[..., 11:14]

    a = _cabi.CompositeArgs()
    a.R, a.S = R, S
    a.zs, a.xyz, a.raw, a.raw_ld = zs.data_ptr(), xyz.data_ptr(), raw.data_ptr(), Cn
    a.has_static_rgb, a.has_transient, a.has_flow = int(not sigma_only), int(want_transient), int(bool(flow_heads))
    a.noise_std = float(noise_std)
    keep = []

    def noise():
        # drawn unconditionally, in the reference's order (rendering.py:207,213,128); only read if std > 0
        n = torch.randn(R, S, **f32)
        keep.append(n)
        return n.data_ptr() if noise_std > 0 else None

    def out(name, *shape):
        t = torch.empty(*shape, **f32)
        return t, t.data_ptr()

    # eval-only camera visibility (rendering.py:190-200)
    if test_time and want_transient and "dataset" in kwargs:
        ds = kwargs["dataset"]
        K = torch.as_tensor(ds.Ks[0], dtype=torch.float32).cpu()
        n_cams = len(ds.cam_train)
        t0 = int(ts[0])
        w2c = []
        for i in range(n_cams):
            c2w = torch.eye(4)
            c2w[:3] = torch.as_tensor(ds.poses[i * ds.N_frames + t0], dtype=torch.float32)
            w2c.append(torch.inverse(c2w)[:3].reshape(12))
        w2c = torch.stack(w2c).to(dev).contiguous()
        vis = torch.empty(R, S, **f32)
        Karr = (C.c_float * 4)(float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]))
        _cabi.check(L.nsff_visibility(_cabi.ptr(xyz), B, Karr, _cabi.ptr(w2c), n_cams, int(ds.img_wh[1]),
                                      int(ds.img_wh[0]), _cabi.ptr(vis), st), "nsff_visibility")
        a.vis = vis.data_ptr()
        keep += [vis, w2c]

    results[f"static_sigmas_{typ}"], a.static_sigmas = out("ss", R, S)
    a.noise_static = noise()
    warp = want_transient and (not test_time) and bool(flow_heads)
    if want_transient:
        results[f"transient_sigmas_{typ}"], a.transient_sigmas = out("ts", R, S)
        a.noise_trans = noise()
        if warp:                                                       # rendering.py:217-232
            warp_mask_fw = NeRF.out_mask(False, False, True, ["bw"])   # forward-warped points want the bw flow
            warp_mask_bw = NeRF.out_mask(False, False, True, ["fw"])
            raw_fw, raw_bw = torch.empty(R, S, 7, **f32), torch.empty(R, S, 7, **f32)
            if pre_warp:
                pts_fw = points(xyz_fw, t_shift=+1, use_t_embedded=False)
                pts_bw = points(xyz_bw, t_shift=-1, use_t_embedded=False)
            else:
                xyz_fw, xyz_bw = torch.empty(R, S, 3, **f32), torch.empty(R, S, 3, **f32)
                pts_fw = points(xyz, offset=raw.data_ptr() + 8 * 4, offset_ld=Cn, xyz_out=xyz_fw, t_shift=+1,
                                use_t_embedded=False)
                pts_bw = points(xyz, offset=raw.data_ptr() + 11 * 4, offset_ld=Cn, xyz_out=xyz_bw, t_shift=-1,
                                use_t_embedded=False)
            model.run_field(warp_mask_fw, raw_fw.view(B, 7), pts=pts_fw)
            a.noise_fw = noise()
            model.run_field(warp_mask_bw, raw_bw.view(B, 7), pts=pts_bw)
            a.noise_bw = noise()
            a.raw_fw, a.raw_bw, a.raw_w_ld = raw_fw.data_ptr(), raw_bw.data_ptr(), 7
            a.xyz_fw, a.xyz_bw = xyz_fw.data_ptr(), xyz_bw.data_ptr()
            results["xyzs_fw"] = xyz_fw
            results["rgb_fw"], a.rgb_fw = out("rgb_fw", R, 3)
            results["xyzs_bw"] = xyz_bw
            results["rgb_bw"], a.rgb_bw = out("rgb_bw", R, 3)
            results["xyzs_fw_bw"], a.xyzs_fw_bw = out("fwbw", R, S, 3)
            results["xyzs_bw_fw"], a.xyzs_bw_fw = out("bwfw", R, S, 3)
            keep += [raw_fw, raw_bw]

    results[f"static_weights_{typ}"], a.static_weights = out("sw", R, S)
    if want_transient:
        results[f"transient_weights_{typ}"], a.transient_weights = out("tw", R, S)
        results[f"weights_{typ}"], a.weights = out("w", R, S)
    full = not (test_time and typ == "coarse")
    if test_time and want_transient:
        results[f"static_alphas_{typ}"], a.static_alphas = out("sa", R, S)
        results[f"transient_alphas_{typ}"], a.transient_alphas = out("ta", R, S)
    if full:
        results[f"depth_{typ}"], a.depth = out("depth", R)
        if want_transient:
            results[f"rgb_{typ}"], a.rgb = out("rgb", R, 3)
            results[f"transient_alpha_{typ}"], a.transient_alpha = out("talpha", R)
            results[f"transient_rgb_{typ}"], a.transient_rgb = out("trgb", R, 3)
            results[f"_static_rgb_{typ}"], a.static_only_rgb = out("sorgb", R, 3)
            results[f"_static_depth_{typ}"], a.static_only_depth = out("sodepth", R)
            if flow_heads:
                results["xyz_fine"], a.xyz_fine = out("xyz_fine", R, 3)
                results["transient_flow_fw"], a.flow_fw = out("ffw", R, 3)
                results["xyz_fw"], a.xyz_fw_ray = out("xfw", R, 3)
                results["transient_flow_bw"], a.flow_bw = out("fbw", R, 3)
                results["xyz_bw"], a.xyz_bw_ray = out("xbw", R, 3)
                if warp and "disocc" in flow_heads:
                    results["disocc_fw"], a.disocc_fw = out("dfw", R, 1)
                    results["disoccs_fw"], a.disoccs_fw = out("dsfw", R, S, 1)
                    results["disocc_bw"], a.disocc_bw = out("dbw", R, 1)
                    results["disoccs_bw"], a.disoccs_bw = out("dsbw", R, S, 1)
        else:
            results[f"rgb_{typ}"], a.rgb = out("rgb", R, 3)
    ev = None
    if COMPOSITE_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record(torch.cuda.current_stream(dev))
    _cabi.check(L.nsff_composite(C.byref(a), st), "nsff_composite")
    if ev is not None:
        ev[1].record(torch.cuda.current_stream(dev))
        COMPOSITE_EVENTS.append((ev[0], ev[1], _composite_bytes(a, B)))
    del keep


def render_rays(models, embeddings, rays, ts, max_t, N_samples=64, perturb=0, noise_std=0, N_importance=0,
                chunk=1024 * 32, test_time=False, **kwargs):
    """Same inputs / outputs as reference rendering.py:52-81.

    models: {'fine'[, 'coarse']: NeRF}; embeddings: {'xyz','dir'[, 't','a']}; rays (R,6) fp32 CUDA;
    ts (R,) int64 or None.  kwargs honoured: output_transient, output_transient_flow, view_dir,
    t_embedded, a_embedded, dataset (eval-only visibility); everything else is ignored like the reference.

    Extension: `keys=[...]` returns only the named result keys, in the reference's order.
    eval.f / the validation loop .cpu() EVERY value of the dict (eval.py:106-107, train.py:119-121): ~1.7 GB per
    512x288x128 frame, almost all of it per-sample tensors the caller never reads; passing e.g.
    keys=PER_RAY_KEYS through eval.f's **kwargs leaves 0.1 % of that to copy.  Unknown names raise KeyError.
    """
    L = _cabi.lib()
    rays = _cabi.dev_tensor(rays if rays.is_contiguous() else rays.contiguous(), "rays")
    if rays.dim() != 2 or rays.shape[1] != 6:
        raise ValueError(f"rays must be (N_rays, 6), got {tuple(rays.shape)}")
    dev = rays.device
    R = rays.shape[0]
    _cabi.select_device(rays)
    _require_octaves(embeddings)
    for m in models.values():
        if not isinstance(m, NeRF):
            raise TypeError("models must be nsff_pl_b200.models.nerf.NeRF instances")
    # differentiable path: hand-written backward kernels behind autograd.Functions (models/autograd.py)
    grad_mode = torch.is_grad_enabled() and (
        any(p.requires_grad for m in models.values() for p in m.parameters()) or
        any(getattr(embeddings.get(k), "weight", None) is not None and embeddings[k].weight.requires_grad
            for k in ("t", "a")))
    if grad_mode and test_time:
        raise NotImplementedError("nsff_pl_b200: test_time=True is an inference mode (the reference calls it under "
                                  "torch.no_grad(), eval.py:81); use test_time=False to train")
    if ts is not None:
        ts = _cabi.dev_tensor(ts if ts.is_contiguous() else ts.contiguous(), "ts", torch.int64)
    st = _cabi.stream_ptr(dev)
    f32 = dict(device=dev, dtype=torch.float32)
    results = {}

    view_dir = kwargs.get("view_dir", None)
    if view_dir is None:
        dirs = rays[:, 3:6].contiguous()
    else:
        dirs = _cabi.dev_tensor(view_dir.contiguous(), "view_dir")

    def fixed(name):
        t = kwargs.get(name, None)
        return None if t is None else _cabi.dev_tensor(t.contiguous(), name)

    # stratified depths + points (rendering.py:313-324, 332)
    rand = torch.rand(R, N_samples, **f32) if perturb > 0 else None
    zs = torch.empty(R, N_samples, **f32)
    xyz = torch.empty(R, N_samples, 3, **f32)
    _cabi.check(L.nsff_sample_points(_cabi.ptr(rays), R, N_samples, float(perturb), _cabi.ptr(rand), _cabi.ptr(zs),
                                     _cabi.ptr(xyz), st), "nsff_sample_points")

    t_embedded = None
    want_transient = False
    if N_importance > 0:                                                # rendering.py:326-348
        model = models["coarse"]
        want_transient = bool(kwargs.get("output_transient", True) and model.encode_transient)
        if want_transient:
            t_embedded = fixed("t_embedded")
            if t_embedded is None and ts is None:
                raise KeyError("ts is required when the transient field is rendered")
        if grad_mode:
            from .autograd import field_pass_train
            field_pass_train(results, model, embeddings, xyz, zs, dirs, ts, max_t, want_transient, [], noise_std,
                             t_embedded, None)
        else:
            _field_pass(results, model, embeddings, xyz, zs, dirs, ts, max_t, want_transient, [], test_time,
                        noise_std, t_embedded, None, kwargs)
        # bins are the mid-points of the UN-jittered depths (rendering.py:315), one row for all rays
        lin = torch.linspace(0, 1, N_samples, **f32)
        zs_mid = (0.5 * (lin[:-1] + lin[1:])).contiguous()
        det = perturb == 0
        parts = [zs]
        w = results["static_weights_coarse"][:, 1:-1].detach()                 # rendering.py:336: no grad through the pdf
        zs_static = sample_pdf(zs_mid.expand(R, -1), w, N_importance, det=det)
        parts.append(zs_static)
        if test_time:
            results["static_zs_fine"] = zs_static
        if want_transient:
            zs_trans = sample_pdf(zs_mid.expand(R, -1), results["transient_weights_coarse"][:, 1:-1].detach(),
                                  N_importance, det=det)
            parts.append(zs_trans)
            if test_time:
                results["transient_zs_fine"] = zs_trans
        S_f = sum(p.shape[1] for p in parts)
        zs = torch.empty(R, S_f, **f32)
        pp = parts + [None] * (3 - len(parts))
        _cabi.check(L.nsff_merge_sorted(_cabi.ptr(pp[0]), pp[0].shape[1], _cabi.ptr(pp[1]), pp[1].shape[1],
                                        _cabi.ptr(pp[2]), 0 if pp[2] is None else pp[2].shape[1], R, _cabi.ptr(zs),
                                        st), "nsff_merge_sorted")
        xyz = torch.empty(R, S_f, 3, **f32)
        _cabi.check(L.nsff_points_from_depths(_cabi.ptr(rays), _cabi.ptr(zs), R, S_f, _cabi.ptr(xyz), st),
                    "nsff_points_from_depths")

    model = models["fine"]                                              # rendering.py:350-360
    a_embedded = None
    if model.encode_appearance:
        a_embedded = fixed("a_embedded")
        if a_embedded is None:
            a_embedded = torch.nn.functional.embedding(ts, embeddings["a"].weight).contiguous()
    if N_importance == 0:
        want_transient = bool(kwargs.get("output_transient", True) and model.encode_transient)
        if want_transient:
            t_embedded = fixed("t_embedded")
            if t_embedded is None and ts is None:
                raise KeyError("ts is required when the transient field is rendered")
    flow_heads = list(kwargs.get("output_transient_flow", [])) if want_transient else []
    if grad_mode:
        from .autograd import field_pass_train
        field_pass_train(results, model, embeddings, xyz, zs, dirs, ts, max_t, want_transient, flow_heads, noise_std,
                         t_embedded, a_embedded)
    else:
        _field_pass(results, model, embeddings, xyz, zs, dirs, ts, max_t, want_transient, flow_heads, test_time,
                    noise_std, t_embedded, a_embedded, kwargs)
    keys = kwargs.get("keys", None)
    if keys is not None:
        missing = [k for k in keys if k not in results]
        if missing:
            raise KeyError(f"render_rays(keys=...): {missing} not produced in this mode; available: {list(results)}")
        wanted = set(keys)
        results = {k: v for k, v in results.items() if k in wanted}
    return results


def _sample_view(t, name, S, cols=3):
    """(R,S,cols) view of a field output (row stride in floats) or a contiguous tensor -> (tensor, ld)."""
    if not t.is_cuda:
        raise NotImplementedError(f"{name} is a CPU tensor: nsff_pl_b200 kernels are CUDA-only (no CPU fallback)")
    if t.dtype != torch.float32 or t.dim() != 3 or t.shape[1] != S or t.shape[2] != cols:
        raise ValueError(f"{name} must be float32 (R, {S}, {cols}), got {t.dtype} {tuple(t.shape)}")
    if t.stride(2) != 1 or t.stride(0) != S * t.stride(1):
        t = t.contiguous()
    return t, t.stride(1)


def interpolate(results_t, results_tp1, dt, K, c2w, img_wh):
    """reference rendering.py:365-460.  results_t / results_tp1: render_rays(test_time=True,
    output_transient_flow=['fw','bw']) of the SAME rays at frames t and t+1; returns ((h,w,3) rgb, (h,w) depth) at
    time t+dt.  All S sample planes are splatted and composited on the device in two launches."""
    L = _cabi.lib()
    # eval.py:106-107 moves every value of both dicts to the host before calling interpolate() and then calls .numpy() on
    # the results (eval.py:201-225): host dicts are staged onto the current CUDA device and the outputs returned on the
    # device the inputs came from.  (The kernels themselves never run on the CPU.)
    host_inputs = not results_t["xyzs_fine"].is_cuda
    if host_inputs:
        if not torch.cuda.is_available():
            raise NotImplementedError("interpolate: nsff_pl_b200 kernels are CUDA-only (no CPU fallback)")
        dev_in = torch.device("cuda", torch.cuda.current_device())
        keys_t = ("xyzs_fine", "zs_fine", "static_rgbs_fine", "transient_rgbs_fine", "transient_flows_fw",
                  "static_alphas_fine", "transient_alphas_fine")
        keys_p = ("transient_rgbs_fine", "transient_flows_bw", "transient_alphas_fine")
        results_t = {k: results_t[k].to(dev_in, non_blocking=True) for k in keys_t}
        results_tp1 = {k: results_tp1[k].to(dev_in, non_blocking=True) for k in keys_p}
    xyzs = _cabi.dev_tensor(results_t["xyzs_fine"].contiguous(), "xyzs_fine")
    R, S = xyzs.shape[:2]
    w, h = int(img_wh[0]), int(img_wh[1])
    if w * h != R:
        raise ValueError(f"img_wh {img_wh} does not match {R} rays")
    dev = xyzs.device
    K = torch.as_tensor(K, dtype=torch.float32).cpu()
    c2w4 = torch.eye(4)
    c2w4[:3] = torch.as_tensor(c2w, dtype=torch.float32).cpu()
    w2c = torch.inverse(c2w4)[:3]
    w2c[1:] *= -1                                   # "right up back" -> "right down forward" (rendering.py:391-393)
    P = (K @ w2c).reshape(12)
    a = _cabi.InterpolateArgs()
    a.R, a.S, a.w, a.h, a.dt = R, S, w, h, float(dt)
    a.fx, a.fy, a.cx, a.cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    for i in range(12):
        a.P[i] = float(P[i])
    zs = _cabi.dev_tensor(results_t["zs_fine"].contiguous(), "zs_fine")
    s_rgb, ld_s = _sample_view(results_t["static_rgbs_fine"], "static_rgbs_fine", S)
    rgb_t, ld_rt = _sample_view(results_t["transient_rgbs_fine"], "transient_rgbs_fine", S)
    ffw, ld_f = _sample_view(results_t["transient_flows_fw"], "transient_flows_fw", S)
    rgb_p, ld_rp = _sample_view(results_tp1["transient_rgbs_fine"], "transient_rgbs_fine (t+1)", S)
    fbw, ld_b = _sample_view(results_tp1["transient_flows_bw"], "transient_flows_bw (t+1)", S)
    if not (ld_s == ld_rt == ld_f):                 # mixed layouts: fall back to packed copies of frame t
        s_rgb, rgb_t, ffw = s_rgb.contiguous(), rgb_t.contiguous(), ffw.contiguous()
        ld_s = ld_rt = ld_f = 3
    if ld_rp != ld_b:
        rgb_p, fbw = rgb_p.contiguous(), fbw.contiguous()
        ld_rp = ld_b = 3
    s_a = _cabi.dev_tensor(results_t["static_alphas_fine"].contiguous(), "static_alphas_fine")
    a_t = _cabi.dev_tensor(results_t["transient_alphas_fine"].contiguous(), "transient_alphas_fine")
    a_p = _cabi.dev_tensor(results_tp1["transient_alphas_fine"].contiguous(), "transient_alphas_fine (t+1)")
    a.xyzs, a.zs = xyzs.data_ptr(), zs.data_ptr()
    a.static_rgb, a.static_alpha = s_rgb.data_ptr(), s_a.data_ptr()
    a.rgb_t, a.alpha_t, a.flow_fw, a.ld_t = rgb_t.data_ptr(), a_t.data_ptr(), ffw.data_ptr(), ld_s
    a.rgb_tp1, a.alpha_tp1, a.flow_bw, a.ld_tp1 = rgb_p.data_ptr(), a_p.data_ptr(), fbw.data_ptr(), ld_rp
    scratch = torch.empty(L.nsff_interpolate_scratch_floats(R, S), device=dev, dtype=torch.float32)
    rgb = torch.empty(h, w, 3, device=dev, dtype=torch.float32)
    depth = torch.empty(h, w, device=dev, dtype=torch.float32)
    _cabi.select_device(xyzs)
    _cabi.check(L.nsff_interpolate(C.byref(a), _cabi.ptr(scratch), _cabi.ptr(rgb), _cabi.ptr(depth),
                                   _cabi.stream_ptr(dev)), "nsff_interpolate")
    if host_inputs:
        return rgb.cpu(), depth.cpu()
    return rgb, depth
