"""Differentiable training path of render_rays: torch.autograd.Function wrappers around the CUDA kernels.

The reference gets its backward from torch autograd over ~25 nn.Linear layers and the einops/cumprod composite
(train.py:178-198 -> rendering.py:83-300).  Here the forward is the same pair of kernels as inference
(tcgen05 field kernel + composite) and the backward is hand-written:
    FieldFn      -> nsff_field_backward_tc (bf16 tcgen05: one chained dgrad launch per pass over the activations the
                    forward saved, weight-gradient GEMMs eight to a launch) or nsff_field_backward_fp32 (verification):
                    recompute + dgrad/wgrad per layer, PE and code gradients
    CompositeFn  -> nsff_composite_backward    (reverse transmittance scans)
Only cheap glue stays in torch (xyz + flow, embedding-table gathers), so their gradients come for free.
"""
import ctypes as C
import re

import torch

from .. import _cabi

_FIELD_OF = [  # parameter-name pattern -> (FieldGrads field stem, has layer index)
    (re.compile(r"^static_xyz_encoding_(\d+)\.0\.(weight|bias)$"), "static", True),
    (re.compile(r"^transient_xyz_encoding_(\d+)\.0\.(weight|bias)$"), "trans", True),
    (re.compile(r"^static_xyz_encoding_final\.(weight|bias)$"), "static_final", False),
    (re.compile(r"^transient_xyz_encoding_final\.(weight|bias)$"), "trans_final", False),
    (re.compile(r"^static_dir_encoding\.0\.(weight|bias)$"), "static_dir", False),
    (re.compile(r"^static_sigma\.(weight|bias)$"), "static_sigma", False),
    (re.compile(r"^static_rgb\.0\.(weight|bias)$"), "static_rgb", False),
    (re.compile(r"^transient_sigma\.(weight|bias)$"), "trans_sigma", False),
    (re.compile(r"^transient_rgb\.0\.(weight|bias)$"), "trans_rgb", False),
    (re.compile(r"^transient_flow_fw\.0\.(weight|bias)$"), "trans_fw", False),
    (re.compile(r"^transient_flow_bw\.0\.(weight|bias)$"), "trans_bw", False),
]


_SAVE_MAX_POINTS = 1 << 19     # 9.9 KB of saved activations per point and pass: 5 GB at this size


def _grad_slot(name):
    for rx, stem, indexed in _FIELD_OF:
        m = rx.match(name)
        if m:
            if indexed:
                return stem, int(m.group(1)) - 1, m.group(2)
            return stem, None, m.group(1)
    raise KeyError(f"parameter {name!r} has no slot in nsff_field_grads")


def _branch_active(stem, mask):
    if stem.startswith("static"):
        return bool(mask & _cabi.OUT_STATIC)
    if stem == "trans_fw":
        return bool(mask & _cabi.OUT_FLOW_FW)
    if stem == "trans_bw":
        return bool(mask & _cabi.OUT_FLOW_BW)
    return bool(mask & _cabi.OUT_TRANSIENT)


# ---- one gradient buffer per model and backward pass --------------------------------------------------------------------
# A training step evaluates the field three times (main pass, two warp passes); autograd would add the three nodes'
# parameter gradients with ~70 elementwise kernels.  The backward kernels accumulate (+=) into the buffers they are given,
# so the nodes of one backward pass share ONE zero-filled flat buffer: the first node that produces the gradient of a
# parameter returns the view, later nodes add into the same memory and return None for it.  The engine runs a leaf's
# AccumulateGrad only after every node that feeds it has run, so it sees the complete sum; the buffer is dropped by an
# engine callback when the pass ends.
_ARENAS = {}


def _arena(model, named, dev):
    task = torch._C._current_graph_task_id()          # the backward pass this node runs in
    a = _ARENAS.get(id(model))
    if a is None or a["task"] != task or task < 0:    # (a stale entry - a pass that ended in an exception - is replaced)
        sizes = [(p.numel() + 3) // 4 * 4 for _, p in named]              # 16-byte slots
        a = {"flat": torch.zeros(sum(sizes), device=dev, dtype=torch.float32), "sizes": sizes, "returned": set(),
             "task": task}
        _ARENAS[id(model)] = a
        if task >= 0:
            torch.autograd.Variable._execution_engine.queue_callback(
                lambda: _ARENAS.pop(id(model), None) if _ARENAS.get(id(model)) is a else None)
    return a


class _CodeRows(torch.autograd.Function):
    """weight[idx] for the per-ray code rows (rendering.py:220,226,331); the backward is one index_add instead of the
    sort-based dense embedding backward (38 us per pass on a 30-row table)."""

    @staticmethod
    def forward(ctx, weight, idx):
        ctx.save_for_backward(idx)
        ctx.shape = weight.shape
        return weight.index_select(0, idx)

    @staticmethod
    def backward(ctx, d_rows):
        (idx,) = ctx.saved_tensors
        g = torch.zeros(ctx.shape, device=d_rows.device, dtype=d_rows.dtype)
        g.index_add_(0, idx, d_rows)
        return g, None


class FieldFn(torch.autograd.Function):
    """raw (B,C) = NeRF heads at `pos` (B,3) for per-ray codes `t_code` / `a_code` (R,*)."""

    @staticmethod
    def forward(ctx, model, mask, S, pos, t_code, a_code, dirs, zs, *params):
        from .nerf import NeRF
        B = pos.shape[0]
        out = torch.empty(B, NeRF.out_channels(mask), device=pos.device, dtype=torch.float32)
        pts = _cabi.Points()
        pts.xyz, pts.dirs, pts.zs = pos.data_ptr(), dirs.data_ptr(), zs.data_ptr()
        pts.t_table = None if t_code is None else t_code.data_ptr()
        pts.a_table = None if a_code is None else a_code.data_ptr()
        pts.B, pts.S = B, S
        # tensor-core backward: let the forward kernel save every hidden activation (its shared-memory chunks already
        # are the backward's tile images) instead of recomputing them layer by layer
        save = None
        if (model.precision == "bf16" and model.backward_precision() == "bf16" and B <= _SAVE_MAX_POINTS
                and not getattr(model, "recompute_in_backward", False)):
            nbytes = _cabi.lib().nsff_field_save_bytes(C.byref(model.field_desc()), B, mask)
            if nbytes:
                save = torch.empty(nbytes, dtype=torch.uint8, device=pos.device)
        with torch.no_grad():
            model.run_field(mask, out, pts=pts, save=save, every_layer=True)
        ctx.saved_acts = save
        ctx.model, ctx.mask, ctx.S = model, mask, S
        ctx.save_for_backward(pos, t_code, a_code, dirs, zs, out)
        return out

    @staticmethod
    def backward(ctx, d_out):
        model, mask, S = ctx.model, ctx.mask, ctx.S
        pos, t_code, a_code, dirs, zs, out = ctx.saved_tensors
        L = _cabi.lib()
        dev = pos.device
        B = pos.shape[0]
        R = B // S
        d_out = d_out.contiguous()
        desc = model.field_desc()
        fp, keep = model.field_params()
        named = list(model.named_parameters())
        fg = _cabi.FieldGrads()
        active = []
        for name, p in named:
            stem, idx, kind = _grad_slot(name)
            ok = _branch_active(stem, mask) and not (stem == "static_dir" and not model.use_viewdir)
            active.append((stem, idx, kind, p) if ok else None)
        # the flat gradient buffer of this backward pass (see _arena): 16-byte slots in parameter order
        arena = _arena(model, named, dev)
        grads, o = [], 0
        for i, (a, n) in enumerate(zip(active, arena["sizes"])):
            if a is None:
                grads.append(None)
                o += n
                continue
            stem, idx, kind, p = a
            gbuf = arena["flat"][o:o + p.numel()].view_as(p)
            o += n
            if i in arena["returned"]:
                grads.append(None)                       # an earlier node of this pass returned the view: add in place
            else:
                arena["returned"].add(i)
                grads.append(gbuf)
            field = stem + ("_w" if kind == "weight" else "_b")
            if idx is None:
                setattr(fg, field, gbuf.data_ptr())
            else:
                getattr(fg, field)[idx] = gbuf.data_ptr()
        need_pos, need_t, need_a = ctx.needs_input_grad[3], ctx.needs_input_grad[4], ctx.needs_input_grad[5]
        need_t = need_t and t_code is not None and bool(mask & _cabi.OUT_TRANSIENT)
        need_a = need_a and a_code is not None and bool(mask & _cabi.OUT_STATIC)
        d_pos = torch.empty(B, 3, device=dev) if need_pos and (mask & _cabi.OUT_TRANSIENT) else None
        d_t = torch.empty(R, t_code.shape[1], device=dev) if need_t else None
        d_a = torch.empty(R, a_code.shape[1], device=dev) if need_a else None
        _cabi.select_device(pos)
        st = _cabi.stream_ptr(dev)
        tc = model.backward_precision() == "bf16"
        bwd_pack = model.field_bwd_pack() if tc else None
        saved = ctx.saved_acts if tc else None
        ctx.saved_acts = None        # consumed here: the node may outlive this call (e.g. a kept `loss`), the 10 KB/point must not
                                     # (a second backward through a retained graph falls back to the recompute path)
        rays_per = R if saved is not None else max(1, (262144 if tc else 32768) // S)
        ld = out.stride(0)
        for r0 in range(0, R, rays_per):
            nr = min(rays_per, R - r0)
            b0, nb = r0 * S, nr * S
            pts = _cabi.Points()
            pts.xyz = pos.data_ptr() + b0 * 12
            pts.dirs = dirs.data_ptr() + r0 * 12
            pts.zs = zs.data_ptr() + b0 * 4
            pts.t_table = None if t_code is None else t_code.data_ptr() + r0 * t_code.shape[1] * 4
            pts.a_table = None if a_code is None else a_code.data_ptr() + r0 * a_code.shape[1] * 4
            pts.B, pts.S = nb, S
            tail = (C.c_void_p(out.data_ptr() + b0 * ld * 4), C.c_void_p(d_out.data_ptr() + b0 * ld * 4), ld,
                    C.byref(fg),
                    None if d_pos is None else C.c_void_p(d_pos.data_ptr() + b0 * 12),
                    None if d_t is None else C.c_void_p(d_t.data_ptr() + r0 * d_t.shape[1] * 4),
                    None if d_a is None else C.c_void_p(d_a.data_ptr() + r0 * d_a.shape[1] * 4))
            if tc:
                scratch = torch.empty(L.nsff_field_bwd_tc_scratch_bytes(C.byref(desc), nb, int(saved is not None)),
                                      device=dev, dtype=torch.uint8)
                rc = L.nsff_field_backward_tc(C.byref(desc), C.byref(fp), _cabi.ptr(bwd_pack), C.byref(pts), mask,
                                              *tail, _cabi.ptr(saved), _cabi.ptr(scratch), st)
                _cabi.check(rc, "nsff_field_backward_tc")
            else:
                scratch = torch.empty(L.nsff_field_bwd_scratch_floats(C.byref(desc), nb), device=dev,
                                      dtype=torch.float32)
                rc = L.nsff_field_backward_fp32(C.byref(desc), C.byref(fp), C.byref(pts), mask, *tail,
                                                _cabi.ptr(scratch), st)
                _cabi.check(rc, "nsff_field_backward_fp32")
        del keep
        return (None, None, None, d_pos, d_t, d_a, None, None, *grads)


# output order of CompositeFn (subset chosen by the mode flags)
_SAMPLE_OUT = ["static_sigmas", "transient_sigmas", "static_weights", "transient_weights", "weights"]
_WARP_SAMPLE_OUT = ["xyzs_fw_bw", "xyzs_bw_fw"]
_RAY_OUT = ["depth", "rgb", "transient_alpha", "transient_rgb", "static_only_rgb", "static_only_depth"]
_FLOW_RAY_OUT = ["xyz_fine", "flow_fw", "xyz_fw_ray", "flow_bw", "xyz_bw_ray"]
_WARP_RAY_OUT = ["rgb_fw", "rgb_bw"]
_DISOCC_OUT = ["disocc_fw", "disoccs_fw", "disocc_bw", "disoccs_bw"]
_SHAPES = {"depth": (), "transient_alpha": (), "static_only_depth": (), "disocc_fw": (1,), "disocc_bw": (1,)}


def composite_outputs(has_flow, warp, disocc, transient=True):
    if not transient:
        return ["static_sigmas", "static_weights", "depth", "rgb"]
    names = list(_SAMPLE_OUT)
    if warp:
        names += _WARP_SAMPLE_OUT
    names += _RAY_OUT
    if has_flow:
        names += _FLOW_RAY_OUT
    if warp:
        names += _WARP_RAY_OUT
        if disocc:
            names += _DISOCC_OUT
    return names


def _composite_args(cfg, zs, xyz, noises, raw, raw_fw, raw_bw, xyz_fw, xyz_bw):
    R, S = zs.shape
    a = _cabi.CompositeArgs()
    a.R, a.S = R, S
    a.zs, a.xyz, a.raw, a.raw_ld = zs.data_ptr(), xyz.data_ptr(), raw.data_ptr(), raw.shape[-1]
    a.has_static_rgb, a.has_transient, a.has_flow = 1, int(cfg.get("transient", True)), int(cfg["has_flow"])
    a.noise_std = float(cfg["noise_std"])
    if cfg["noise_std"] > 0:
        a.noise_static = noises[0].data_ptr()
        if cfg.get("transient", True):
            a.noise_trans = noises[1].data_ptr()
    if cfg["warp"]:
        a.raw_fw, a.raw_bw, a.raw_w_ld = raw_fw.data_ptr(), raw_bw.data_ptr(), 7
        a.xyz_fw, a.xyz_bw = xyz_fw.data_ptr(), xyz_bw.data_ptr()
        if cfg["noise_std"] > 0:
            a.noise_fw, a.noise_bw = noises[2].data_ptr(), noises[3].data_ptr()
    return a


MAX_TRAIN_SAMPLES = 256      # nsff_composite_bwd.cu: kMaxChunksBwd (8) x 32 lanes


class CompositeFn(torch.autograd.Function):
    """All outputs of rendering.py:202-298 for the static+transient training composite."""

    @staticmethod
    def forward(ctx, cfg, zs, xyz, n_s, n_t, n_fw, n_bw, raw, raw_fw, raw_bw, xyz_fw, xyz_bw):
        ctx.set_materialize_grads(False)      # undefined output gradients stay None (NULL for the kernel), no zero fills
        R, S = zs.shape
        if S > MAX_TRAIN_SAMPLES:
            # nsff_composite_backward keeps a ray's samples in registers (32 lanes x 8 chunks): say so before the forward
            # instead of failing inside loss.backward()
            raise NotImplementedError(
                f"nsff_pl_b200: training composites support at most {MAX_TRAIN_SAMPLES} samples per ray, got {S} "
                f"(N_samples + 2*N_importance with a transient coarse model); render under torch.no_grad() or reduce "
                f"N_samples / N_importance")
        dev = zs.device
        names = composite_outputs(cfg["has_flow"], cfg["warp"], cfg["disocc"], cfg.get("transient", True))
        a = _composite_args(cfg, zs, xyz, (n_s, n_t, n_fw, n_bw), raw, raw_fw, raw_bw, xyz_fw, xyz_bw)
        outs = []
        for nm in names:
            if nm in _SAMPLE_OUT:
                shape = (R, S)
            elif nm in _WARP_SAMPLE_OUT:
                shape = (R, S, 3)
            elif nm in ("disoccs_fw", "disoccs_bw"):
                shape = (R, S, 1)
            else:
                shape = (R,) + _SHAPES.get(nm, (3,))
            t = torch.empty(*shape, device=dev, dtype=torch.float32)
            setattr(a, nm, t.data_ptr())
            outs.append(t)
        _cabi.select_device(zs)
        _cabi.check(_cabi.lib().nsff_composite(C.byref(a), _cabi.stream_ptr(dev)), "nsff_composite")
        ctx.cfg, ctx.names = cfg, names
        ctx.save_for_backward(zs, xyz, n_s, n_t, n_fw, n_bw, raw, raw_fw, raw_bw, xyz_fw, xyz_bw)
        ctx.mark_non_differentiable(*[t for nm, t in zip(names, outs) if nm in _DISOCC_OUT])   # rendering.py:290-291
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        cfg, names = ctx.cfg, ctx.names
        zs, xyz, n_s, n_t, n_fw, n_bw, raw, raw_fw, raw_bw, xyz_fw, xyz_bw = ctx.saved_tensors
        dev = zs.device
        a = _composite_args(cfg, zs, xyz, (n_s, n_t, n_fw, n_bw), raw, raw_fw, raw_bw, xyz_fw, xyz_bw)
        g = _cabi.CompositeGrads()
        keep = {}
        for nm, go in zip(names, gouts):
            if go is None or nm in _DISOCC_OUT:
                continue
            keep[nm] = go.contiguous()
            setattr(g, nm, keep[nm].data_ptr())
        d_raw = torch.empty_like(raw)
        g.d_raw = d_raw.data_ptr()
        d_raw_fw = d_raw_bw = None
        if cfg["warp"]:
            d_raw_fw, d_raw_bw = torch.empty_like(raw_fw), torch.empty_like(raw_bw)
            g.d_raw_fw, g.d_raw_bw = d_raw_fw.data_ptr(), d_raw_bw.data_ptr()
        _cabi.select_device(zs)
        _cabi.check(_cabi.lib().nsff_composite_backward(C.byref(a), C.byref(g), _cabi.stream_ptr(dev)),
                    "nsff_composite_backward")
        return (None, None, None, None, None, None, None, d_raw, d_raw_fw, d_raw_bw,
                keep.get("xyzs_fw_bw"), keep.get("xyzs_bw_fw"))


def field_pass_train(results, model, embeddings, xyz, zs, dirs, ts, max_t, want_transient, flow_heads, noise_std,
                     t_embedded, a_embedded):
    """Differentiable counterpart of rendering._field_pass for test_time=False (same keys, same order)."""
    from .nerf import NeRF
    typ = model.typ
    R, S = zs.shape
    B = R * S
    dev = zs.device
    f32 = dict(device=dev, dtype=torch.float32)
    params = list(model.parameters())
    results[f"zs_{typ}"] = zs
    results[f"xyzs_{typ}"] = xyz
    zs_flat = zs.reshape(B)

    if not want_transient:                                                 # static-only field (rendering.py:297-298)
        mask = NeRF.out_mask(False, True, False, [])
        raw = FieldFn.apply(model, mask, S, xyz.reshape(B, 3), None, a_embedded, dirs, zs_flat, *params).view(R, S, 4)
        results[f"static_rgbs_{typ}"] = raw[..., :3]
        n_s = torch.randn(R, S, **f32)
        cfg = dict(has_flow=False, warp=False, disocc=False, noise_std=float(noise_std), transient=False)
        o = dict(zip(composite_outputs(False, False, False, False),
                     CompositeFn.apply(cfg, zs, xyz, n_s, None, None, None, raw.contiguous(), None, None, None, None)))
        results[f"static_sigmas_{typ}"] = o["static_sigmas"]
        results[f"static_weights_{typ}"] = o["static_weights"]
        results[f"depth_{typ}"] = o["depth"]
        results[f"rgb_{typ}"] = o["rgb"]
        return

    def t_rows(shift):
        if shift == 0 and t_embedded is not None:
            return t_embedded
        idx = ts if shift == 0 else (torch.clamp(ts + 1, max=max_t) if shift > 0 else torch.clamp(ts - 1, min=0))
        return _CodeRows.apply(embeddings["t"].weight, idx)                                 # rendering.py:220,226,331

    mask = NeRF.out_mask(False, True, True, flow_heads)
    Cn = NeRF.out_channels(mask)
    raw = FieldFn.apply(model, mask, S, xyz.reshape(B, 3), t_rows(0), a_embedded, dirs, zs_flat, *params).view(R, S, Cn)
    results[f"static_rgbs_{typ}"] = raw[..., :3]
    results[f"transient_rgbs_{typ}"] = raw[..., 4:7]
    if flow_heads:
        results["transient_flows_fw"] = raw[..., 8:11]
        results["transient_flows_bw"] = raw[..., 11:14]

    def noise():
        return torch.randn(R, S, **f32)       # drawn in the reference's order (rendering.py:207,213,128)

    n_s, n_t = noise(), noise()
    warp = bool(flow_heads)
    raw_fw = raw_bw = xyz_fw = xyz_bw = n_fw = n_bw = None
    if warp:                                                               # rendering.py:217-232
        xyz_fw = xyz + raw[..., 8:11]
        raw_fw = FieldFn.apply(model, NeRF.out_mask(False, False, True, ["bw"]), S, xyz_fw.reshape(B, 3), t_rows(+1),
                               a_embedded, dirs, zs_flat, *params).view(R, S, 7)
        n_fw = noise()
        xyz_bw = xyz + raw[..., 11:14]
        raw_bw = FieldFn.apply(model, NeRF.out_mask(False, False, True, ["fw"]), S, xyz_bw.reshape(B, 3), t_rows(-1),
                               a_embedded, dirs, zs_flat, *params).view(R, S, 7)
        n_bw = noise()
    cfg = dict(has_flow=bool(flow_heads), warp=warp, disocc=warp and "disocc" in flow_heads, noise_std=float(noise_std))
    outs = dict(zip(composite_outputs(cfg["has_flow"], cfg["warp"], cfg["disocc"]),
                    CompositeFn.apply(cfg, zs, xyz, n_s, n_t, n_fw, n_bw, raw.contiguous(),
                                      None if raw_fw is None else raw_fw.contiguous(),
                                      None if raw_bw is None else raw_bw.contiguous(),
                                      None if xyz_fw is None else xyz_fw.contiguous(),
                                      None if xyz_bw is None else xyz_bw.contiguous())))
    # ---- same key order as rendering.py:206-295 ----
    results[f"static_sigmas_{typ}"] = outs["static_sigmas"]
    results[f"transient_sigmas_{typ}"] = outs["transient_sigmas"]
    if warp:
        results["xyzs_fw"] = xyz_fw
        results["rgb_fw"] = outs["rgb_fw"]
        results["xyzs_bw"] = xyz_bw
        results["rgb_bw"] = outs["rgb_bw"]
        results["xyzs_fw_bw"] = outs["xyzs_fw_bw"]
        results["xyzs_bw_fw"] = outs["xyzs_bw_fw"]
    results[f"static_weights_{typ}"] = outs["static_weights"]
    results[f"transient_weights_{typ}"] = outs["transient_weights"]
    results[f"weights_{typ}"] = outs["weights"]
    results[f"depth_{typ}"] = outs["depth"]
    results[f"rgb_{typ}"] = outs["rgb"]
    results[f"transient_alpha_{typ}"] = outs["transient_alpha"]
    results[f"transient_rgb_{typ}"] = outs["transient_rgb"]
    results[f"_static_rgb_{typ}"] = outs["static_only_rgb"]
    results[f"_static_depth_{typ}"] = outs["static_only_depth"]
    if flow_heads:
        results["xyz_fine"] = outs["xyz_fine"]
        results["transient_flow_fw"] = outs["flow_fw"]
        results["xyz_fw"] = outs["xyz_fw_ray"]
        results["transient_flow_bw"] = outs["flow_bw"]
        results["xyz_bw"] = outs["xyz_bw_ray"]
        if cfg["disocc"]:
            for k in _DISOCC_OUT:
                results[k] = outs[k]
