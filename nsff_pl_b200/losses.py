"""NeRFWLoss - drop-in for the reference's `losses.NeRFWLoss` (losses.py:31-171), SURVEY.md section 8f #2.

Same constructor, same attributes (`lambda_geo_d`, `lambda_geo_f`, `lambda_reg`, `lambda_ent`, `z_far`,
`thickness_filter`, `topk`), same buffers (`Ks`, `Ps` registered by the caller, `max_t` set by the caller:
train.py:135-138), same `forward(inputs, targets, **kwargs) -> dict` with the same keys in the same order.  The ~300
ATen kernels the reference's formulation costs (forward + autograd) become three launches forward and one backward
(csrc/nsff_loss.cu).

Differences, both deliberate:
  * `flow_fw_l` / `flow_bw_l` are always present when flows are on (value 0 if no ray has a valid target).  The
    reference drops the key in that case (losses.py:120-127), which it can only know by synchronising the device on
    `valid.any()`; `strict_keys=True` restores that behaviour (and the synchronisation).
  * hard example mining (`topk < 1`, losses.py:163-165) is not implemented: NotImplementedError.
"""
import ctypes as C

import torch
from torch import nn

from . import _cabi

TERM_KEYS = ("col_l", "disp_l", "entropy_l", "cross_entropy_l", "flow_fw_l", "flow_bw_l", "pho_l", "cyc_l",
             "reg_temp_sm_l", "reg_min_l", "reg_sp_sm_l")
_DIFF = ("rgb_fine", "depth_fine", "rgb_coarse", "depth_coarse", "tw", "sw", "xyz_fw", "xyz_bw", "rgb_fw", "rgb_bw",
         "xyzs_fw_bw", "xyzs_bw_fw", "xyzs_fw", "xyzs_bw")


def _f32(t, name):
    if t is None:
        return None
    t = _cabi.dev_tensor(t, name)
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


class _LossFn(torch.autograd.Function):
    """terms (11,) = f(differentiable render outputs); everything else rides in `ctx_in` (a dict of tensors/scalars)."""

    @staticmethod
    def forward(ctx, cfg, *diff):
        L = _cabi.lib()
        t = dict(zip(_DIFF, diff))
        t.update(cfg["const"])
        dev = t["rgb_fine"].device
        a = _cabi.LossArgs()
        a.R, a.S = cfg["R"], cfg["S"]
        for k in ("has_flow", "thickness", "n_reg", "n_frames", "max_t"):
            setattr(a, k, int(cfg[k]))
        for k in ("lambda_geo_d", "lambda_geo_f", "lambda_reg", "lambda_ent", "cross_entropy_w"):
            setattr(a, k, float(cfg[k]))
        stats = torch.empty(L.nsff_loss_stats_floats(), device=dev, dtype=torch.float32)
        sums = torch.empty(len(TERM_KEYS), device=dev, dtype=torch.float32)
        terms = torch.empty(len(TERM_KEYS), device=dev, dtype=torch.float32)
        t.update(stats=stats, term_sums=sums, terms=terms)
        cfg["const"]["stats"] = stats
        for name in _cabi.LossArgs._PTRS_IN:
            v = t.get(name)
            setattr(a, name, None if v is None else v.data_ptr())
        _cabi.select_device(terms)
        _cabi.check(L.nsff_nerfw_loss_forward(C.byref(a), _cabi.stream_ptr(dev)), "nsff_nerfw_loss_forward")
        ctx.cfg, ctx.tensors = cfg, t
        ctx.needs = [d is not None for d in diff]
        return terms

    @staticmethod
    def backward(ctx, g_terms):
        L = _cabi.lib()
        t, cfg = ctx.tensors, ctx.cfg
        dev = g_terms.device
        a = _cabi.LossArgs()
        a.R, a.S = cfg["R"], cfg["S"]
        for k in ("has_flow", "thickness", "n_reg", "n_frames", "max_t"):
            setattr(a, k, int(cfg[k]))
        for k in ("lambda_geo_d", "lambda_geo_f", "lambda_reg", "lambda_ent", "cross_entropy_w"):
            setattr(a, k, float(cfg[k]))
        g = g_terms.contiguous().float()
        t = dict(t, g_terms=g)
        for name in _cabi.LossArgs._PTRS_IN:
            v = t.get(name)
            setattr(a, name, None if v is None else v.data_ptr())
        grads = []
        for name, need in zip(_DIFF, ctx.needs):
            if not need or (not cfg["has_flow"] and name not in ("rgb_fine", "depth_fine", "rgb_coarse", "depth_coarse")):
                grads.append(None)
                continue
            gbuf = torch.empty_like(t[name])
            setattr(a, "d_" + name, gbuf.data_ptr())
            grads.append(gbuf)
        _cabi.select_device(g)
        _cabi.check(L.nsff_nerfw_loss_backward(C.byref(a), _cabi.stream_ptr(dev)), "nsff_nerfw_loss_backward")
        return (None, *grads)


class NeRFWLoss(nn.Module):
    """See the module docstring; term definitions and the paper references are in losses.py:31-56."""

    def __init__(self, lambda_geo=0.04, lambda_reg=0.1, thickness=1, topk=1.0, strict_keys=False):
        super().__init__()
        self.lambda_geo_d = self.lambda_geo_f = lambda_geo
        self.lambda_reg = lambda_reg
        self.lambda_ent = 1e-3
        self.z_far = 0.95
        self.thickness_filter = torch.ones(1, 1, max(thickness, 1))      # kept for attribute compatibility
        self.topk = topk
        self.strict_keys = strict_keys
        # self.Ks (n_cams,3,3), self.Ps (n_cams,N_frames,3,4) are registered as buffers and self.max_t is set by the
        # caller, exactly as train.py:135-138 does for the reference class

    def forward(self, inputs, targets, **kwargs):
        if self.topk < 1:
            raise NotImplementedError("nsff_pl_b200.losses.NeRFWLoss: hard example mining (topk < 1) is not implemented")
        has_flow = bool(kwargs["output_transient_flow"])
        rgb = _f32(inputs["rgb_fine"], "rgb_fine")
        R = rgb.shape[0]
        const = {"rgbs": _f32(targets["rgbs"], "rgbs"), "disps": _f32(targets["disps"], "disps")}
        diff = {"rgb_fine": rgb, "depth_fine": _f32(inputs["depth_fine"], "depth_fine"),
                "rgb_coarse": None, "depth_coarse": None}
        if "rgb_coarse" in inputs and "depth_coarse" in inputs:          # (losses.py:76-83: the coarse pair, x0.1 for colour)
            diff["rgb_coarse"] = _f32(inputs["rgb_coarse"], "rgb_coarse")
            diff["depth_coarse"] = _f32(inputs["depth_coarse"], "depth_coarse")
        S = 1
        cfg = dict(R=R, has_flow=has_flow, thickness=int(self.thickness_filter.shape[-1]), n_frames=1, max_t=0,
                   lambda_geo_d=self.lambda_geo_d, lambda_geo_f=self.lambda_geo_f, lambda_reg=self.lambda_reg,
                   lambda_ent=self.lambda_ent, cross_entropy_w=0.0)
        if R <= 2048:                       # the stats kernel finds the medians itself (bitonic sort in shared memory)
            const["medians"] = torch.empty(3, device=rgb.device, dtype=torch.float32)
        else:
            meds = [torch.median(diff["depth_fine"].detach()), torch.median(-const["disps"])]
            meds.append(torch.median(diff["depth_coarse"].detach()) if diff["depth_coarse"] is not None else meds[0])
            const["medians"] = torch.stack(meds).float()
        if has_flow:
            S = inputs["transient_weights_fine"].shape[1]
            cfg["cross_entropy_w"] = self.lambda_ent / 5 * min(kwargs["epoch"] / 10, 1.0)
            cfg["n_frames"], cfg["max_t"] = int(self.Ps.shape[1]), int(self.max_t)
            diff.update(tw=_f32(inputs["transient_weights_fine"], "transient_weights_fine"),
                        sw=_f32(inputs["static_weights_fine"], "static_weights_fine"),
                        xyz_fw=_f32(inputs["xyz_fw"], "xyz_fw"), xyz_bw=_f32(inputs["xyz_bw"], "xyz_bw"),
                        rgb_fw=_f32(inputs["rgb_fw"], "rgb_fw"), rgb_bw=_f32(inputs["rgb_bw"], "rgb_bw"),
                        xyzs_fw_bw=_f32(inputs["xyzs_fw_bw"], "xyzs_fw_bw"), xyzs_bw_fw=_f32(inputs["xyzs_bw_fw"], "xyzs_bw_fw"),
                        xyzs_fw=_f32(inputs["xyzs_fw"], "xyzs_fw"), xyzs_bw=_f32(inputs["xyzs_bw"], "xyzs_bw"))
            const.update(disocc_fw=_f32(inputs["disocc_fw"].detach(), "disocc_fw"),
                         disocc_bw=_f32(inputs["disocc_bw"].detach(), "disocc_bw"),
                         disoccs_fw=_f32(inputs["disoccs_fw"].detach(), "disoccs_fw"),
                         disoccs_bw=_f32(inputs["disoccs_bw"].detach(), "disoccs_bw"),
                         xyzs_fine=_f32(inputs["xyzs_fine"].detach(), "xyzs_fine"),
                         uv_fw=_f32(targets["uv_fw"], "uv_fw"), uv_bw=_f32(targets["uv_bw"], "uv_bw"),
                         cam_ids=_cabi.dev_tensor(targets["cam_ids"].long().contiguous(), "cam_ids", torch.int64),
                         ts=_cabi.dev_tensor(targets["ts"].long().contiguous(), "ts", torch.int64),
                         Ks=_f32(self.Ks, "Ks"), Ps=_f32(self.Ps, "Ps"))
        if "weights" in kwargs:
            const["ray_weights"] = _f32(kwargs["weights"].reshape(-1), "weights")
        cfg.update(S=int(S), n_reg=int(S * self.z_far), const=const)
        terms = _LossFn.apply(cfg, *[diff.get(k) for k in _DIFF])
        keys = TERM_KEYS if has_flow else TERM_KEYS[:2]
        parts = terms.unbind(0)            # one autograd node (backward: one stack) instead of eleven selects
        ret = {k: parts[TERM_KEYS.index(k)] for k in keys}
        if has_flow and self.strict_keys:
            nv = const_valid_counts(cfg)          # host synchronisation, like the reference's valid.any()
            for k, n in zip(("flow_fw_l", "flow_bw_l"), nv):
                if n == 0:
                    del ret[k]
        return ret


def const_valid_counts(cfg):
    """(n_valid_fw, n_valid_bw) of the last forward that used `cfg` - reads the statistics buffer (synchronises)."""
    st = cfg["const"]["stats"] if "stats" in cfg["const"] else None
    if st is None:
        raise RuntimeError("no forward has run")
    v = st[15:17].tolist()
    return int(v[0]), int(v[1])


loss_dict = {"nerfw": NeRFWLoss}
