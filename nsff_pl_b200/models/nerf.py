"""Drop-in `models/nerf.py`: PosEmbedding and NeRF with the reference's constructor arguments, attribute
names, state_dict keys and forward() signatures (reference: models/nerf.py:4-30 and 33-213), whose
forward passes run on the hand-written sm_100a kernels behind include/nsff_b200.h.

Parameters stay ordinary fp32 ``nn.Parameter``s (so checkpoints, optimizers and DDP see exactly the
reference's tensors); the tensor-core path keeps a lazily refreshed bf16 "field pack" next to them.
"""
import ctypes as C

import torch
from torch import nn

from .. import _cabi

__all__ = ["PosEmbedding", "NeRF"]

# Optional profiling hook (bench.py): when set to a list, every tensor-core field launch appends
# (start_event, end_event, algorithmic_flops) recorded on the launching stream.
FIELD_EVENTS = None

# multiply-accumulates per point (exact from the layer shapes of reference nerf.py:74-116)
_MAC_STATIC, _MAC_STATIC_SIGMA_ONLY = 630_016, 491_264
_MAC_TRANS_BASE = 581_376          # trunk + final + sigma head; +768 per 3-wide head (rgb, fw, bw)


class PosEmbedding(nn.Module):
    """x -> [x, sin(f0 x), cos(f0 x), ...]; same ctor/attrs as reference nerf.py:4-15."""

    def __init__(self, max_logscale, N_freqs, logscale=True):
        super().__init__()
        self.funcs = [torch.sin, torch.cos]
        exps = torch.linspace(0, max_logscale, N_freqs)
        self.freqs = 2 ** exps if logscale else torch.linspace(1, 2 ** max_logscale, N_freqs)
        self._freqs_dev = {}

    def _freqs_on(self, device):
        key = (device.type, device.index)
        if key not in self._freqs_dev:
            self._freqs_dev[key] = self.freqs.to(device=device, dtype=torch.float32).contiguous()
        return self._freqs_dev[key]

    def is_octaves(self, n):
        """True iff freqs == 2**arange(n) (what the fused renderer kernels hard-wire)."""
        return len(self.freqs) == n and bool((self.freqs == 2.0 ** torch.arange(n)).all())

    def forward(self, x):
        """x: (B, 3) CUDA fp32 -> (B, 6*N_freqs+3)   (reference nerf.py:17-30)."""
        if isinstance(x, torch.Tensor) and x.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError(
                "nsff_pl_b200: PosEmbedding.forward records no autograd graph - call it under torch.no_grad(); the "
                "differentiable path is render_rays (models/autograd.py), which encodes positions inside the field kernel.")
        x = _cabi.dev_tensor(x.contiguous() if isinstance(x, torch.Tensor) else x, "x")
        if x.dim() != 2 or x.shape[1] != 3:
            raise ValueError(f"PosEmbedding expects (B, 3), got {tuple(x.shape)}")
        nf = len(self.freqs)
        out = torch.empty(x.shape[0], 3 + 6 * nf, device=x.device, dtype=torch.float32)
        _cabi.select_device(x)
        _cabi.check(_cabi.lib().nsff_pos_embedding(_cabi.ptr(x), x.shape[0], _cabi.ptr(self._freqs_on(x.device)), nf,
                                                   _cabi.ptr(out), _cabi.stream_ptr(x.device)), "nsff_pos_embedding")
        return out


def _act_linear(n_in, n_out, act):
    return nn.Sequential(nn.Linear(n_in, n_out), act)


class NeRF(nn.Module):
    """Static (+ optional dynamic/"transient") 8x256 field.  Constructor arguments, attributes read by the
    renderer (.typ, .encode_appearance, .encode_transient, ...) and parameter names follow reference
    nerf.py:34-116.  Extra attribute: ``precision`` in {'bf16' (tcgen05 tensor cores, default), 'fp32'
    (CUDA-core verification path)}."""

    def __init__(self, typ, D=8, W=256, skips=[4], in_channels_xyz=63, use_viewdir=True, in_channels_dir=27,
                 encode_appearance=False, in_channels_a=48, encode_transient=False, in_channels_t=16,
                 output_flow=False, flow_scale=0.2):
        super().__init__()
        self.typ, self.D, self.W, self.skips = typ, D, W, skips
        self.in_channels_xyz, self.use_viewdir, self.in_channels_dir = in_channels_xyz, use_viewdir, in_channels_dir
        self.encode_appearance = False if typ == "coarse" else encode_appearance
        self.in_channels_a = in_channels_a if encode_appearance else 0
        self.encode_transient = encode_transient
        self.in_channels_t = in_channels_t if encode_transient else 0
        self.output_flow = self.encode_transient and output_flow
        self.precision = "bf16"

        def trunk(prefix, n_in):
            for i in range(D):
                width_in = n_in if i == 0 else (W + n_in if i in skips else W)
                setattr(self, f"{prefix}_xyz_encoding_{i + 1}", _act_linear(width_in, W, nn.ReLU(True)))
            setattr(self, f"{prefix}_xyz_encoding_final", nn.Linear(W, W))

        # creation order == reference order, so a given torch seed yields identical initial weights
        trunk("static", in_channels_xyz)
        if use_viewdir:
            self.static_dir_encoding = _act_linear(W + in_channels_dir + self.in_channels_a, W, nn.ReLU(True))
        self.static_sigma = nn.Linear(W, 1)
        self.static_rgb = _act_linear(W, 3, nn.Sigmoid())
        if encode_transient:
            trunk("transient", in_channels_xyz + in_channels_t)
            self.transient_sigma = nn.Linear(W, 1)
            self.transient_rgb = _act_linear(W, 3, nn.Sigmoid())
            if typ == "fine" and self.output_flow:
                self.flow_scale = flow_scale
                self.transient_flow_fw = _act_linear(W, 3, nn.Tanh())
                self.transient_flow_bw = _act_linear(W, 3, nn.Tanh())
        self._pack = None
        self._pack_key = None

    # ---- C-ABI views of this module ----------------------------------------------------------------
    @property
    def has_flow_heads(self):
        return hasattr(self, "transient_flow_fw")

    def field_desc(self):
        if len(self.skips) != 1:
            raise NotImplementedError("nsff_pl_b200 kernels support exactly one skip layer")
        return _cabi.FieldDesc(W=self.W, D=self.D, skip=self.skips[0], in_xyz=self.in_channels_xyz,
                               in_dir=self.in_channels_dir, in_a=self.in_channels_a, in_t=self.in_channels_t,
                               use_viewdir=int(self.use_viewdir), has_transient=int(self.encode_transient),
                               has_flow=int(self.has_flow_heads), flow_scale=float(getattr(self, "flow_scale", 0.2)))

    def _lin(self, name):
        m = getattr(self, name, None)
        if m is None:
            return None, None
        m = m[0] if isinstance(m, nn.Sequential) else m
        for p, n in ((m.weight, name + ".weight"), (m.bias, name + ".bias")):
            _cabi.dev_tensor(p.data, n)
        return m.weight, m.bias

    def field_params(self):
        """FieldParams struct of raw fp32 device pointers (+ the tensors, to keep them alive)."""
        fp = _cabi.FieldParams()
        keep = []

        def put(field, idx, name):
            w, b = self._lin(name)
            keep.extend([w, b])
            wp = None if w is None else w.data_ptr()
            bp = None if b is None else b.data_ptr()
            if idx is None:
                setattr(fp, field + "_w", wp)
                setattr(fp, field + "_b", bp)
            else:
                getattr(fp, field + "_w")[idx] = wp
                getattr(fp, field + "_b")[idx] = bp

        if self.D > 8:
            raise NotImplementedError("D > 8 is not supported")
        for i in range(self.D):
            put("static", i, f"static_xyz_encoding_{i + 1}")
            if self.encode_transient:
                put("trans", i, f"transient_xyz_encoding_{i + 1}")
        put("static_final", None, "static_xyz_encoding_final")
        put("static_dir", None, "static_dir_encoding")
        put("static_sigma", None, "static_sigma")
        put("static_rgb", None, "static_rgb")
        if self.encode_transient:
            put("trans_final", None, "transient_xyz_encoding_final")
            put("trans_sigma", None, "transient_sigma")
            put("trans_rgb", None, "transient_rgb")
            put("trans_fw", None, "transient_flow_fw")
            put("trans_bw", None, "transient_flow_bw")
        return fp, keep

    def field_pack(self):
        """bf16 swizzled weight image for the tcgen05 kernel; rebuilt when any parameter changed
        (optimizer step, load_state_dict, .to()) - detected through tensor version counters."""
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._pack is None or key != self._pack_key:
            desc = self.field_desc()
            fp, keep = self.field_params()
            L = _cabi.lib()
            nbytes = C.c_size_t(0)
            _cabi.check(L.nsff_field_pack_bytes(C.byref(desc), C.byref(nbytes)), "nsff_field_pack_bytes")
            dev = params[0].device
            if self._pack is None or self._pack.numel() != nbytes.value or self._pack.device != dev:
                self._pack = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            _cabi.select_device(self._pack)
            _cabi.check(L.nsff_field_pack(C.byref(desc), C.byref(fp), _cabi.ptr(self._pack), nbytes.value,
                                          _cabi.stream_ptr(dev)), "nsff_field_pack")
            del keep
            self._pack_key = key
        return self._pack

    def field_bwd_pack(self):
        """W and W^T bf16 blocks for the tensor-core backward (nsff_field_backward_tc); same cache rule."""
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if getattr(self, "_bwd_pack", None) is None or key != self._bwd_pack_key:
            desc = self.field_desc()
            fp, keep = self.field_params()
            L = _cabi.lib()
            nbytes = L.nsff_field_bwd_tc_pack_bytes(C.byref(desc))
            if nbytes == 0:
                raise NotImplementedError("tensor-core backward needs the 8x256 / skip-4 NSFF field shape")
            dev = params[0].device
            self._bwd_pack = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _cabi.select_device(self._bwd_pack)
            _cabi.check(L.nsff_field_bwd_tc_pack(C.byref(desc), C.byref(fp), _cabi.ptr(self._bwd_pack),
                                                 _cabi.stream_ptr(dev)), "nsff_field_bwd_tc_pack")
            del keep
            self._bwd_pack_key = key
        return self._bwd_pack

    def backward_precision(self):
        """'bf16' (tcgen05 backward) or 'fp32' (CUDA-core verification backward); follows ``precision`` unless
        ``grad_precision`` is set on the module."""
        gp = getattr(self, "grad_precision", None) or self.precision
        if gp not in ("bf16", "fp32"):
            raise ValueError(f"unknown grad_precision {gp!r} (use 'bf16' or 'fp32')")
        return gp

    @staticmethod
    def out_mask(sigma_only, output_static, output_transient, output_transient_flow):
        m = 0
        if output_static:
            m |= _cabi.OUT_STATIC
        if output_transient:
            m |= _cabi.OUT_TRANSIENT
        if sigma_only:
            m |= _cabi.OUT_SIGMA_ONLY
        else:
            if "fw" in output_transient_flow:
                m |= _cabi.OUT_FLOW_FW
            if "bw" in output_transient_flow:
                m |= _cabi.OUT_FLOW_BW
        return m

    @staticmethod
    def out_channels(mask):
        if mask & _cabi.OUT_SIGMA_ONLY:
            return bool(mask & _cabi.OUT_STATIC) + bool(mask & _cabi.OUT_TRANSIENT)
        return 4 * bool(mask & _cabi.OUT_STATIC) + 4 * bool(mask & _cabi.OUT_TRANSIENT) + \
            3 * bool(mask & _cabi.OUT_FLOW_FW) + 3 * bool(mask & _cabi.OUT_FLOW_BW)

    def macs_per_point(self, mask):
        """Algorithmic (unpadded) multiply-accumulates per point of a forward with head mask `mask`
        (W=256, D=8 shapes)."""
        sigma_only = bool(mask & _cabi.OUT_SIGMA_ONLY)
        n = 0
        if mask & _cabi.OUT_STATIC:
            n += _MAC_STATIC_SIGMA_ONLY if sigma_only else _MAC_STATIC
        if mask & _cabi.OUT_TRANSIENT:
            n += _MAC_TRANS_BASE
            if not sigma_only:
                n += 768 * (1 + bool(mask & _cabi.OUT_FLOW_FW) + bool(mask & _cabi.OUT_FLOW_BW))
        return n

    def _guard_grad(self):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "nsff_pl_b200: NeRF.forward on a pre-embedded row records no autograd graph - call it under "
                "torch.no_grad(); training goes through render_rays (models/autograd.py), which is differentiable.")

    def executed_macs_per_point(self, mask, every_layer=False):
        """Multiply-accumulates per point the tensor-core kernel actually issues (unpadded): the inference program
        folds the activation-free 256x256 `*_xyz_encoding_final` layers into their successors (nsff_field_tc.cu,
        build_layout), i.e. 65,536 fewer per branch - except in sigma-only mode for the static branch (no final
        layer there) and when every layer is evaluated (training forward)."""
        n = self.macs_per_point(mask)
        if every_layer or self.precision != "bf16":
            return n
        if (mask & _cabi.OUT_STATIC) and not (mask & _cabi.OUT_SIGMA_ONLY):
            n -= 256 * 256
        if mask & _cabi.OUT_TRANSIENT:
            n -= 256 * 256
        return n

    def run_field(self, mask, out, pts=None, x=None, save=None, every_layer=False):
        """Launch the MLP on raw points (`pts`: _cabi.Points) or an embedded matrix `x` into `out` (B, C).
        `save` (uint8 CUDA tensor of nsff_field_save_bytes): training mode, the kernel also writes every hidden
        activation for the tensor-core backward.  `every_layer`: evaluate the reference's layers one by one
        (NSFF_OUT_EVERY_LAYER; forwards whose backward differentiates the un-folded network)."""
        L = _cabi.lib()
        desc = self.field_desc()
        _cabi.select_device(out)
        st = _cabi.stream_ptr(out.device)
        B = out.shape[0]
        if self.precision == "bf16":
            pack = self.field_pack()
            ev = None
            if FIELD_EVENTS is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record(torch.cuda.current_stream(out.device))
            cmask = mask | (_cabi.OUT_EVERY_LAYER if every_layer else 0)
            if pts is not None and save is not None:
                rc = L.nsff_field_forward_save(C.byref(desc), _cabi.ptr(pack), C.byref(pts), mask, _cabi.ptr(out),
                                               out.stride(0), _cabi.ptr(save), st)
            elif pts is not None:
                rc = L.nsff_field_forward(C.byref(desc), _cabi.ptr(pack), C.byref(pts), cmask, _cabi.ptr(out),
                                          out.stride(0), st)
            else:
                rc = L.nsff_field_forward_embedded(C.byref(desc), _cabi.ptr(pack), _cabi.ptr(x), x.stride(0), B, cmask,
                                                   _cabi.ptr(out), out.stride(0), st)
            _cabi.check(rc, "nsff_field_forward")
            if ev is not None:
                ev[1].record(torch.cuda.current_stream(out.device))
                FIELD_EVENTS.append((ev[0], ev[1], 2.0 * B * self.macs_per_point(mask),
                                     2.0 * B * self.executed_macs_per_point(mask, every_layer or save is not None)))
        elif self.precision == "fp32":
            fp, keep = self.field_params()
            step = 1 << 16                       # bound the fp32 scratch (3 hidden planes per point)
            if pts is not None:                  # ... in whole rays (the per-ray code / direction rows are indexed by ray)
                step = max(1, step // int(pts.S)) * int(pts.S)
            for b0 in range(0, B, step):
                nb = min(step, B - b0)
                scratch = torch.empty(L.nsff_field_fp32_scratch_floats(C.byref(desc), nb), device=out.device,
                                      dtype=torch.float32)
                if pts is not None:
                    sub = _cabi.Points()
                    C.memmove(C.byref(sub), C.byref(pts), C.sizeof(sub))
                    if b0 % pts.S != 0:
                        raise AssertionError("fp32 chunking must split on ray boundaries")
                    ray0 = b0 // pts.S
                    sub.xyz = pts.xyz + b0 * 12
                    if pts.offset:
                        sub.offset = pts.offset + b0 * pts.offset_ld * 4
                    if pts.xyz_out:
                        sub.xyz_out = pts.xyz_out + b0 * 12
                    if pts.dirs:
                        sub.dirs = pts.dirs + ray0 * 12
                    if pts.t_index:
                        sub.t_index = pts.t_index + ray0 * 8
                    elif pts.t_table:
                        sub.t_table = pts.t_table + ray0 * self.in_channels_t * 4
                    if pts.a_index:
                        sub.a_index = pts.a_index + ray0 * 8
                    elif pts.a_table:
                        sub.a_table = pts.a_table + ray0 * self.in_channels_a * 4
                    if pts.zs:
                        sub.zs = pts.zs + b0 * 4
                    sub.B = nb
                    rc = L.nsff_field_forward_fp32(C.byref(desc), C.byref(fp), C.byref(sub), None, 0, 0, mask,
                                                   C.c_void_p(out.data_ptr() + b0 * out.stride(0) * 4), out.stride(0),
                                                   _cabi.ptr(scratch), st)
                else:
                    rc = L.nsff_field_forward_fp32(C.byref(desc), C.byref(fp), None,
                                                   C.c_void_p(x.data_ptr() + b0 * x.stride(0) * 4), x.stride(0), nb,
                                                   mask, C.c_void_p(out.data_ptr() + b0 * out.stride(0) * 4),
                                                   out.stride(0), _cabi.ptr(scratch), st)
                _cabi.check(rc, "nsff_field_forward_fp32")
            del keep
        else:
            raise ValueError(f"unknown precision {self.precision!r} (use 'bf16' or 'fp32')")

    def forward(self, x, sigma_only=False, output_static=True, output_transient=True, output_transient_flow=[]):
        """Same contract as reference nerf.py:118-213: `x` is the embedded row
        [xyz | dir | a | t] (or [xyz | t] / [xyz] with sigma_only); returns the concatenated heads."""
        if sigma_only and not output_static:
            raise ValueError("sigma_only requires output_static (reference nerf.py:162-180)")
        self._guard_grad()
        x = _cabi.dev_tensor(x if x.is_contiguous() else x.contiguous(), "x")
        cx, cd, ca, ct = self.in_channels_xyz, self.in_channels_dir, self.in_channels_a, self.in_channels_t
        want = cx + (ct if output_transient else 0) if sigma_only else cx + cd + ca + (ct if output_transient else 0)
        if x.dim() != 2 or x.shape[1] != want:
            raise ValueError(f"expected x of shape (B, {want}) for this call mode, got {tuple(x.shape)}")
        if sigma_only and self.precision == "fp32":
            # the fp32 entry reads the full-row column layout; widen [xyz | t] to it
            full = torch.zeros(x.shape[0], cx + cd + ca + ct, device=x.device, dtype=torch.float32)
            full[:, :cx] = x[:, :cx]
            if output_transient:
                full[:, cx + cd + ca:] = x[:, cx:]
            x = full
        elif (not sigma_only) and (not output_transient) and self.precision == "fp32" and ct:
            x = torch.cat([x, torch.zeros(x.shape[0], ct, device=x.device)], 1)
        mask = self.out_mask(sigma_only, output_static, output_transient, output_transient_flow)
        out = torch.empty(x.shape[0], self.out_channels(mask), device=x.device, dtype=torch.float32)
        if x.shape[0]:
            self.run_field(mask, out, x=x)
        return out
