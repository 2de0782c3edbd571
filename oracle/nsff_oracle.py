"""CPU oracle for the NSFF render_rays hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional, fp32, CPU restatement (torch CPU tensors, no
nn.Module, no einops) of the algorithm of kwea123/nsff_pl's volume renderer.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it; the product package
``nsff_pl_b200`` never does (tests/test_host_logic.py greps for that).

Parity status: PINNED.  The reference holds no golden vectors of its own
(SURVEY.md §4), so the oracle is pinned against outputs of the unmodified
reference run in the build container: ``tests/golden/make_golden.py`` imports
``/root/reference/models/{nerf,rendering}.py`` and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays every fixture
through this file.

Each function cites the reference lines it follows (paths under
/root/reference/).  Networks are passed as ``FieldSpec`` (plain dict of fp32
tensors keyed by the reference's state_dict names + the handful of
constructor flags the renderer reads).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Z_FAR = 0.95           # models/rendering.py:316
LAST_DELTA_STATIC = 100.0      # models/rendering.py:203
LAST_DELTA_TRANSIENT = 1e-3    # models/rendering.py:204
GREY_BG = 0.8          # models/rendering.py:265


# --------------------------------------------------------------------------
# field description
# --------------------------------------------------------------------------
@dataclasses.dataclass
class FieldSpec:
    """What models/nerf.py:33-116 stores in a NeRF module, as plain data."""
    typ: str                       # 'coarse' | 'fine'
    params: Dict[str, torch.Tensor]  # reference state_dict names -> fp32 tensors
    D: int = 8
    W: int = 256
    skips: Sequence[int] = (4,)
    in_channels_xyz: int = 63
    use_viewdir: bool = True
    in_channels_dir: int = 27
    encode_appearance: bool = False
    in_channels_a: int = 0
    encode_transient: bool = False
    in_channels_t: int = 0
    output_flow: bool = False
    flow_scale: float = 0.2

    @staticmethod
    def from_module(m) -> "FieldSpec":
        """Build from anything that quacks like the reference NeRF module."""
        sd = {k: v.detach().to(torch.float32).cpu().clone() for k, v in m.state_dict().items()}
        return FieldSpec(
            typ=m.typ, params=sd, D=m.D, W=m.W, skips=tuple(m.skips),
            in_channels_xyz=m.in_channels_xyz, use_viewdir=m.use_viewdir,
            in_channels_dir=m.in_channels_dir,
            encode_appearance=bool(m.encode_appearance), in_channels_a=m.in_channels_a,
            encode_transient=bool(m.encode_transient), in_channels_t=m.in_channels_t,
            output_flow=bool(m.output_flow), flow_scale=float(getattr(m, "flow_scale", 0.2)))


@dataclasses.dataclass
class Encoders:
    """embeddings dict of the reference (train.py:47-59) as plain data."""
    xyz_freqs: torch.Tensor                 # PosEmbedding(9,10).freqs
    dir_freqs: torch.Tensor                 # PosEmbedding(3,4).freqs
    t_table: Optional[torch.Tensor] = None  # nn.Embedding(N_frames, N_tau).weight
    a_table: Optional[torch.Tensor] = None  # nn.Embedding(N_frames, N_a).weight


def pos_freqs(max_logscale: int, n_freqs: int, logscale: bool = True) -> torch.Tensor:
    """models/nerf.py:12-15."""
    if logscale:
        return 2 ** torch.linspace(0, max_logscale, n_freqs)
    return torch.linspace(1, 2 ** max_logscale, n_freqs)


def pos_embedding(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """models/nerf.py:17-30: [x, sin(f0 x), cos(f0 x), sin(f1 x), ...]."""
    cols = [x]
    for f in freqs:
        cols.append(torch.sin(f * x))
        cols.append(torch.cos(f * x))
    return torch.cat(cols, dim=-1)


# --------------------------------------------------------------------------
# the two MLPs  (models/nerf.py:118-213)
# --------------------------------------------------------------------------
def _lin(p, name, x):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def _trunk(spec: FieldSpec, prefix: str, skip_in: torch.Tensor) -> torch.Tensor:
    """8x(Linear+ReLU) with the input re-concatenated at the skip layer
    (models/nerf.py:163-167 static, 192-196 transient)."""
    h = skip_in
    for i in range(spec.D):
        if i in spec.skips:
            h = torch.cat([skip_in, h], dim=1)
        h = torch.relu(_lin(spec.params, f"{prefix}_xyz_encoding_{i + 1}.0", h))
    return h


def nerf_forward(spec: FieldSpec, x: torch.Tensor, sigma_only: bool = False,
                 output_static: bool = True, output_transient: bool = True,
                 output_transient_flow: Sequence[str] = ()) -> torch.Tensor:
    """models/nerf.py:118-213, all four call modes."""
    p = spec.params
    cx, cd, ca, ct = spec.in_channels_xyz, spec.in_channels_dir, spec.in_channels_a, spec.in_channels_t
    # column split (models/nerf.py:142-160)
    if sigma_only:
        if output_transient:
            in_xyz, in_t = x[:, :cx], x[:, cx:cx + ct]
            assert x.shape[1] == cx + ct
        else:
            in_xyz = x
    elif output_transient:
        assert x.shape[1] == cx + cd + ca + ct
        in_xyz, in_dir = x[:, :cx], x[:, cx:cx + cd]
        in_a, in_t = x[:, cx + cd:cx + cd + ca], x[:, cx + cd + ca:]
    else:
        assert x.shape[1] == cx + cd + ca
        in_xyz, in_dir, in_a = x[:, :cx], x[:, cx:cx + cd], x[:, cx + cd:]

    static = None
    if output_static:
        h = _trunk(spec, "static", in_xyz)
        s_sigma = _lin(p, "static_sigma", h)                       # nerf.py:169
        if sigma_only:
            if not output_transient:
                return s_sigma
            ht = _trunk(spec, "transient", torch.cat([in_xyz, in_t], 1))
            ft = _lin(p, "transient_xyz_encoding_final", ht)
            return torch.cat([s_sigma, _lin(p, "transient_sigma", ft)], 1)   # nerf.py:173-180
        feat = _lin(p, "static_xyz_encoding_final", h)             # nerf.py:182
        if spec.use_viewdir:
            feat = torch.relu(_lin(p, "static_dir_encoding.0", torch.cat([feat, in_dir, in_a], 1)))
        s_rgb = torch.sigmoid(_lin(p, "static_rgb.0", feat))
        static = torch.cat([s_rgb, s_sigma], 1)
        if not output_transient:
            return static

    ht = _trunk(spec, "transient", torch.cat([in_xyz, in_t], 1))
    ft = _lin(p, "transient_xyz_encoding_final", ht)
    pieces = [torch.sigmoid(_lin(p, "transient_rgb.0", ft)), _lin(p, "transient_sigma", ft)]
    if "fw" in output_transient_flow:
        pieces.append(spec.flow_scale * torch.tanh(_lin(p, "transient_flow_fw.0", ft)))
    if "bw" in output_transient_flow:
        pieces.append(spec.flow_scale * torch.tanh(_lin(p, "transient_flow_bw.0", ft)))
    transient = torch.cat(pieces, 1)
    return torch.cat([static, transient], 1) if output_static else transient


# --------------------------------------------------------------------------
# eval-only camera visibility (datasets/ray_utils.py:127-181)
# --------------------------------------------------------------------------
def ndc2world(xyz: torch.Tensor, K: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    rz = 2 / (xyz[..., 2] - 1 - eps)
    rx = -rz * xyz[..., 0] * cx / fx
    ry = -rz * xyz[..., 1] * cy / fy
    return torch.stack([rx, ry, rz], -1)


def add_camera_visibility(vis: torch.Tensor, xyz_w: torch.Tensor, K: torch.Tensor,
                          H: int, W: int, c2w: torch.Tensor) -> None:
    """datasets/ray_utils.py:154-181; accumulates 0/1 into ``vis`` in place."""
    c2w4 = torch.eye(4)
    c2w4[:3] = c2w
    w2c = torch.inverse(c2w4)
    cam = w2c[:3, :3] @ xyz_w.T + w2c[:3, 3:]
    front = cam[2] < 0
    if front.sum() == 0:
        return
    cam = cam[:, front]
    cam[1:] *= -1
    img = K @ cam
    img = img[:2] / img[2]
    vis[front] += ((img[0] >= 0) & (img[0] < W) & (img[1] >= 0) & (img[1] < H)).float()


# --------------------------------------------------------------------------
# inverse-CDF resampling (models/rendering.py:10-49)
# --------------------------------------------------------------------------
def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_importance: int,
               det: bool = False, eps: float = 1e-5) -> torch.Tensor:
    n_rays, n_w = weights.shape
    w = weights + eps
    pdf = w / w.sum(1, keepdim=True)
    cdf = torch.cat([torch.zeros(n_rays, 1), torch.cumsum(pdf, -1)], -1)
    if det:
        u = torch.linspace(0, 1, n_importance).expand(n_rays, n_importance)
    else:
        u = torch.rand(n_rays, n_importance)
    u = u.contiguous()
    idx = torch.searchsorted(cdf, u, right=True)
    lo = torch.clamp_min(idx - 1, 0)
    hi = torch.clamp_max(idx, n_w)
    c_lo, c_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    b_lo, b_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    den = c_hi - c_lo
    den = torch.where(den < eps, torch.ones_like(den), den)
    return b_lo + (u - c_lo) / den * (b_hi - b_lo)


# --------------------------------------------------------------------------
# alpha compositing helpers (models/rendering.py:202-215, 234-241)
# --------------------------------------------------------------------------
def _excl_cumprod_one_minus(alpha: torch.Tensor) -> torch.Tensor:
    """T_i = prod_{j<i} (1 - alpha_j)   (rendering.py:234-235)."""
    sh = torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha], -1)
    return torch.cumprod(sh[:, :-1], -1)


def _wsum(w: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """sum over samples of w (R,S) times v (R,S,C) -> (R,C)."""
    return (w.unsqueeze(-1) * v).sum(1)


# --------------------------------------------------------------------------
# render_rays (models/rendering.py:52-362)
# --------------------------------------------------------------------------
def render_rays(models: Dict[str, FieldSpec], enc: Encoders, rays: torch.Tensor,
                ts: Optional[torch.Tensor], max_t: int, N_samples: int = 64,
                perturb: float = 0, noise_std: float = 0, N_importance: int = 0,
                chunk: int = 1024 * 32, test_time: bool = False, **kwargs) -> Dict[str, torch.Tensor]:
    """Same result dict, same RNG draw order, as models/rendering.py:52-362.
    ``chunk`` is accepted and ignored (chunking does not change results)."""
    results: Dict[str, torch.Tensor] = {}
    n_rays = rays.shape[0]
    rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
    dir_embedded = pos_embedding(kwargs.get("view_dir", rays_d), enc.dir_freqs)   # :308

    def t_code(idx):
        return F.embedding(idx, enc.t_table)

    def field_pass(spec: FieldSpec, xyz: torch.Tensor, zs: torch.Tensor,
                   want_transient: bool, flow_heads: List[str],
                   t_embedded, a_embedded) -> None:
        """`inference` closure of the reference (rendering.py:83-300)."""
        typ = spec.typ
        S = xyz.shape[1]
        results[f"zs_{typ}"] = zs
        results[f"xyzs_{typ}"] = xyz
        pts = xyz.reshape(-1, 3)

        def per_point(v):                      # 'n1 c -> (n1 n2) c'
            return v.unsqueeze(1).expand(n_rays, S, v.shape[-1]).reshape(n_rays * S, -1)

        if typ == "coarse" and test_time:                                    # :151-162
            cols = [pos_embedding(pts, enc.xyz_freqs)]
            if want_transient:
                cols.append(per_point(t_embedded))
            out = nerf_forward(spec, torch.cat(cols, 1), sigma_only=True,
                               output_transient=want_transient).reshape(n_rays, S, -1)
            s_raw = out[..., 0]
            t_raw = out[..., 1] if want_transient else None
            s_rgb = t_rgb = flow_fw = flow_bw = None
        else:                                                                 # :163-188
            cols = [pos_embedding(pts, enc.xyz_freqs), per_point(dir_embedded)]
            if spec.encode_appearance:
                cols.append(per_point(a_embedded))
            if want_transient:
                cols.append(per_point(t_embedded))
            out = nerf_forward(spec, torch.cat(cols, 1), output_transient=want_transient,
                               output_transient_flow=flow_heads).reshape(n_rays, S, -1)
            results[f"static_rgbs_{typ}"] = s_rgb = out[..., :3]
            s_raw = out[..., 3]
            t_rgb = t_raw = flow_fw = flow_bw = None
            if want_transient:
                results[f"transient_rgbs_{typ}"] = t_rgb = out[..., 4:7]
                t_raw = out[..., 7]
                if flow_heads:
                    far = (zs > Z_FAR).unsqueeze(-1)
                    flow_fw = torch.where(far, torch.zeros(()), out[..., 8:11])
                    flow_bw = torch.where(far, torch.zeros(()), out[..., 11:14])
                    results["transient_flows_fw"] = flow_fw
                    results["transient_flows_bw"] = flow_bw

        if test_time and want_transient and "dataset" in kwargs:              # :190-200
            ds = kwargs["dataset"]
            K = ds.Ks[0]
            vis = torch.zeros(len(pts))
            pts_w = ndc2world(pts, K)
            for i in range(len(ds.cam_train)):
                pose = torch.as_tensor(ds.poses[i * ds.N_frames + int(ts[0])], dtype=torch.float32)
                add_camera_visibility(vis, pts_w, K, ds.img_wh[1], ds.img_wh[0], pose)
            t_raw = torch.where(vis.view_as(t_raw) == 0, torch.full((), -10.0), t_raw)

        dz = zs[:, 1:] - zs[:, :-1]                                           # :202-204
        d_static = torch.cat([dz, torch.full_like(dz[:, :1], LAST_DELTA_STATIC)], -1)
        d_trans = torch.cat([dz, torch.full_like(dz[:, :1], LAST_DELTA_TRANSIENT)], -1)

        s_sigma = F.softplus(s_raw + torch.randn_like(s_raw) * noise_std)     # :206-207
        results[f"static_sigmas_{typ}"] = s_sigma
        alpha = 1 - torch.exp(-d_static * s_sigma)

        if want_transient:
            a_s = alpha
            t_sigma = F.softplus(t_raw + torch.randn_like(t_raw) * noise_std)  # :212-213
            results[f"transient_sigmas_{typ}"] = t_sigma
            a_t = 1 - torch.exp(-d_trans * t_sigma)
            alpha = 1 - (1 - a_s) * (1 - a_t)

            if (not test_time) and flow_heads:                                # :217-232
                def warp(xyz_w, code, head):
                    """render_transient_warping (rendering.py:98-140)."""
                    cols_w = [pos_embedding(xyz_w.reshape(-1, 3), enc.xyz_freqs), per_point(dir_embedded)]
                    if spec.encode_appearance:
                        cols_w.append(per_point(a_embedded))
                    cols_w.append(per_point(code))
                    o = nerf_forward(spec, torch.cat(cols_w, 1), output_static=False,
                                     output_transient=True,
                                     output_transient_flow=[head]).reshape(n_rays, S, -1)
                    rgb_w, raw_w = o[..., :3], o[..., 3]
                    fl_w = torch.where((zs > Z_FAR).unsqueeze(-1), torch.zeros(()), o[..., 4:7])
                    a_tw = 1 - torch.exp(-d_trans * F.softplus(raw_w + torch.randn_like(raw_w) * noise_std))
                    a_w = 1 - (1 - a_s) * (1 - a_tw)
                    T_w = _excl_cumprod_one_minus(a_w)
                    rgb_map = _wsum(a_s * T_w, s_rgb) + _wsum(a_tw * T_w, rgb_w)
                    return rgb_map, fl_w, a_tw * T_w

                results["xyzs_fw"] = xyz_fw = xyz + flow_fw
                results["rgb_fw"], fl_fw_bw, w_t_fw = \
                    warp(xyz_fw, t_code(torch.clamp(ts + 1, max=max_t)), "bw")
                results["xyzs_bw"] = xyz_bw = xyz + flow_bw
                results["rgb_bw"], fl_bw_fw, w_t_bw = \
                    warp(xyz_bw, t_code(torch.clamp(ts - 1, min=0)), "fw")
                results["xyzs_fw_bw"] = xyz_fw + fl_fw_bw
                results["xyzs_bw_fw"] = xyz_bw + fl_bw_fw

        T = _excl_cumprod_one_minus(alpha)                                    # :234-235
        w_all = alpha * T
        if want_transient:
            w_s, w_t = a_s * T, a_t * T
            results[f"static_weights_{typ}"] = w_s
            results[f"transient_weights_{typ}"] = w_t
            results[f"weights_{typ}"] = w_all
        else:
            results[f"static_weights_{typ}"] = w_all
        if test_time:                                                         # :249-254
            if want_transient:
                results[f"static_alphas_{typ}"] = a_s
                results[f"transient_alphas_{typ}"] = a_t
            if typ == "coarse":
                return

        results[f"depth_{typ}"] = (w_all * zs).sum(1)                         # :256
        if want_transient:
            t_map = _wsum(w_t, t_rgb)
            results[f"rgb_{typ}"] = _wsum(w_s, s_rgb) + t_map
            results[f"transient_alpha_{typ}"] = t_alpha = w_t.sum(1)
            results[f"transient_rgb_{typ}"] = t_map + GREY_BG * (1 - t_alpha.unsqueeze(-1))
            w_only_s = a_s * _excl_cumprod_one_minus(a_s)                      # :270-278
            results[f"_static_rgb_{typ}"] = _wsum(w_only_s, s_rgb)
            results[f"_static_depth_{typ}"] = (w_only_s * zs).sum(1)
            if flow_heads:                                                    # :280-295
                results["xyz_fine"] = _wsum(w_all, xyz)
                results["transient_flow_fw"] = _wsum(w_all, flow_fw)
                results["xyz_fw"] = results["xyz_fine"] + results["transient_flow_fw"]
                results["transient_flow_bw"] = _wsum(w_all, flow_bw)
                results["xyz_bw"] = results["xyz_fine"] + results["transient_flow_bw"]
                if (not test_time) and "disocc" in flow_heads:
                    occ_fw, occ_bw = (w_t_fw - w_t).detach(), (w_t_bw - w_t).detach()      # :290-291
                    results["disocc_fw"] = 1 - occ_fw.sum(1, keepdim=True).abs()
                    results["disoccs_fw"] = 1 - occ_fw.unsqueeze(-1).abs()
                    results["disocc_bw"] = 1 - occ_bw.sum(1, keepdim=True).abs()
                    results["disoccs_bw"] = 1 - occ_bw.unsqueeze(-1).abs()
        else:
            results[f"rgb_{typ}"] = _wsum(w_all, s_rgb)                        # :298

    # ---- stratified depths (rendering.py:313-324) -------------------------
    zs = torch.linspace(0, 1, N_samples).expand(n_rays, N_samples)
    zs_mid = 0.5 * (zs[:, :-1] + zs[:, 1:])
    if perturb > 0:
        upper = torch.cat([zs_mid, zs[:, -1:]], -1)
        lower = torch.cat([zs[:, :1], zs_mid], -1)
        zs = lower + (upper - lower) * (perturb * torch.rand(n_rays, N_samples))

    t_embedded = a_embedded = None
    want_transient = False
    if N_importance > 0:                                                       # :326-348
        spec = models["coarse"]
        want_transient = kwargs.get("output_transient", True) and spec.encode_transient
        if want_transient:
            t_embedded = kwargs["t_embedded"] if "t_embedded" in kwargs else t_code(ts)
        xyz_c = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * zs.unsqueeze(-1)
        field_pass(spec, xyz_c, zs, want_transient, [], t_embedded, None)
        det = perturb == 0
        zs_static = sample_pdf(zs_mid, results["static_weights_coarse"][:, 1:-1].detach(), N_importance, det=det)   # :336
        merged = [zs, zs_static]
        if test_time:
            results["static_zs_fine"] = zs_static
        if want_transient:
            zs_trans = sample_pdf(zs_mid, results["transient_weights_coarse"][:, 1:-1].detach(), N_importance, det=det)   # :343
            merged.append(zs_trans)
            if test_time:
                results["transient_zs_fine"] = zs_trans
        zs = torch.sort(torch.cat(merged, -1), -1)[0]
        if "_zs_fine" in kwargs:
            # TEST HOOK (not in the reference): evaluate the fine pass on externally supplied depths, so that a bf16
            # implementation whose resampled depths differ in the last bits can still be compared sample by sample
            zs = kwargs["_zs_fine"]

    spec = models["fine"]                                                      # :350-360
    if spec.encode_appearance:
        a_embedded = kwargs["a_embedded"] if "a_embedded" in kwargs else F.embedding(ts, enc.a_table)
    if N_importance == 0:
        want_transient = kwargs.get("output_transient", True) and spec.encode_transient
        if want_transient:
            t_embedded = kwargs["t_embedded"] if "t_embedded" in kwargs else t_code(ts)
    flow_heads = list(kwargs.get("output_transient_flow", [])) if want_transient else []
    xyz_f = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * zs.unsqueeze(-1)
    field_pass(spec, xyz_f, zs, want_transient, flow_heads, t_embedded, a_embedded)
    return results


# --------------------------------------------------------------------------
# test-time frame interpolation (models/rendering.py:365-460) with the forward
# "average" softmax-splat of models/softsplat.py:6-44,303-329 restated on CPU
# --------------------------------------------------------------------------
def splat_average(inp: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """FunctionSoftsplat(tenInput (1,C,H,W), tenFlow (1,2,H,W), None, 'average'): every source pixel adds
    input*bilinear_weight to the 4 integer neighbours of (x+flow_x, y+flow_y); a ones channel accumulates the
    weights and the result is divided by it (zero weight -> 1)."""
    _, Cc, H, W = inp.shape
    src = torch.cat([inp, torch.ones(1, 1, H, W)], 1)[0].reshape(Cc + 1, -1)          # (C+1, HW)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    ox = (xs + flow[0, 0]).reshape(-1)
    oy = (ys + flow[0, 1]).reshape(-1)
    x0, y0 = torch.floor(ox), torch.floor(oy)
    out = torch.zeros(Cc + 1, H * W)
    for dx, dy in ((0, 0), (1, 0), (0, 1), (1, 1)):
        xi, yi = x0 + dx, y0 + dy
        wgt = (1 - (ox - xi).abs()) * (1 - (oy - yi).abs())
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        idx = (yi * W + xi).long()[ok]
        out.index_add_(1, idx, src[:, ok] * wgt[ok])
    norm = out[-1:]
    norm = torch.where(norm == 0, torch.ones_like(norm), norm)
    return (out[:-1] / norm).reshape(1, Cc, H, W)


def interpolate(results_t, results_tp1, dt, K, c2w, img_wh):
    """models/rendering.py:365-460: per-sample-plane forward splat of the dynamic RGBA of frame t (by dt * flow_fw)
    and of frame t+1 (by (1-dt) * flow_bw), blended and composited front to back with the static plane."""
    w, h = img_wh
    xyzs = results_t["xyzs_fine"]
    R, S = xyzs.shape[:2]
    c2w4 = torch.eye(4)
    c2w4[:3] = c2w
    w2c = torch.inverse(c2w4)[:3]
    w2c[1:] *= -1
    P = K @ w2c
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    grid = torch.stack([xs, ys], -1)                                                  # (h, w, 2) = (x, y)
    pts = xyzs.reshape(-1, 3)
    world = ndc2world(pts, K)

    def plane_flow(flows, scale):
        moved = ndc2world((xyzs + flows).reshape(-1, 3), K)
        moved = world + scale * (moved - world)
        uvd = P[:, :3] @ moved.T + P[:, 3:]
        uv = (uvd[:2] / uvd[2]).reshape(2, h, w, S)
        return uv.permute(3, 0, 1, 2) - grid.permute(2, 0, 1).unsqueeze(0)           # (S, 2, h, w)

    of_fw = plane_flow(results_t["transient_flows_fw"], dt)
    of_bw = plane_flow(results_tp1["transient_flows_bw"], 1 - dt)

    def rgba(res):
        rgb = res["transient_rgbs_fine"].reshape(h, w, S, 3).permute(2, 3, 0, 1)
        a = res["transient_alphas_fine"].reshape(h, w, S).permute(2, 0, 1).unsqueeze(1)
        return torch.cat([rgb, a], 1)                                                 # (S, 4, h, w)

    rgba_t, rgba_tp1 = rgba(results_t), rgba(results_tp1)
    static_rgb = results_t["static_rgbs_fine"].reshape(h, w, S, 3)
    static_a = results_t["static_alphas_fine"].reshape(h, w, S, 1)
    zs = results_t["zs_fine"].reshape(h, w, S)
    acc = torch.zeros(h, w, 4)
    depth = torch.zeros(h, w)
    for s in range(S):
        fw = splat_average(rgba_t[s:s + 1], of_fw[s:s + 1])[0].permute(1, 2, 0)
        bw = splat_average(rgba_tp1[s:s + 1], of_bw[s:s + 1])[0].permute(1, 2, 0)
        c_rgb = fw[..., :3] * fw[..., 3:] * (1 - dt) + bw[..., :3] * bw[..., 3:] * dt + static_rgb[:, :, s] * static_a[:, :, s]
        c_a = 1 - (1 - (fw[..., 3:] * (1 - dt) + bw[..., 3:] * dt)) * (1 - static_a[:, :, s])
        acc[..., :3] += (1 - acc[..., 3:]) * c_rgb
        depth += (1 - acc[..., 3]) * c_a[..., 0] * zs[..., s]
        acc[..., 3:] += (1 - acc[..., 3:]) * c_a
    return acc[..., :3], depth


# --------------------------------------------------------------------------
# ray generation for a test frame (datasets/ray_utils.py:7-106, monocular.py:268-276)
# --------------------------------------------------------------------------
def frame_rays(H: int, W: int, K: torch.Tensor, c2w: torch.Tensor, near: float = 1.0) -> torch.Tensor:
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    dirs = torch.stack([(xs - cx) / fx, -(ys - cy) / fy, -torch.ones_like(xs)], -1).reshape(-1, 3)   # :27-29
    rays_d = dirs @ c2w[:, :3].T                                                                       # :50-54
    rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    rays_o = c2w[:, 3].expand_as(rays_d)
    shift_near = -min(-1.0, float(c2w[2, 3]))                                                          # monocular.py:270
    t = -(shift_near + rays_o[..., 2]) / rays_d[..., 2]                                                # :88-104
    rays_o = rays_o + t[..., None] * rays_d
    ox_oz, oy_oz = rays_o[..., 0] / rays_o[..., 2], rays_o[..., 1] / rays_o[..., 2]
    o0, o1 = -1. / (cx / fx) * ox_oz, -1. / (cy / fy) * oy_oz
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (cx / fx) * (rays_d[..., 0] / rays_d[..., 2] - ox_oz)
    d1 = -1. / (cy / fy) * (rays_d[..., 1] / rays_d[..., 2] - oy_oz)
    return torch.cat([torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, 1 - o2], -1)], 1)
