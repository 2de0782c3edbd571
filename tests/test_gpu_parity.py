"""GPU parity tests: nsff_pl_b200 (CUDA, through the C ABI) vs the CPU oracle on identical seeded inputs.

Tolerances (stated; every bound is abs error <= tol * max(1, |ref|_max)):
  * precision='fp32' (CUDA-core verification path): 2e-5 on first-order results; 2e-4 on the warp-derived
    tensors (rgb_fw/bw, xyzs_fw_bw/bw_fw, disocc*), whose inputs pass through sin(512*(x+flow)) a second
    time and amplify fp32 summation-order differences.  Hierarchical mode: the resampled depths are an
    inverse CDF, ill-conditioned where the coarse pdf is ~0, so per-sample tensors of the fine pass sit at
    slightly different depths; the coarse pass and the per-ray outputs are still held to 2e-4.
  * precision='bf16' (tcgen05 path; bf16 operands, fp32 accumulate), DEFAULT-INIT weights - the state the
    BASELINE.md section-4 envelope was measured on: rgb <= 5e-3 and PSNR >= 50 dB, depth <= 1e-3,
    flow <= 1e-4 (north_star's "stated float tolerance"; measured here: 9e-5 / 90 dB / 1e-5 / 4e-5).
  * precision='bf16', ENERGISED weights (tests/scenes.py: gains that make densities/flows vary strongly and
    amplify bf16 rounding ~30x): per-ray rgb/depth/flow <= 1.5e-2, PSNR >= 50 dB, per-sample heads <= 1e-2,
    sigmas <= 3e-2 relative, second-order cycle points xyzs_fw_bw/bw_fw <= 0.1.
"""
import numpy as np
import pytest
import torch

import golden_io
from parity_util import build_fields, compare_dicts, max_abs, psnr, synthetic_rays
from scenes import FakeDataset
from oracle import nsff_oracle as O

pytestmark = pytest.mark.gpu

from tolerances import BF16_DEFAULT_INIT_TOL, BF16_TOL, FP32_TOL, FP32_WARP_TOL, PER_RAY_KEYS  # noqa: E402


def _run_both(models, emb, specs, enc, rays, ts, max_t, **kw):
    from nsff_pl_b200.models.rendering import render_rays
    dev = "cuda"
    kw_gpu = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
    torch.manual_seed(1234)
    got = render_rays(models, emb, rays.to(dev), None if ts is None else ts.to(dev), max_t, **kw_gpu)
    torch.cuda.synchronize()
    torch.manual_seed(1234)
    want = O.render_rays(specs, enc, rays, ts, max_t, **kw)
    return got, want


def _compare_hierarchical(got, want, coarse_tol, ray_tol, zs_tol):
    """Coarse pass + resampled depths + per-ray fine outputs; per-sample fine tensors live at (slightly)
    different depths on the two sides, so only their shapes/finiteness are checked."""
    assert list(got.keys()) == list(want.keys())
    tol = {"_coarse": coarse_tol, "zs_fine": zs_tol}
    tol.update({k: ray_tol for k in PER_RAY_KEYS})
    tol["rgb_fw"] = tol["rgb_bw"] = ray_tol
    compare_dicts(got, want, tol, default_tol=None)


# ------------------------------------------------------------------------------------------------
# small ops
# ------------------------------------------------------------------------------------------------
def test_library_loaded_and_device_is_blackwell():
    from nsff_pl_b200 import _cabi
    assert _cabi.lib().nsff_abi_version() == 1
    assert torch.cuda.get_device_capability()[0] == 10, "these kernels are sm_100a only"


def test_pos_embedding_matches_oracle():
    from nsff_pl_b200.models.nerf import PosEmbedding
    torch.manual_seed(0)
    x = torch.randn(1000, 3)
    for ml, n in ((9, 10), (3, 4)):
        got = PosEmbedding(ml, n)(x.cuda())
        want = O.pos_embedding(x, O.pos_freqs(ml, n))
        assert got.shape == want.shape
        assert max_abs(got, want) <= 2e-6 * 512      # sin/cos of |arg| up to ~2000: few-ulp argument error
    assert PosEmbedding(9, 10)(torch.zeros(0, 3).cuda()).shape == (0, 63)


def test_sample_pdf_matches_oracle():
    from nsff_pl_b200.models.rendering import sample_pdf
    meta, a = golden_io.load("ops_w64")
    got = sample_pdf(a["pdf.bins"].cuda(), a["pdf.weights"].cuda(), meta["n_importance"], det=True)
    assert max_abs(got, a["pdf.det"]) <= 2e-6
    # random u: both sides consume their own generator, so compare through the oracle with the SAME u
    torch.manual_seed(3)
    bins, w = a["pdf.bins"], a["pdf.weights"]
    u = torch.rand(bins.shape[0], 33)
    import nsff_pl_b200._cabi as cabi
    out = torch.empty(bins.shape[0], 33, device="cuda")
    bc, wc, uc = bins.cuda(), w.cuda(), u.cuda()
    cabi.check(cabi.lib().nsff_sample_pdf(cabi.ptr(bc), bc.stride(0), cabi.ptr(wc), wc.stride(0), bins.shape[0],
                                          w.shape[1], 33, cabi.ptr(uc), 1e-5, cabi.ptr(out), cabi.stream_ptr()),
               "nsff_sample_pdf")
    torch.manual_seed(0)
    real_rand = torch.rand
    try:
        torch.rand = lambda *s, **k: u
        want = O.sample_pdf(bins, w, 33, det=False)
    finally:
        torch.rand = real_rand
    assert max_abs(out, want) <= 2e-6


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
def test_nerf_forward_modes(precision, tol):
    """NeRF.forward drop-in on embedded rows, all call modes of nerf.py:129-141."""
    models, emb, specs, enc = build_fields(W=256, coarse=True, seed=21, precision=precision)
    torch.manual_seed(4)
    x3 = torch.rand(300, 3) * 2 - 1
    xin = torch.cat([O.pos_embedding(x3, enc.xyz_freqs), O.pos_embedding(torch.randn(300, 3) * 0.3, enc.dir_freqs),
                     torch.randn(300, 48)], 1)
    f, c = models["fine"], models["coarse"]
    sf, sc = specs["fine"], specs["coarse"]
    xs = torch.cat([xin[:, :63], xin[:, 90:]], 1)
    cases = [
        (f(xin.cuda(), output_transient_flow=["fw", "bw"]), O.nerf_forward(sf, xin, output_transient_flow=["fw", "bw"])),
        (f(xin.cuda()), O.nerf_forward(sf, xin)),
        (f(xin[:, :90].cuda(), output_transient=False), O.nerf_forward(sf, xin[:, :90], output_transient=False)),
        (f(xin.cuda(), output_static=False, output_transient_flow=["bw"]),
         O.nerf_forward(sf, xin, output_static=False, output_transient_flow=["bw"])),
        (f(xin.cuda(), output_static=False, output_transient_flow=["fw"]),
         O.nerf_forward(sf, xin, output_static=False, output_transient_flow=["fw"])),
        (c(xs.cuda(), sigma_only=True), O.nerf_forward(sc, xs, sigma_only=True)),
        (c(xin[:, :63].cuda(), sigma_only=True, output_transient=False),
         O.nerf_forward(sc, xin[:, :63], sigma_only=True, output_transient=False)),
        (c(xin.cuda()), O.nerf_forward(sc, xin)),
    ]
    for i, (got, want) in enumerate(cases):
        assert got.shape == want.shape, f"mode {i}: {got.shape} vs {want.shape}"
        scale = max(1.0, float(want.abs().max()))
        assert max_abs(got, want) <= tol * scale, f"mode {i}: err {max_abs(got, want):.3e} (scale {scale:.2f})"


# ------------------------------------------------------------------------------------------------
# render_rays, fp32 verification path: tight tolerance, every mode, ragged sizes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R,S", [(40, 128), (7, 50), (1, 2), (33, 192)])
def test_render_train_warp_fp32(R, S):
    models, emb, specs, enc = build_fields(W=256, seed=5, precision="fp32", n_frames=8)
    rays, ts = synthetic_rays(R, n_frames=8, seed=R)
    got, want = _run_both(models, emb, specs, enc, rays, ts, 7, N_samples=S, perturb=0, noise_std=0,
                          N_importance=0, test_time=False, output_transient_flow=["fw", "bw", "disocc"])
    compare_dicts(got, want, FP32_WARP_TOL, default_tol=FP32_TOL)


def test_render_modes_fp32():
    models, emb, specs, enc = build_fields(W=256, coarse=True, seed=6, precision="fp32", n_frames=8)
    rays, ts = synthetic_rays(24, n_frames=8, seed=3)
    base = dict(N_samples=32, perturb=0, noise_std=0)
    for kw in (dict(N_importance=0, test_time=True, output_transient_flow=["fw", "bw"]),
               dict(N_importance=0, test_time=True),
               dict(N_importance=0, test_time=True, output_transient=False),
               dict(N_importance=0, test_time=False),
               dict(N_importance=16, test_time=True, output_transient_flow=["fw", "bw"]),
               dict(N_importance=8, test_time=False, output_transient_flow=["fw", "bw", "disocc"]),
               dict(N_importance=0, test_time=True, output_transient_flow=["fw", "bw"],
                    view_dir=torch.randn(24, 3) * 0.2, t_embedded=torch.randn(24, 48))):
        got, want = _run_both(models, emb, specs, enc, rays, ts, 7, **base, **kw)
        if kw["N_importance"]:
            _compare_hierarchical(got, want, coarse_tol=FP32_TOL, ray_tol=1e-3, zs_tol=1e-3)
        else:
            compare_dicts(got, want, FP32_WARP_TOL, default_tol=FP32_TOL)


def test_render_noise_and_perturb_fp32():
    """perturb / noise tensors are drawn by torch on each side's own generator: feed identical draws."""
    from nsff_pl_b200.models import rendering as R_
    models, emb, specs, enc = build_fields(W=256, seed=7, precision="fp32", n_frames=8)
    rays, ts = synthetic_rays(16, n_frames=8, seed=9)
    torch.manual_seed(77)
    draws = [torch.rand(16, 32)] + [torch.randn(16, 32) for _ in range(4)]
    it_cpu, it_gpu = iter(draws), iter(draws)
    real_rand, real_randn, real_randn_like = torch.rand, torch.randn, torch.randn_like
    try:
        torch.rand = lambda *s, **k: next(it_gpu).to(k.get("device", "cpu"))
        torch.randn = lambda *s, **k: next(it_gpu).to(k.get("device", "cpu"))
        got = R_.render_rays(models, emb, rays.cuda(), ts.cuda(), 7, N_samples=32, perturb=1.0, noise_std=1.0,
                             test_time=False, output_transient_flow=["fw", "bw", "disocc"])
        torch.cuda.synchronize()
        torch.rand = lambda *s, **k: next(it_cpu)
        torch.randn_like = lambda t, **k: next(it_cpu)
        want = O.render_rays(specs, enc, rays, ts, 7, N_samples=32, perturb=1.0, noise_std=1.0, test_time=False,
                             output_transient_flow=["fw", "bw", "disocc"])
    finally:
        torch.rand, torch.randn, torch.randn_like = real_rand, real_randn, real_randn_like
    compare_dicts(got, want, FP32_WARP_TOL, default_tol=FP32_TOL)


def test_render_static_only_and_visibility_fp32():
    models, emb, specs, enc = build_fields(W=256, transient=False, seed=8, precision="fp32")
    rays, _ = synthetic_rays(256, seed=1)
    got, want = _run_both(models, emb, specs, enc, rays, None, 0, N_samples=64, test_time=True)   # config 1 shape
    compare_dicts(got, want, FP32_WARP_TOL, default_tol=FP32_TOL)
    models, emb, specs, enc = build_fields(W=256, seed=9, precision="fp32", n_frames=6)
    rays, ts = synthetic_rays(24, n_frames=6, seed=11)
    ts = torch.full_like(ts, 2)
    ds = FakeDataset(n_frames=6, n_cams=2, img_wh=(32, 18), seed=6)
    got, want = _run_both(models, emb, specs, enc, rays, ts, 5, N_samples=32, test_time=True,
                          output_transient_flow=["fw", "bw"], dataset=ds)
    compare_dicts(got, want, FP32_WARP_TOL, default_tol=FP32_TOL)


def test_golden_w64_cases_through_fp32_path():
    """The reference-generated fixtures (W=64) replayed through the CUDA fp32 path."""
    from nsff_pl_b200.models.nerf import NeRF, PosEmbedding
    from nsff_pl_b200.models.rendering import render_rays
    for case in ("train_warp_w64", "test_flow_w64", "cfg1_static_w64", "hier_test_w64", "appearance_train_w64"):
        meta, a = golden_io.load(case)
        models = {}
        for key, m in meta["models"].items():
            net = NeRF(m["typ"], D=m["D"], W=m["W"], skips=m["skips"], in_channels_xyz=63, in_channels_dir=27,
                       use_viewdir=m["use_viewdir"], encode_appearance=m["encode_appearance"], in_channels_a=48,
                       encode_transient=m["encode_transient"], in_channels_t=48, output_flow=m["output_flow"])
            net.load_state_dict(golden_io.model_state(a, key))
            net.requires_grad_(False)
            net.precision = "fp32"
            models[key] = net.cuda()
        emb = {"xyz": PosEmbedding(9, 10), "dir": PosEmbedding(3, 4)}
        for k in ("t", "a"):
            if f"emb.{k}" in a:
                e = torch.nn.Embedding(*a[f"emb.{k}"].shape)
                e.weight.data.copy_(a[f"emb.{k}"])
                emb[k] = e.cuda().requires_grad_(False)
        kw = dict(meta["call"])
        max_t = kw.pop("max_t")
        ts = a["ts"].cuda() if "ts" in a else None
        got = render_rays(models, emb, a["rays"].cuda(), ts, max_t, **kw)
        assert list(got.keys()) == meta["result_keys"], case
        want = golden_io.expected(a)
        if kw["N_importance"]:
            _compare_hierarchical({k: got[k] for k in want}, want, coarse_tol=FP32_TOL, ray_tol=1e-3, zs_tol=1e-3)
        else:
            compare_dicts(got, want, FP32_WARP_TOL, default_tol=FP32_TOL, check_order=False)


# ------------------------------------------------------------------------------------------------
# render_rays, tcgen05 bf16 path
# ------------------------------------------------------------------------------------------------
def test_render_train_warp_bf16_default_init_reference_tolerance():
    """north_star's parity statement: rendered RGB, depth and flow within the stated tolerance of the
    reference path, on the reference's own default-initialised networks, full warp."""
    import parity_util
    real = parity_util.energise_
    parity_util.energise_ = lambda m, **k: m
    try:
        models, emb, specs, enc = build_fields(W=256, seed=5, precision="bf16", n_frames=8)
    finally:
        parity_util.energise_ = real
    rays, ts = synthetic_rays(128, n_frames=8, seed=64)
    got, want = _run_both(models, emb, specs, enc, rays, ts, 7, N_samples=128, perturb=0, noise_std=0,
                          N_importance=0, test_time=False, output_transient_flow=["fw", "bw", "disocc"])
    errs = compare_dicts(got, want, BF16_DEFAULT_INIT_TOL, default_tol=None)
    assert psnr(got["rgb_fine"], want["rgb_fine"]) >= 50.0, errs


def test_render_hierarchical_bf16():
    models, emb, specs, enc = build_fields(W=256, coarse=True, seed=6, precision="bf16", n_frames=8)
    rays, ts = synthetic_rays(64, n_frames=8, seed=3)
    got, want = _run_both(models, emb, specs, enc, rays, ts, 7, N_samples=64, N_importance=64, test_time=True,
                          output_transient_flow=["fw", "bw"])
    assert got["zs_fine"].shape == (64, 192)                       # 64 + 2*64 fine samples
    assert bool((got["zs_fine"][:, 1:] >= got["zs_fine"][:, :-1]).all()), "merged depths must be sorted"
    _compare_hierarchical(got, want, coarse_tol=3e-2, ray_tol=1.5e-2, zs_tol=None)
    assert psnr(got["rgb_fine"], want["rgb_fine"]) >= 45.0
    # Inverse-CDF resampling is ill-conditioned where the coarse pdf ~ 0, so the bf16 path's resampled depths differ in
    # a few places and per-sample fine tensors cannot be compared index by index above.  Re-run the oracle's fine pass on
    # the depths the GPU actually used (oracle test hook `_zs_fine`): now EVERY fine-pass tensor is comparable.
    want2 = O.render_rays(specs, enc, rays, ts, 7, N_samples=64, N_importance=64, test_time=True,
                          output_transient_flow=["fw", "bw"], _zs_fine=got["zs_fine"].cpu())
    fine = {k: v for k, v in got.items() if k.endswith("_fine") and not k.endswith("zs_fine")}
    compare_dicts(fine, {k: want2[k] for k in fine}, BF16_TOL, default_tol=None, check_order=False)


@pytest.mark.parametrize("R,S", [(64, 128), (5, 77), (300, 128)])
def test_render_train_warp_bf16(R, S):
    models, emb, specs, enc = build_fields(W=256, seed=5, precision="bf16", n_frames=8)
    rays, ts = synthetic_rays(R, n_frames=8, seed=R)
    got, want = _run_both(models, emb, specs, enc, rays, ts, 7, N_samples=S, perturb=0, noise_std=0,
                          N_importance=0, test_time=False, output_transient_flow=["fw", "bw", "disocc"])
    errs = compare_dicts(got, want, BF16_TOL, default_tol=None)
    assert psnr(got["rgb_fine"], want["rgb_fine"]) >= 50.0, errs
    assert psnr(got["rgb_fw"], want["rgb_fw"]) >= 50.0, errs


def test_bf16_path_matches_bf16_emulation_tightly():
    """Catches layout/packing bugs that a loose bf16 tolerance would hide: emulate the kernel's rounding
    points (bf16 inputs/weights/hidden activations, fp32 accumulate) on the CPU and compare at 5e-3 relative (a bf16 ulp: rounding
    boundaries flip with accumulation order; layout or packing bugs give O(1) errors)."""
    models, emb, specs, enc = build_fields(W=256, seed=13, precision="bf16")
    f, sf = models["fine"], specs["fine"]
    torch.manual_seed(2)
    x3 = torch.rand(512, 3) * 2 - 1
    xin = torch.cat([O.pos_embedding(x3, enc.xyz_freqs), O.pos_embedding(torch.randn(512, 3) * 0.3, enc.dir_freqs),
                     torch.randn(512, 48)], 1)
    got = f(xin.cuda(), output_transient_flow=["fw", "bw"]).cpu()

    def r(t):
        return t.to(torch.bfloat16).to(torch.float32)

    p = {k: (r(v) if k.endswith("weight") else v) for k, v in sf.params.items()}

    def lin(name, x):
        return x @ p[name + ".weight"].T + p[name + ".bias"]

    def trunk(prefix, skip_in):
        h = skip_in
        for i in range(8):
            if i == 4:
                h = torch.cat([skip_in, h], 1)
            h = r(torch.relu(lin(f"{prefix}_xyz_encoding_{i + 1}.0", h)))
        return h

    x = r(xin)
    h = trunk("static", x[:, :63])
    s_sigma = lin("static_sigma", h)
    feat = r(lin("static_xyz_encoding_final", h))
    feat = r(torch.relu(lin("static_dir_encoding.0", torch.cat([feat, x[:, 63:90]], 1))))
    s_rgb = torch.sigmoid(lin("static_rgb.0", feat))
    ht = trunk("transient", torch.cat([x[:, :63], x[:, 90:]], 1))
    ft = r(lin("transient_xyz_encoding_final", ht))
    want = torch.cat([s_rgb, s_sigma, torch.sigmoid(lin("transient_rgb.0", ft)), lin("transient_sigma", ft),
                      0.2 * torch.tanh(lin("transient_flow_fw.0", ft)), 0.2 * torch.tanh(lin("transient_flow_bw.0", ft))], 1)
    scale = max(1.0, float(want.abs().max()))
    assert max_abs(got, want) <= 5e-3 * scale, f"err {max_abs(got, want):.3e} scale {scale:.2f}"


def test_full_frame_properties_bf16():
    """Size-independent checks at the benchmark shape's per-ray size (S=128) on a larger ray batch:
    weights are a sub-stochastic partition, static-only == joint when the dynamic field is switched off,
    rays are independent (a shuffled batch renders to the shuffled result)."""
    from nsff_pl_b200.models.rendering import render_rays
    models, emb, specs, enc = build_fields(W=256, seed=17, precision="bf16")
    rays, ts = synthetic_rays(4096, n_frames=30, seed=2)
    rays, ts = rays.cuda(), ts.cuda()
    kw = dict(N_samples=128, test_time=False, output_transient_flow=["fw", "bw", "disocc"])
    out = render_rays(models, emb, rays, ts, 29, **kw)
    w = out["weights_fine"]
    assert torch.isfinite(w).all() and (w >= 0).all() and (w.sum(1) <= 1 + 1e-4).all()
    assert (out["transient_flows_fw"][out["zs_fine"] > 0.95] == 0).all()
    perm = torch.randperm(4096, device="cuda")
    out_p = render_rays(models, emb, rays[perm].contiguous(), ts[perm].contiguous(), 29, **kw)
    for k in ("rgb_fine", "depth_fine", "rgb_fw", "disocc_bw", "weights_fine"):
        assert torch.equal(out_p[k], out[k][perm]), f"{k}: rays are not independent / not deterministic"


def test_w256_golden_fixture_bf16_and_fp32():
    from nsff_pl_b200.models.nerf import NeRF, PosEmbedding
    from nsff_pl_b200.models.rendering import render_rays
    from scenes import energise_
    meta, a = golden_io.load("train_warp_w256_seeded")
    torch.manual_seed(meta["build_seed"])
    fine = NeRF("fine", D=8, W=256, skips=[4], in_channels_xyz=63, in_channels_dir=27, use_viewdir=True,
                encode_appearance=False, in_channels_a=48, encode_transient=True, in_channels_t=48,
                output_flow=True, flow_scale=0.2)
    t_emb = torch.nn.Embedding(meta["n_frames"], 48)
    energise_(fine)
    csum = float(sum(v.double().abs().sum() for v in fine.state_dict().values()))
    if abs(csum - meta["weight_abs_sum"]) > 1e-6 * meta["weight_abs_sum"]:
        pytest.skip("torch RNG differs from the build container")
    fine.requires_grad_(False)
    fine.cuda()
    emb = {"xyz": PosEmbedding(9, 10), "dir": PosEmbedding(3, 4), "t": t_emb.cuda().requires_grad_(False)}
    kw = dict(meta["call"])
    max_t = kw.pop("max_t")
    for precision, tol, default in (("fp32", FP32_WARP_TOL, FP32_TOL), ("bf16", BF16_TOL, None)):
        fine.precision = precision
        got = render_rays({"fine": fine}, emb, a["rays"].cuda(), a["ts"].cuda(), max_t, **kw)
        assert list(got.keys()) == meta["result_keys"]
        compare_dicts(got, golden_io.expected(a), tol, default_tol=default, check_order=False)


def test_cpu_tensors_fail_loudly_and_grad_modes():
    from nsff_pl_b200.models.rendering import render_rays
    models, emb, specs, enc = build_fields(W=256, seed=1)
    rays, ts = synthetic_rays(8, seed=0)
    with pytest.raises(NotImplementedError):
        render_rays(models, emb, rays, ts.cuda(), 29, N_samples=8)            # CPU rays: no fallback
    models["fine"].requires_grad_(True)
    with pytest.raises(NotImplementedError):                                  # test_time is an inference mode
        render_rays(models, emb, rays.cuda(), ts.cuda(), 29, N_samples=8, test_time=True)
    out = render_rays(models, emb, rays.cuda(), ts.cuda(), 29, N_samples=8)   # training mode: differentiable
    assert out["rgb_fine"].requires_grad and out["rgb_fine"].shape == (8, 3)
    with torch.no_grad():
        out = render_rays(models, emb, rays.cuda(), ts.cuda(), 29, N_samples=8)
    assert not out["rgb_fine"].requires_grad


def test_ray_sharded_render_is_bitwise_identical_to_unsharded():
    """Config 4's invariant: rays are independent, so rendering a frame in contiguous
    ray blocks (what each rank of an N-GPU job does, nsff_pl_b200/parallel.py) and concatenating equals the
    one-shot render bit for bit - for uneven block sizes too."""
    from nsff_pl_b200 import parallel
    from nsff_pl_b200.models.rendering import render_rays
    models, emb, _, _ = build_fields(W=256, seed=23, precision="bf16")
    rays, ts = synthetic_rays(1000, n_frames=30, seed=8)
    rays, ts = rays.cuda(), ts.cuda()
    kw = dict(N_samples=128, test_time=False, output_transient_flow=["fw", "bw", "disocc"])
    full = render_rays(models, emb, rays, ts, 29, **kw)
    for world in (2, 8, 3):
        parts = []
        for rank in range(world):
            lo, hi = parallel.shard_bounds(1000, rank, world)
            parts.append(parallel.pack_pixels(render_rays(models, emb, rays[lo:hi].contiguous(), ts[lo:hi].contiguous(),
                                                          29, **kw)))
        assert torch.equal(torch.cat(parts, 0), parallel.pack_pixels(full)), f"world={world}"


def test_interpolate_matches_oracle_and_reference_fixture():
    """Section 8f #1: test-time interpolation + 'average' splat, one launch for all planes.  (i) the reference-generated
    fixture (per-sample tensors from the reference's render_rays) replayed through the CUDA path; (ii) end to end
    from our own render_rays(test_time=True) views (strided inputs), vs the oracle on the same results.
    Tolerance 2e-5: atomics make the splat accumulation order non-deterministic at the 1e-7 level."""
    from nsff_pl_b200.models.rendering import interpolate, render_rays
    # interpolate_cuda_w64: the unmodified reference interpolate() INCLUDING its own CUDA splat kernel, run on a B200
    # (tests/golden/make_splat_golden.py); interpolate_w64: the same with the splat op restated on the CPU (round 1)
    for fixture in ("interpolate_cuda_w64", "interpolate_w64"):
        meta, a = golden_io.load(fixture)
        rt = {k: a[f"t.{k}"].cuda() for k in meta["keys"]}
        rp = {k: a[f"tp1.{k}"].cuda() for k in meta["keys"]}
        for dt in meta["dts"]:
            rgb, depth = interpolate(rt, rp, dt, a["K"], a["c2w"], tuple(meta["img_wh"]))
            assert max_abs(rgb, a[f"out.rgb.{dt}"]) <= 2e-5 and max_abs(depth, a[f"out.depth.{dt}"]) <= 2e-5, (fixture, dt)
    # end to end at W=256 on the tensor-core path
    models, emb, specs, enc = build_fields(W=256, seed=41, precision="bf16", n_frames=8)
    w, h = 24, 16
    rays, _ = synthetic_rays(w * h, n_frames=8, seed=3, img_wh=(w, h), frame=2)
    ts = torch.full((w * h,), 2, dtype=torch.long)
    kw = dict(N_samples=32, test_time=True, output_transient_flow=["fw", "bw"])
    gt = render_rays(models, emb, rays.cuda(), ts.cuda(), 7, **kw)
    gp = render_rays(models, emb, rays.cuda(), (ts + 1).cuda(), 7, **kw)
    ds = FakeDataset(n_frames=8, n_cams=1, img_wh=(w, h), seed=2)
    K, c2w = ds.Ks[0], torch.tensor(ds.poses[0])
    rgb, depth = interpolate(gt, gp, 0.4, K, c2w, (w, h))
    cpu = lambda d: {k: v.detach().cpu() for k, v in d.items()}
    rgb_o, depth_o = O.interpolate(cpu(gt), cpu(gp), 0.4, K, c2w, (w, h))
    assert rgb.shape == (h, w, 3) and depth.shape == (h, w)
    assert max_abs(rgb, rgb_o) <= 2e-5 and max_abs(depth, depth_o) <= 2e-5


def test_render_appearance_code_bf16_and_fp32():
    """NeRF-W appearance code (encode_appearance=True, nerf.py:85-87): a 48-wide per-frame code joins the dir layer;
    on the tensor-core path it is a fourth prologue chunk."""
    for precision, tol, default in (("fp32", FP32_WARP_TOL, FP32_TOL), ("bf16", BF16_TOL, None)):
        models, emb, specs, enc = build_fields(W=256, appearance=True, seed=19, precision=precision, n_frames=8)
        rays, ts = synthetic_rays(40, n_frames=8, seed=6)
        got, want = _run_both(models, emb, specs, enc, rays, ts, 7, N_samples=48, test_time=False,
                              output_transient_flow=["fw", "bw", "disocc"])
        compare_dicts(got, want, tol, default_tol=default)
        got, want = _run_both(models, emb, specs, enc, rays, ts, 7, N_samples=48, test_time=True,
                              a_embedded=torch.randn(40, 48))
        compare_dicts(got, want, tol, default_tol=default)


@pytest.mark.gpu
def test_frame_rays_on_device_match_reference_fixture_and_oracle():
    """f4: rays built on the device (one kernel) vs the reference's CPU ray_utils output (fixture) and, at the
    benchmark frame size, vs the oracle; the NDC invariant o_z + d_z == 1 holds for every ray."""
    from nsff_pl_b200.datasets import ray_utils as RU
    meta, a = golden_io.load("frame_rays")
    w, h = meta["img_wh"]
    K = torch.as_tensor(a["K"])
    for i in range(meta["n"]):
        got = RU.frame_rays(h, w, K, a[f"c2w.{i}"], near=meta["near"]).cpu()
        ref = torch.as_tensor(a[f"rays.{i}"])
        assert max_abs(got, ref) <= 2e-6 * float(ref.abs().max()), (i, max_abs(got, ref))
    ds = FakeDataset(n_frames=4, n_cams=1, img_wh=(512, 288), seed=3)
    c2w = torch.tensor(ds.poses[1])
    got = RU.frame_rays(288, 512, ds.Ks[0], c2w)
    ref = O.frame_rays(288, 512, ds.Ks[0], c2w)
    assert got.shape == (288 * 512, 6)
    assert max_abs(got.cpu(), ref) <= 2e-6 * float(ref.abs().max())
    assert float((got[:, 2] + got[:, 5] - 1).abs().max()) < 1e-6
    # get_ndc_rays alone on arbitrary world rays
    g = torch.Generator().manual_seed(4)
    ro = torch.randn(1000, 3, generator=g) * 0.1
    rd = torch.nn.functional.normalize(torch.randn(1000, 3, generator=g) * 0.3 + torch.tensor([0., 0., -1.]), dim=-1)
    o_d, d_d = RU.get_ndc_rays(ds.Ks[0], 1.0, 1.0, ro.cuda(), rd.cuda())
    fx, fy, cx, cy = ds.Ks[0][0, 0], ds.Ks[0][1, 1], ds.Ks[0][0, 2], ds.Ks[0][1, 2]
    t = -(1.0 + ro[:, 2]) / rd[:, 2]
    o = ro + t[:, None] * rd
    o_ref = torch.stack([-1 / (cx / fx) * o[:, 0] / o[:, 2], -1 / (cy / fy) * o[:, 1] / o[:, 2], 1 + 2 / o[:, 2]], -1)
    assert max_abs(o_d.cpu(), o_ref) < 1e-5
    assert float((o_d[:, 2] + d_d[:, 2] - 1).abs().max()) < 1e-6


@pytest.mark.gpu
def test_empty_ray_batch_returns_empty_results_with_the_same_keys():
    """Edge case: R = 0 (a ragged last chunk of eval.f can be empty).  No kernel launch may fail; every result keeps its
    key, dtype and trailing shape with a leading dimension of 0."""
    from nsff_pl_b200.models.rendering import render_rays
    models, emb, _, _ = build_fields(W=256, seed=2, precision="bf16", n_frames=8)
    rays, ts = synthetic_rays(3, n_frames=8, seed=7)
    for kw in (dict(N_samples=16, test_time=False, output_transient_flow=["fw", "bw", "disocc"]),
               dict(N_samples=16, test_time=True, output_transient_flow=["fw", "bw"])):
        with torch.no_grad():
            one = render_rays(models, emb, rays[:1].cuda(), ts[:1].cuda(), 7, **kw)
            none = render_rays(models, emb, rays[:0].cuda(), ts[:0].cuda(), 7, **kw)
        torch.cuda.synchronize()
        assert list(one.keys()) == list(none.keys())
        for k in one:
            assert none[k].shape == (0,) + tuple(one[k].shape[1:]), k
            assert none[k].dtype == one[k].dtype


# ---------------------------------------------------------------------------------------------------------------------
# composite: the staged kernel (bulk copies into per-warp shared-memory stages) and the direct kernel share one
# per-sample body, so their results must be IDENTICAL bit for bit - on every mode and on ragged chunk counts.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_samples", [128, 40, 24, 4])
@pytest.mark.parametrize("mode", ["static", "transient", "flow_test", "warp_train", "warp_noise"])
def test_staged_composite_is_bit_identical_to_the_direct_kernel(monkeypatch, mode, n_samples):
    from nsff_pl_b200.models.rendering import render_rays
    transient = mode != "static"
    models, emb, _, _ = build_fields(W=256, transient=transient, seed=41, precision="bf16", n_frames=8)
    rays, ts = synthetic_rays(333, n_frames=8, seed=42)
    kw = dict(N_samples=n_samples, perturb=0, noise_std=0, test_time=True)
    if mode == "transient":
        kw.update(output_transient=True, output_transient_flow=[])
    elif mode == "flow_test":                      # test_time: flows composited, no warp passes
        kw.update(output_transient_flow=["fw", "bw"])
    elif mode == "warp_train":
        kw.update(output_transient_flow=["fw", "bw", "disocc"], test_time=False)
    elif mode == "warp_noise":
        kw.update(output_transient_flow=["fw", "bw", "disocc"], test_time=False, perturb=1.0, noise_std=1.0)
    outs = []
    for direct in ("0", "1"):
        monkeypatch.setenv("NSFF_COMP_DIRECT", direct)
        torch.manual_seed(7)
        with torch.no_grad():
            outs.append(render_rays(models, emb, rays.cuda(), ts.cuda() if transient else None, 7, **kw))
    assert list(outs[0]) == list(outs[1])
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_merged_n256_ops_equal_the_two_half_ops(monkeypatch):
    """field kernel: the K-chunks both halves of a layer read first are issued as N = 256 MMAs into the two adjacent half
    accumulators (NSFF_TC_MERGE=1, default) instead of two N = 128 MMAs: same products, same accumulation order per
    accumulator element - the heads must agree to rounding noise (measured: bit-identical)."""
    from nsff_pl_b200.models.rendering import render_rays
    models, emb, _, _ = build_fields(W=256, seed=51, precision="bf16", n_frames=8)
    rays, ts = synthetic_rays(300, n_frames=8, seed=52)
    kw = dict(N_samples=64, perturb=0, noise_std=0, test_time=False, output_transient_flow=["fw", "bw", "disocc"])
    outs = []
    for merge in ("1", "0"):
        monkeypatch.setenv("NSFF_TC_MERGE", merge)
        with torch.no_grad():
            outs.append(render_rays(models, emb, rays.cuda(), ts.cuda(), 7, **kw))
    for k in outs[0]:
        assert max_abs(outs[0][k], outs[1][k]) <= 1e-5 * max(1.0, float(outs[1][k].abs().max())), k
