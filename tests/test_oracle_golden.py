"""Pins oracle/nsff_oracle.py against outputs of the UNMODIFIED reference (tests/golden/*.npz).

CPU only.  Tolerance: the oracle uses the same fp32 torch ops in the same order, so agreement is
expected to be bit-exact or within a few ulp; the assertion allows 2e-6 abs + 2e-6 rel.
"""
import numpy as np
import pytest
import torch

import golden_io
from scenes import FakeDataset, energise_
from oracle import nsff_oracle as O

ATOL = RTOL = 2e-6
RENDER_CASES = [c for c in golden_io.list_cases("*_w64")
                if c not in ("ops_w64", "interpolate_w64", "interpolate_cuda_w64") and not c.startswith("grads_")]
GRAD_CASES = golden_io.list_cases("grads_*")


def _spec(meta_m, state):
    return O.FieldSpec(typ=meta_m["typ"], params=state, D=meta_m["D"], W=meta_m["W"], skips=tuple(meta_m["skips"]),
                       in_channels_xyz=meta_m["in_channels_xyz"], use_viewdir=meta_m["use_viewdir"],
                       in_channels_dir=meta_m["in_channels_dir"], encode_appearance=meta_m["encode_appearance"],
                       in_channels_a=meta_m["in_channels_a"], encode_transient=meta_m["encode_transient"],
                       in_channels_t=meta_m["in_channels_t"], output_flow=meta_m["output_flow"],
                       flow_scale=meta_m["flow_scale"])


def _encoders(arrays):
    return O.Encoders(xyz_freqs=O.pos_freqs(9, 10), dir_freqs=O.pos_freqs(3, 4),
                      t_table=arrays.get("emb.t"), a_table=arrays.get("emb.a"))


def _close(name, got, want, tol=None):
    assert got.shape == want.shape, f"{name}: shape {tuple(got.shape)} != {tuple(want.shape)}"
    err = (got.double() - want.double()).abs()
    bound = (ATOL if tol is None else tol) + (RTOL if tol is None else tol) * want.double().abs()
    assert bool((err <= bound).all()), f"{name}: max abs err {err.max().item():.3e}"


def test_pos_embedding_and_forward_modes():
    meta, a = golden_io.load("ops_w64")
    x = a["pe.x"]
    _close("pe.xyz", O.pos_embedding(x, O.pos_freqs(9, 10)), a["pe.xyz"])
    _close("pe.dir", O.pos_embedding(x, O.pos_freqs(3, 4)), a["pe.dir"])
    fine = _spec(meta["models"]["fine"], golden_io.model_state(a, "fine"))
    coarse = _spec(meta["models"]["coarse"], golden_io.model_state(a, "coarse"))
    xin = a["fwd.x"]
    _close("full_flow", O.nerf_forward(fine, xin, output_transient_flow=["fw", "bw", "disocc"]), a["fwd.fine.full_flow"])
    _close("full_noflow", O.nerf_forward(fine, xin), a["fwd.fine.full_noflow"])
    _close("static", O.nerf_forward(fine, xin[:, :90], output_transient=False), a["fwd.fine.static"])
    _close("warp_fw", O.nerf_forward(fine, xin, output_static=False, output_transient_flow=["fw"]), a["fwd.fine.warp_fw"])
    _close("warp_bw", O.nerf_forward(fine, xin, output_static=False, output_transient_flow=["bw"]), a["fwd.fine.warp_bw"])
    xs = torch.cat([xin[:, :63], xin[:, 90:]], 1)
    _close("sigma_both", O.nerf_forward(coarse, xs, sigma_only=True), a["fwd.coarse.sigma_both"])
    _close("sigma_static", O.nerf_forward(coarse, xin[:, :63], sigma_only=True, output_transient=False),
           a["fwd.coarse.sigma_static"])
    _close("coarse_full", O.nerf_forward(coarse, xin), a["fwd.coarse.full"])


def test_sample_pdf():
    meta, a = golden_io.load("ops_w64")
    _close("pdf.det", O.sample_pdf(a["pdf.bins"], a["pdf.weights"], meta["n_importance"], det=True), a["pdf.det"])
    torch.manual_seed(meta["pdf_seed"])
    _close("pdf.rand", O.sample_pdf(a["pdf.bins"], a["pdf.weights"], meta["n_importance"], det=False), a["pdf.rand"])


@pytest.mark.parametrize("case", RENDER_CASES)
def test_render_rays_matches_reference(case):
    meta, a = golden_io.load(case)
    models = {k: _spec(m, golden_io.model_state(a, k)) for k, m in meta["models"].items()}
    kw = dict(meta["call"])
    max_t = kw.pop("max_t")
    for k in meta["tensor_kwargs"]:
        kw[k] = a[f"kw.{k}"]
    if meta["has_dataset"]:
        ds = FakeDataset(n_frames=meta["n_frames"], n_cams=2, img_wh=(32, 18), seed=6)
        assert np.allclose(ds.poses, a["ds.poses"].numpy())
        kw["dataset"] = ds
    torch.manual_seed(meta["seed"])
    got = O.render_rays(models, _encoders(a), a["rays"], a.get("ts"), max_t, **kw)
    want = golden_io.expected(a)
    assert list(got.keys()) == meta["result_keys"], "result dict keys / order differ from the reference"
    for k in want:
        _close(k, got[k], want[k])


@pytest.mark.parametrize("case", GRAD_CASES)
def test_oracle_autograd_matches_reference_autograd(case):
    """The oracle is also the gradient reference of tests/test_gpu_backward.py, so its autograd is pinned too:
    same differentiable-output set (disocc* detached, resampling pdf detached), same loss value, same gradients."""
    meta, a = golden_io.load(case)
    models = {k: _spec(m, golden_io.model_state(a, k)) for k, m in meta["models"].items()}
    for sp in models.values():
        for k in sp.params:
            sp.params[k] = sp.params[k].clone().requires_grad_(True)
    enc = _encoders(a)
    enc.t_table = enc.t_table.clone().requires_grad_(True)
    kw = dict(meta["call"])
    max_t = kw.pop("max_t")
    got = O.render_rays(models, enc, a["rays"], a["ts"], max_t, **kw)
    assert [k for k, v in got.items() if v.requires_grad] == meta["grad_keys"]
    g = torch.Generator().manual_seed(meta["loss_seed"])
    loss = 0.0
    for k, v in got.items():
        w = torch.randn(v.shape, generator=g)
        if v.requires_grad:
            loss = loss + (v * w).sum()
    assert abs(float(loss) - meta["loss"]) <= 1e-4 * max(1.0, abs(meta["loss"]))
    loss.backward()
    for key, sp in models.items():
        for n, p in sp.params.items():
            ref = a.get(f"grad.{key}.{n}")
            if ref is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, f"{key}.{n}"
                continue
            err = float((p.grad - ref).norm() / (ref.norm() + 1e-12))
            assert err <= 1e-4, f"{key}.{n}: {err:.3e}"
    ref = a["grad.emb.t"]
    assert float((enc.t_table.grad - ref).norm() / ref.norm()) <= 1e-4


def test_interpolate_matches_reference():
    """rendering.py:365-460 run by the reference itself with the splat op swapped for the oracle's restatement."""
    meta, a = golden_io.load("interpolate_w64")
    rt = {k: a[f"t.{k}"] for k in meta["keys"]}
    rp = {k: a[f"tp1.{k}"] for k in meta["keys"]}
    for dt in meta["dts"]:
        rgb, depth = O.interpolate(rt, rp, dt, a["K"], a["c2w"], tuple(meta["img_wh"]))
        _close(f"rgb@{dt}", rgb, a[f"out.rgb.{dt}"])
        _close(f"depth@{dt}", depth, a[f"out.depth.{dt}"])


def test_interpolate_and_splat_match_the_reference_cuda_kernel():
    """tests/golden/interpolate_cuda_w64.npz: the UNMODIFIED reference interpolate() run on a B200 including its own CUDA
    splat kernel (softsplat.py:6-44, NVRTC-compiled by tests/golden/make_splat_golden.py) - pins the oracle's restatement
    of the splat against the real thing.  4e-6: the kernel's atomicAdd order is not deterministic."""
    meta, a = golden_io.load("interpolate_cuda_w64")
    assert "CUDA splat kernel" in meta["what"]
    rt = {k: a[f"t.{k}"] for k in meta["keys"]}
    rp = {k: a[f"tp1.{k}"] for k in meta["keys"]}
    for dt in meta["dts"]:
        rgb, depth = O.interpolate(rt, rp, dt, a["K"], a["c2w"], tuple(meta["img_wh"]))
        _close(f"rgb@{dt}", rgb, a[f"out.rgb.{dt}"], tol=4e-6)
        _close(f"depth@{dt}", depth, a[f"out.depth.{dt}"], tol=4e-6)
    _close("splat", O.splat_average(a["splat.in"], a["splat.flow"]), a["splat.out"], tol=4e-6)


def test_splat_average_known_answers():
    """softsplat.py:6-44 semantics on hand-checkable cases: integer flow moves pixels, half-pixel flow splits 50/50,
    out-of-frame contributions are dropped, untouched pixels stay 0 (weight 0 -> divide by 1)."""
    inp = torch.arange(1.0, 7.0).reshape(1, 1, 2, 3)
    flow = torch.zeros(1, 2, 2, 3)
    assert torch.equal(O.splat_average(inp, flow), inp)
    flow[:, 0] = 1.0                                           # everything one pixel to the right
    out = O.splat_average(inp, flow)
    assert torch.equal(out[0, 0], torch.tensor([[0.0, 1.0, 2.0], [0.0, 4.0, 5.0]]))
    flow[:, 0] = 0.5                                           # split between x and x+1: average of the two sources
    out = O.splat_average(inp, flow)
    assert torch.allclose(out[0, 0], torch.tensor([[1.0, 1.5, 2.5], [4.0, 4.5, 5.5]]))


def test_full_width_seeded_case():
    """W=256 fixture: weights are rebuilt from the seed (checksum-guarded), outputs are the reference's."""
    from nsff_pl_b200.models.nerf import NeRF
    meta, a = golden_io.load("train_warp_w256_seeded")
    torch.manual_seed(meta["build_seed"])
    fine = NeRF("fine", D=8, W=256, skips=[4], in_channels_xyz=63, in_channels_dir=27, use_viewdir=True,
                encode_appearance=False, in_channels_a=48, encode_transient=True, in_channels_t=48,
                output_flow=True, flow_scale=0.2)
    t_table = torch.nn.Embedding(meta["n_frames"], 48).weight.detach()
    energise_(fine)
    csum = float(sum(v.double().abs().sum() for v in fine.state_dict().values()))
    if abs(csum - meta["weight_abs_sum"]) > 1e-6 * meta["weight_abs_sum"] or \
            abs(float(t_table.double().abs().sum()) - meta["emb_abs_sum"]) > 1e-6 * meta["emb_abs_sum"]:
        pytest.skip("torch RNG / default init differs from the build container; seeded fixture not comparable")
    kw = dict(meta["call"])
    max_t = kw.pop("max_t")
    enc = O.Encoders(O.pos_freqs(9, 10), O.pos_freqs(3, 4), t_table=t_table)
    got = O.render_rays({"fine": O.FieldSpec.from_module(fine)}, enc, a["rays"], a["ts"], max_t, **kw)
    assert list(got.keys()) == meta["result_keys"]
    for k, v in golden_io.expected(a).items():
        _close(k, got[k], v)


def test_frame_rays_matches_reference():
    """ray_utils.py get_ray_directions -> get_rays -> get_ndc_rays, composed as monocular.py:268-276 does."""
    meta, a = golden_io.load("frame_rays")
    w, h = meta["img_wh"]
    for i in range(meta["n"]):
        got = O.frame_rays(h, w, torch.as_tensor(a["K"]), torch.as_tensor(a[f"c2w.{i}"]), near=meta["near"])
        assert torch.equal(got, torch.as_tensor(a[f"rays.{i}"])), i
