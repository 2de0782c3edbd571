"""CPU: the staged UNMODIFIED reference (baseline/_ref, oracle/ref_loader.py), the caller harness of tests/ref_harness.py
and the reference's NeRFWLoss (losses.py:31-171) with the kornia.filter2d shim.  Skipped when the reference has not been
staged (it is staged by __graft_entry__.build() in the build container and travels with the gpurun snapshot)."""
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not (ref_loader.available() or ref_loader.stage()), reason="reference not staged")

TRAIN_KW = dict(output_transient=True, output_transient_flow=["fw", "bw", "disocc"])


def _fields(n_frames=6, coarse=False, seed=0):
    from parity_util import build_fields
    models, emb, specs, enc = build_fields(W=64, coarse=coarse, seed=seed, n_frames=n_frames, device=None)
    rmodels, remb = ref_loader.build_reference_fields(models, emb)
    return models, emb, specs, enc, rmodels, remb


def test_staged_reference_matches_the_oracle_bit_for_bit():
    from oracle import nsff_oracle as O
    from parity_util import compare_dicts, synthetic_rays
    ref = ref_loader.load()
    _, _, specs, enc, rmodels, remb = _fields()
    rays, ts = synthetic_rays(24, n_frames=6, seed=3)
    kw = dict(N_samples=16, perturb=0, noise_std=0, N_importance=0, test_time=False, output_transient_flow=["fw", "bw", "disocc"])
    with torch.no_grad():
        want = ref.rendering.render_rays(rmodels, remb, rays, ts, 5, **kw)
        got = O.render_rays(specs, enc, rays, ts, 5, **kw)
    errs = compare_dicts(got, want, {}, default_tol=1e-6)
    assert max(errs.values()) <= 1e-6


def test_filter2d_shim_is_a_zero_padded_box_sum():
    ref = ref_loader.load()
    import sys
    f2d = ref.losses.filter2d
    x = torch.rand(1, 1, 5, 9)
    assert torch.equal(f2d(x, torch.ones(1, 1, 1), "constant"), x)                    # thickness 1: identity
    y = f2d(x, torch.ones(1, 1, 3), "constant")
    pad = torch.nn.functional.pad(x, (1, 1))
    assert torch.allclose(y, pad[..., :-2] + pad[..., 1:-1] + pad[..., 2:])
    y4 = f2d(x, torch.ones(1, 1, 4), "constant")                                       # even window: 1 in front, 2 behind
    pad = torch.nn.functional.pad(x, (1, 2))
    assert torch.allclose(y4, sum(pad[..., i:i + 9] for i in range(4)))
    assert "kornia" not in sys.modules or not hasattr(sys.modules["kornia"], "__nsff_shim__")   # shims do not leak


def test_system_forward_chunking_equals_one_call_and_loss_backpropagates():
    import ref_harness as H
    ref = ref_loader.load()
    _, _, _, _, rmodels, remb = _fields()
    from parity_util import synthetic_rays
    rays, ts = synthetic_rays(20, n_frames=6, seed=5)
    for m in rmodels.values():
        m.requires_grad_(True)
    remb["t"].requires_grad_(True)
    whole = H.system_forward(ref.rendering.render_rays, rmodels, remb, rays, ts, 6, 12, 0, 0, 0, 64, **TRAIN_KW)
    chunks = H.system_forward(ref.rendering.render_rays, rmodels, remb, rays, ts, 6, 12, 0, 0, 0, 7, **TRAIN_KW)   # ragged
    assert list(whole) == list(chunks)
    for k in whole:
        assert torch.allclose(whole[k], chunks[k], atol=2e-4, rtol=1e-4), k      # (MKL blocks GEMMs differently per batch size)
    loss = H.make_loss(ref, 6, img_wh=(32, 18))
    batch = H.make_batch(rays, ts, img_wh=(32, 18))
    terms = loss(chunks, batch, epoch=3, **TRAIN_KW)
    assert {"col_l", "disp_l", "entropy_l", "cross_entropy_l", "pho_l", "cyc_l", "reg_temp_sm_l", "reg_min_l",
            "reg_sp_sm_l"} <= set(terms)
    total = sum(terms.values())
    assert torch.isfinite(total)
    total.backward()
    g = rmodels["fine"].transient_flow_fw[0].weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0


def test_eval_f_moves_every_value_to_the_host_and_concatenates_ragged_chunks():
    import ref_harness as H
    ref = ref_loader.load()
    _, _, _, _, rmodels, remb = _fields(coarse=True)
    from parity_util import synthetic_rays
    rays, ts = synthetic_rays(13, n_frames=6, seed=6)
    out = H.eval_f(ref.rendering.render_rays, rmodels, remb, rays, ts, 5, 8, 4, 5, output_transient_flow=["fw", "bw"])
    one = H.eval_f(ref.rendering.render_rays, rmodels, remb, rays, ts, 5, 8, 4, 64, output_transient_flow=["fw", "bw"])
    assert all(v.device.type == "cpu" and v.shape[0] == 13 for v in out.values())
    for k in one:
        assert torch.allclose(out[k], one[k], atol=2e-3, rtol=2e-3), k      # (resampled depths move with the GEMM blocking)
