"""GPU: nsff_pl_b200.losses.NeRFWLoss (csrc/nsff_loss.cu) against the UNMODIFIED reference
losses.NeRFWLoss (losses.py:31-171, staged in baseline/_ref) run on the CPU: every term of the dict and the gradient of
their sum w.r.t. every differentiable input.  fp32 on both sides: 2e-5 relative on the terms, 2e-4 relative L2 on the
gradients (sums are accumulated in a different order)."""
import pytest
import torch

from oracle import ref_loader

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ref_loader.available() or ref_loader.stage()), reason="reference not staged")]

DIFF = ("rgb_fine", "depth_fine", "rgb_coarse", "depth_coarse", "transient_weights_fine", "static_weights_fine", "xyz_fw",
        "xyz_bw", "rgb_fw", "rgb_bw", "xyzs_fw_bw", "xyzs_bw_fw", "xyzs_fw", "xyzs_bw")


def _case(R, S, n_frames, seed, coarse):
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.rand(*s, generator=g)
    xyzs = torch.cat([rnd(R, S, 2) * 2 - 1, -rnd(R, S, 1) * 0.9], -1)             # NDC points, z in (-0.9, 0)
    inp = {"rgb_fine": rnd(R, 3), "depth_fine": rnd(R), "transient_weights_fine": rnd(R, S) * 0.2,
           "static_weights_fine": rnd(R, S) * 0.2, "xyz_fw": torch.cat([rnd(R, 2) * 2 - 1, -rnd(R, 1) * 0.9], -1),
           "xyz_bw": torch.cat([rnd(R, 2) * 2 - 1, -rnd(R, 1) * 0.9], -1), "disocc_fw": rnd(R, 1), "disocc_bw": rnd(R, 1),
           "rgb_fw": rnd(R, 3), "rgb_bw": rnd(R, 3), "disoccs_fw": rnd(R, S, 1), "disoccs_bw": rnd(R, S, 1),
           "xyzs_fw_bw": xyzs + 0.05 * torch.randn(R, S, 3, generator=g), "xyzs_bw_fw": xyzs + 0.05 * torch.randn(R, S, 3, generator=g),
           "xyzs_fine": xyzs, "xyzs_fw": xyzs + 0.03 * torch.randn(R, S, 3, generator=g),
           "xyzs_bw": xyzs + 0.03 * torch.randn(R, S, 3, generator=g)}
    if coarse:
        inp["rgb_coarse"], inp["depth_coarse"] = rnd(R, 3), rnd(R)
    ts = torch.randint(0, n_frames, (R,), generator=g)
    ts[0], ts[1] = 0, n_frames - 1                                                    # no bw / no fw neighbour
    tgt = {"rgbs": rnd(R, 3), "disps": rnd(R) + 0.1, "cam_ids": torch.zeros(R, dtype=torch.long), "ts": ts,
           "uv_fw": rnd(R, 2) * 60, "uv_bw": rnd(R, 2) * 60}
    return inp, tgt


@pytest.mark.parametrize("R,S,thickness,coarse", [(37, 23, 1, False), (64, 20, 3, True), (5, 128, 4, False)])
def test_fused_loss_matches_the_reference_loss(R, S, thickness, coarse):
    import ref_harness as H
    from nsff_pl_b200.losses import NeRFWLoss, TERM_KEYS
    ref = ref_loader.load()
    n_frames = 8
    inp, tgt = _case(R, S, n_frames, seed=R + S, coarse=coarse)
    kwargs = dict(epoch=4, output_transient=True, output_transient_flow=["fw", "bw", "disocc"])
    rloss = H.make_loss(ref, n_frames, img_wh=(64, 36), thickness=thickness)
    leaves_r = {k: v.clone().requires_grad_(k in DIFF) for k, v in inp.items()}
    rterms = rloss(leaves_r, tgt, **kwargs)
    sum(rterms.values()).backward()

    ours = NeRFWLoss(thickness=thickness).cuda()
    ours.register_buffer("Ks", rloss.Ks.cuda())
    ours.register_buffer("Ps", rloss.Ps.cuda())
    ours.max_t = rloss.max_t
    leaves = {k: v.clone().cuda().requires_grad_(k in DIFF) for k, v in inp.items()}
    terms = ours(leaves, {k: v.cuda() for k, v in tgt.items()}, **kwargs)
    sum(terms.values()).backward()
    torch.cuda.synchronize()

    assert list(terms) == [k for k in TERM_KEYS if k in rterms or k.startswith("flow_")]
    for k, v in rterms.items():
        assert abs(float(terms[k]) - float(v)) <= 2e-5 * max(1.0, abs(float(v))), (k, float(terms[k]), float(v))
    for k in DIFF:
        if k not in inp:
            continue
        r, gt = leaves_r[k].grad, leaves[k].grad
        assert (r is None) == (gt is None), k
        if r is None:
            continue
        err = float((gt.cpu().double() - r.double()).norm() / (r.double().norm() + 1e-20))
        assert err <= 2e-4, f"{k}: relative gradient error {err:.3e}"


def test_fused_loss_without_flows_and_with_ray_weights():
    import ref_harness as H
    from nsff_pl_b200.losses import NeRFWLoss
    ref = ref_loader.load()
    inp, tgt = _case(48, 8, 8, seed=3, coarse=True)
    w = torch.rand(48, generator=torch.Generator().manual_seed(9)) + 0.5
    kw = dict(epoch=0, output_transient_flow=[])
    rl = H.make_loss(ref, 8)
    lr = {k: v.clone().requires_grad_(k in DIFF) for k, v in inp.items()}
    rt = rl(lr, tgt, weights=w, **kw)
    sum(rt.values()).backward()
    lo = {k: v.clone().cuda().requires_grad_(k in DIFF) for k, v in inp.items()}
    ot = NeRFWLoss().cuda()(lo, {k: v.cuda() for k, v in tgt.items()}, weights=w.cuda(), **kw)
    sum(ot.values()).backward()
    assert list(ot) == list(rt) == ["col_l", "disp_l"]
    for k in rt:
        assert abs(float(ot[k]) - float(rt[k])) <= 2e-5 * max(1.0, abs(float(rt[k]))), k
    for k in ("rgb_fine", "depth_fine", "rgb_coarse", "depth_coarse"):
        r, g = lr[k].grad.double(), lo[k].grad.cpu().double()
        assert float((g - r).norm() / (r.norm() + 1e-20)) <= 2e-4, k
    with pytest.raises(NotImplementedError):
        NeRFWLoss(topk=0.5).cuda()(lo, {k: v.cuda() for k, v in tgt.items()}, **kw)
