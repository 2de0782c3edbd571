#!/usr/bin/env python
"""Generate tests/golden/interpolate_cuda_w64.npz by running the UNMODIFIED reference interpolate()
(models/rendering.py:365-460) INCLUDING its own CUDA splat kernel (models/softsplat.py:6-44, specialised by the
reference's cupy_kernel(), :148-196).  Needs a GPU and the staged reference (baseline/_ref, oracle/ref_loader.py):

    gpurun -- 'python tests/golden/make_splat_golden.py gpurun_out/interpolate_cuda_w64.npz'
    cp gpurun_out/interpolate_cuda_w64.npz tests/golden/

cupy is not installed: oracle/ref_loader.py's `cupy` shim compiles the kernel string with NVRTC and launches it through
the driver API (cuda-python), which is all the reference asks of cupy (softsplat.py:200-202, 228-236).  Nothing of the
reference is restated here: render_rays, the projection, the splat, the normalisation and the MPI composite are all
executed by the reference's own code.  Also stored: one raw FunctionSoftsplat('average') call on random inputs.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, ".."), os.path.join(HERE, "..", "..")]
from oracle import ref_loader                      # noqa: E402
from parity_util import build_fields, synthetic_rays   # noqa: E402
from scenes import FakeDataset                     # noqa: E402


def main(out_path):
    assert torch.cuda.is_available(), "needs a GPU (the reference's splat op is CUDA-only)"
    ref = ref_loader.load()
    N_FRAMES = 6
    models, emb, _, _ = build_fields(W=64, seed=31, n_frames=N_FRAMES, device=None)
    rmodels, remb = ref_loader.build_reference_fields(models, emb)
    w, h = 12, 7
    rays, _ = synthetic_rays(w * h, n_frames=N_FRAMES, seed=14, img_wh=(w, h), frame=2)
    ts = torch.full((w * h,), 2, dtype=torch.long)
    kw = dict(N_samples=10, N_importance=0, test_time=True, output_transient_flow=["fw", "bw"])
    with torch.no_grad():
        rt = ref.rendering.render_rays(rmodels, remb, rays, ts, N_FRAMES - 1, **kw)
        rp = ref.rendering.render_rays(rmodels, remb, rays, ts + 1, N_FRAMES - 1, **kw)
    ds = FakeDataset(n_frames=N_FRAMES, n_cams=1, img_wh=(w, h), seed=8)
    K, c2w = ds.Ks[0], torch.tensor(ds.poses[1])
    arrays, how = {}, "as-is"
    real_splat = ref.rendering.FunctionSoftsplat
    try:
        for dt in (0.0, 0.3, 0.75):
            try:
                rgb, depth = ref.rendering.interpolate(rt, rp, dt, K, c2w, (w, h))
            except AssertionError:
                # the op asserts contiguous inputs (softsplat.py:217-222) and current torch keeps the permuted strides of
                # `of_fw[s:s+1]` across .cuda(): hand the SAME op contiguous copies of the same values
                how = "inputs made contiguous before the reference's FunctionSoftsplat (it asserts contiguity)"
                ref.rendering.FunctionSoftsplat = lambda tenInput, tenFlow, tenMetric, strType: real_splat(
                    tenInput.contiguous(), tenFlow.contiguous(), tenMetric, strType)
                rgb, depth = ref.rendering.interpolate(rt, rp, dt, K, c2w, (w, h))
            arrays[f"out.rgb.{dt}"], arrays[f"out.depth.{dt}"] = rgb.cpu().numpy(), depth.cpu().numpy()
    finally:
        ref.rendering.FunctionSoftsplat = real_splat
    keys = ("xyzs_fine", "zs_fine", "static_rgbs_fine", "static_alphas_fine", "transient_rgbs_fine",
            "transient_alphas_fine", "transient_flows_fw", "transient_flows_bw")
    for k in keys:
        arrays[f"t.{k}"], arrays[f"tp1.{k}"] = rt[k].numpy().copy(), rp[k].numpy().copy()
    arrays["K"], arrays["c2w"] = K.numpy(), c2w.numpy()
    g = torch.Generator().manual_seed(5)
    s_in = torch.rand(1, 4, 9, 13, generator=g)
    s_flow = 3.0 * torch.randn(1, 2, 9, 13, generator=g)          # some targets leave the frame
    s_out = real_splat(s_in.cuda(), s_flow.cuda(), None, "average").cpu()
    arrays["splat.in"], arrays["splat.flow"], arrays["splat.out"] = s_in.numpy(), s_flow.numpy(), s_out.numpy()
    meta = dict(img_wh=[w, h], dts=[0.0, 0.3, 0.75], keys=list(keys), splat_inputs=how,
                device=torch.cuda.get_device_name(0), torch=torch.__version__,
                what="unmodified reference interpolate() incl. its CUDA splat kernel (NVRTC-compiled)")
    arrays = {k: np.ascontiguousarray(v) for k, v in arrays.items()}
    arrays["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(out_path, **arrays)
    print("wrote", out_path, meta)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "interpolate_cuda_w64.npz"))
