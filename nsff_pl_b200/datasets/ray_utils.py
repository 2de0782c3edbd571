"""On-device ray generation: the reference builds every test frame's rays on the CPU with
kornia/torch (datasets/ray_utils.py:7-106, datasets/monocular.py:268-276) and ships 3.5 MB per frame over PCIe; here
one kernel writes the (H*W, 6) NDC rays straight into HBM.  Same function names / argument meaning as the reference."""
import ctypes as C

import torch

from .. import _cabi

__all__ = ["get_ndc_rays", "frame_rays"]


def _K4(K):
    K = torch.as_tensor(K, dtype=torch.float32).cpu()
    return (C.c_float * 4)(float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]))


def get_ndc_rays(K, near, shift_near, rays_o, rays_d):
    """reference ray_utils.py:62-106 (scalar near / shift_near).  rays_o, rays_d: (N,3) CUDA fp32."""
    rays_o = _cabi.dev_tensor(rays_o.contiguous(), "rays_o")
    rays_d = _cabi.dev_tensor(rays_d.contiguous(), "rays_d")
    out_o, out_d = torch.empty_like(rays_o), torch.empty_like(rays_d)
    _cabi.select_device(rays_o)
    _cabi.check(_cabi.lib().nsff_ndc_rays(_K4(K), float(near), float(shift_near), _cabi.ptr(rays_o), _cabi.ptr(rays_d),
                                          rays_o.shape[0], _cabi.ptr(out_o), _cabi.ptr(out_d),
                                          _cabi.stream_ptr(rays_o.device)), "nsff_ndc_rays")
    return out_o, out_d


def frame_rays(H, W, K, c2w, near=1.0, device="cuda"):
    """(H*W, 6) NDC rays of a frame: get_ray_directions(H, W, K) -> get_rays(directions, c2w) ->
    get_ndc_rays(K, near, shift_near, ...) with shift_near = -min(-1, c2w[2,3])  (monocular.py:268-272)."""
    c2w = torch.as_tensor(c2w, dtype=torch.float32).cpu().reshape(3, 4)
    shift_near = -min(-1.0, float(c2w[2, 3]))
    dev = torch.device(device)
    rays = torch.empty(H * W, 6, device=dev, dtype=torch.float32)
    c2w12 = (C.c_float * 12)(*[float(v) for v in c2w.reshape(-1)])
    _cabi.select_device(rays)
    _cabi.check(_cabi.lib().nsff_frame_rays(_K4(K), c2w12, float(near), shift_near, int(H), int(W), _cabi.ptr(rays),
                                            _cabi.stream_ptr(dev)), "nsff_frame_rays")
    return rays
