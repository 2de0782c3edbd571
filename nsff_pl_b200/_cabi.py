"""ctypes binding of include/nsff_b200.h (libnsff_b200.so) + tensor-checking call wrappers.

Everything here is plumbing: it validates torch tensors the way the reference's own native ops do
(models/softsplat.py:217-222 asserts contiguity / CUDA-ness; alt_cuda_corr/correlation.cpp:18-24
CHECK_CUDA / CHECK_CONTIGUOUS) and hands raw device pointers + the current CUDA stream to the C ABI.
There is no CPU implementation behind any of these calls: a missing library or a CPU tensor raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# NSFF_LIB: an alternative build of the same library (e.g. the -DNSFF_TIMELINE build of tools/timeline.py)
LIB_PATH = os.environ.get("NSFF_LIB") or os.path.join(_HERE, "libnsff_b200.so")

OUT_STATIC, OUT_TRANSIENT, OUT_FLOW_FW, OUT_FLOW_BW, OUT_SIGMA_ONLY, OUT_EVERY_LAYER = 1, 2, 4, 8, 16, 32

_lib = None


class NsffError(RuntimeError):
    pass


class FieldDesc(C.Structure):
    _fields_ = [("W", C.c_int32), ("D", C.c_int32), ("skip", C.c_int32), ("in_xyz", C.c_int32),
                ("in_dir", C.c_int32), ("in_a", C.c_int32), ("in_t", C.c_int32), ("use_viewdir", C.c_int32),
                ("has_transient", C.c_int32), ("has_flow", C.c_int32), ("flow_scale", C.c_float)]


class FieldParams(C.Structure):
    _fields_ = [("static_w", C.c_void_p * 8), ("static_b", C.c_void_p * 8),
                ("static_final_w", C.c_void_p), ("static_final_b", C.c_void_p),
                ("static_dir_w", C.c_void_p), ("static_dir_b", C.c_void_p),
                ("static_sigma_w", C.c_void_p), ("static_sigma_b", C.c_void_p),
                ("static_rgb_w", C.c_void_p), ("static_rgb_b", C.c_void_p),
                ("trans_w", C.c_void_p * 8), ("trans_b", C.c_void_p * 8),
                ("trans_final_w", C.c_void_p), ("trans_final_b", C.c_void_p),
                ("trans_sigma_w", C.c_void_p), ("trans_sigma_b", C.c_void_p),
                ("trans_rgb_w", C.c_void_p), ("trans_rgb_b", C.c_void_p),
                ("trans_fw_w", C.c_void_p), ("trans_fw_b", C.c_void_p),
                ("trans_bw_w", C.c_void_p), ("trans_bw_b", C.c_void_p)]


class Points(C.Structure):
    _fields_ = [("xyz", C.c_void_p), ("offset", C.c_void_p), ("offset_ld", C.c_int32),
                ("xyz_out", C.c_void_p), ("dirs", C.c_void_p), ("t_table", C.c_void_p),
                ("t_index", C.c_void_p), ("t_shift", C.c_int32), ("max_t", C.c_int32),
                ("a_table", C.c_void_p), ("a_index", C.c_void_p), ("zs", C.c_void_p),
                ("B", C.c_int64), ("S", C.c_int32), ("xyz_fw_out", C.c_void_p), ("xyz_bw_out", C.c_void_p)]


class CompositeArgs(C.Structure):
    _fields_ = [("R", C.c_int64), ("S", C.c_int32),
                ("zs", C.c_void_p), ("xyz", C.c_void_p), ("raw", C.c_void_p), ("raw_ld", C.c_int32),
                ("has_static_rgb", C.c_int32), ("has_transient", C.c_int32), ("has_flow", C.c_int32),
                ("noise_static", C.c_void_p), ("noise_trans", C.c_void_p), ("noise_std", C.c_float),
                ("vis", C.c_void_p),
                ("raw_fw", C.c_void_p), ("raw_bw", C.c_void_p), ("raw_w_ld", C.c_int32),
                ("xyz_fw", C.c_void_p), ("xyz_bw", C.c_void_p),
                ("noise_fw", C.c_void_p), ("noise_bw", C.c_void_p),
                ("static_sigmas", C.c_void_p), ("transient_sigmas", C.c_void_p),
                ("static_weights", C.c_void_p), ("transient_weights", C.c_void_p), ("weights", C.c_void_p),
                ("static_alphas", C.c_void_p), ("transient_alphas", C.c_void_p),
                ("xyzs_fw_bw", C.c_void_p), ("xyzs_bw_fw", C.c_void_p),
                ("disoccs_fw", C.c_void_p), ("disoccs_bw", C.c_void_p),
                ("depth", C.c_void_p), ("rgb", C.c_void_p), ("transient_alpha", C.c_void_p),
                ("transient_rgb", C.c_void_p), ("static_only_rgb", C.c_void_p), ("static_only_depth", C.c_void_p),
                ("xyz_fine", C.c_void_p), ("flow_fw", C.c_void_p), ("xyz_fw_ray", C.c_void_p),
                ("flow_bw", C.c_void_p), ("xyz_bw_ray", C.c_void_p),
                ("rgb_fw", C.c_void_p), ("rgb_bw", C.c_void_p), ("disocc_fw", C.c_void_p), ("disocc_bw", C.c_void_p)]


class FieldGrads(C.Structure):        # same layout as FieldParams, but writable fp32 gradient buffers
    _fields_ = FieldParams._fields_


class CompositeGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("static_sigmas", "transient_sigmas", "static_weights", "transient_weights", "weights",
                 "xyzs_fw_bw", "xyzs_bw_fw", "depth", "rgb", "transient_alpha", "transient_rgb",
                 "static_only_rgb", "static_only_depth", "xyz_fine", "flow_fw", "xyz_fw_ray", "flow_bw", "xyz_bw_ray",
                 "rgb_fw", "rgb_bw", "d_raw", "d_raw_fw", "d_raw_bw")]


class InterpolateArgs(C.Structure):
    _fields_ = [("R", C.c_int64), ("S", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("dt", C.c_float), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("P", C.c_float * 12),
                ("xyzs", C.c_void_p), ("zs", C.c_void_p), ("static_rgb", C.c_void_p), ("static_alpha", C.c_void_p),
                ("rgb_t", C.c_void_p), ("alpha_t", C.c_void_p), ("flow_fw", C.c_void_p), ("ld_t", C.c_int32),
                ("rgb_tp1", C.c_void_p), ("alpha_tp1", C.c_void_p), ("flow_bw", C.c_void_p), ("ld_tp1", C.c_int32)]


class LossArgs(C.Structure):
    _PTRS_IN = ("rgb_fine", "depth_fine", "rgb_coarse", "depth_coarse", "tw", "sw", "xyz_fw", "xyz_bw", "rgb_fw", "rgb_bw",
                "disocc_fw", "disocc_bw", "disoccs_fw", "disoccs_bw", "xyzs_fw_bw", "xyzs_bw_fw", "xyzs_fine", "xyzs_fw",
                "xyzs_bw", "rgbs", "disps", "uv_fw", "uv_bw", "cam_ids", "ts", "Ks", "Ps", "ray_weights", "medians",
                "stats", "term_sums", "terms", "g_terms")
    _PTRS_GRAD = ("d_rgb_fine", "d_depth_fine", "d_rgb_coarse", "d_depth_coarse", "d_tw", "d_sw", "d_xyz_fw", "d_xyz_bw",
                  "d_rgb_fw", "d_rgb_bw", "d_xyzs_fw_bw", "d_xyzs_bw_fw", "d_xyzs_fw", "d_xyzs_bw")
    _fields_ = ([("R", C.c_int64), ("S", C.c_int32), ("has_flow", C.c_int32), ("thickness", C.c_int32),
                 ("n_reg", C.c_int32), ("n_frames", C.c_int32), ("max_t", C.c_int32),
                 ("lambda_geo_d", C.c_float), ("lambda_geo_f", C.c_float), ("lambda_reg", C.c_float),
                 ("lambda_ent", C.c_float), ("cross_entropy_w", C.c_float)]
                + [(n, C.c_void_p) for n in _PTRS_IN] + [(n, C.c_void_p) for n in _PTRS_GRAD])


class ImgGemmArgs(C.Structure):
    _fields_ = [("a_img", C.c_void_p), ("a_chunks", C.c_int32), ("b_blocks", C.c_void_p), ("N", C.c_int32),
                ("bias", C.c_void_p), ("epi", C.c_int32), ("out_img", C.c_void_p), ("mask_img", C.c_void_p),
                ("out_rows", C.c_void_p), ("ld", C.c_int32), ("ncols", C.c_int32), ("beta", C.c_int32),
                ("n_rows", C.c_int64), ("relu_bits", C.c_void_p), ("mask_bits", C.c_void_p), ("colsum", C.c_void_p)]


class ImgWgradArgs(C.Structure):
    _fields_ = [("a_img", C.c_void_p), ("b_img", C.c_void_p), ("b_chunks", C.c_int32), ("n_rows", C.c_int64),
                ("acc_t", C.c_void_p)]


# name -> (restype, argtypes); also the list tests/test_cabi_symbols.py checks against the header
PROTOTYPES = {
    "nsff_abi_version": (C.c_int, []),
    "nsff_error_string": (C.c_char_p, [C.c_int]),
    "nsff_last_cuda_error": (C.c_char_p, []),
    "nsff_set_device": (C.c_int, [C.c_int]),
    "nsff_launch_count": (C.c_uint64, []),
    "nsff_pos_embedding": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "nsff_sample_points": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "nsff_points_from_depths": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "nsff_field_pack_bytes": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(C.c_size_t)]),
    "nsff_field_pack": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(FieldParams), C.c_void_p, C.c_size_t, C.c_void_p]),
    "nsff_field_forward": (C.c_int, [C.POINTER(FieldDesc), C.c_void_p, C.POINTER(Points), C.c_int, C.c_void_p,
                                     C.c_int32, C.c_void_p]),
    "nsff_field_forward_embedded": (C.c_int, [C.POINTER(FieldDesc), C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                              C.c_int, C.c_void_p, C.c_int32, C.c_void_p]),
    "nsff_field_fp32_scratch_floats": (C.c_size_t, [C.POINTER(FieldDesc), C.c_int64]),
    "nsff_field_forward_fp32": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(FieldParams), C.POINTER(Points),
                                          C.c_void_p, C.c_int32, C.c_int64, C.c_int, C.c_void_p, C.c_int32,
                                          C.c_void_p, C.c_void_p]),
    "nsff_field_bwd_scratch_floats": (C.c_size_t, [C.POINTER(FieldDesc), C.c_int64]),
    "nsff_field_backward_fp32": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(FieldParams), C.POINTER(Points), C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(FieldGrads), C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsff_field_bwd_tc_pack_bytes": (C.c_size_t, [C.POINTER(FieldDesc)]),
    "nsff_field_bwd_tc_pack": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(FieldParams), C.c_void_p, C.c_void_p]),
    "nsff_field_save_bytes": (C.c_size_t, [C.POINTER(FieldDesc), C.c_int64, C.c_int]),
    "nsff_field_forward_save": (C.c_int, [C.POINTER(FieldDesc), C.c_void_p, C.POINTER(Points), C.c_int, C.c_void_p,
                                          C.c_int32, C.c_void_p, C.c_void_p]),
    "nsff_field_bwd_tc_scratch_bytes": (C.c_size_t, [C.POINTER(FieldDesc), C.c_int64, C.c_int]),
    "nsff_field_backward_tc": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(FieldParams), C.c_void_p, C.POINTER(Points),
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(FieldGrads), C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsff_img_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "nsff_img_from_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32,
                                     C.c_int32, C.c_void_p]),
    "nsff_img_to_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int64, C.c_void_p]),
    "nsff_img_pack_weight": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p]),
    "nsff_img_gemm": (C.c_int, [C.POINTER(ImgGemmArgs), C.c_void_p]),
    "nsff_img_wgrad": (C.c_int, [C.POINTER(ImgWgradArgs), C.c_void_p]),
    "nsff_composite_backward": (C.c_int, [C.POINTER(CompositeArgs), C.POINTER(CompositeGrads), C.c_void_p]),
    "nsff_interpolate_scratch_floats": (C.c_size_t, [C.c_int64, C.c_int]),
    "nsff_interpolate": (C.c_int, [C.POINTER(InterpolateArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsff_loss_stats_floats": (C.c_int, []),
    "nsff_nerfw_loss_forward": (C.c_int, [C.POINTER(LossArgs), C.c_void_p]),
    "nsff_nerfw_loss_backward": (C.c_int, [C.POINTER(LossArgs), C.c_void_p]),
    "nsff_frame_rays": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p]),
    "nsff_ndc_rays": (C.c_int, [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsff_visibility": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.c_void_p, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p, C.c_void_p]),
    "nsff_composite": (C.c_int, [C.POINTER(CompositeArgs), C.c_void_p]),
    "nsff_sample_pdf": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                  C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "nsff_merge_sorted": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64,
                                    C.c_void_p, C.c_void_p]),
}


def lib():
    """Load libnsff_b200.so (once).  Fails loudly: there is no other implementation to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NsffError(
                f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built.  Run "
                "`python -m nsff_pl_b200.build` (or __graft_entry__.build()); nsff_pl_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if handle.nsff_abi_version() != 1:
            raise NsffError("libnsff_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        L = lib()
        msg = L.nsff_error_string(rc).decode()
        if rc == -3:
            msg += ": " + L.nsff_last_cuda_error().decode()
        raise NsffError(f"{what} failed: {msg} (code {rc})")


def dev_tensor(t, name, dtype=torch.float32):
    """CHECK_CUDA + CHECK_CONTIGUOUS + dtype; returns the tensor (never copies)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise NotImplementedError(f"{name} is a CPU tensor: nsff_pl_b200 kernels are CUDA-only (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def select_device(t):
    check(lib().nsff_set_device(t.device.index if t.device.index is not None else torch.cuda.current_device()),
          "nsff_set_device")


def launch_count():
    return int(lib().nsff_launch_count())
