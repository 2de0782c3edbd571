"""CPU: host-side logic of nsff_pl_b200 that needs no GPU - module surface, parameter names, head masks,
ray sharding - plus the structural rule that the product never imports the oracle."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_imports_the_oracle_or_reads_the_reference():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nsff_pl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "/root/reference" in txt or "nsff_oracle" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product code must not touch oracle/ or the reference tree: {bad}"


def test_state_dict_names_and_shapes_match_the_reference_checkpoint_surface():
    from nsff_pl_b200.models.nerf import NeRF
    fine = NeRF("fine", encode_transient=True, in_channels_t=48, output_flow=True)
    sd = fine.state_dict()
    assert sum(v.numel() for v in sd.values()) == 1_218_574          # SURVEY.md section 2.1 [probe]
    expect = {"static_xyz_encoding_1.0.weight": (256, 63), "static_xyz_encoding_5.0.weight": (256, 319),
              "static_xyz_encoding_final.weight": (256, 256), "static_dir_encoding.0.weight": (256, 283),
              "static_sigma.weight": (1, 256), "static_rgb.0.weight": (3, 256),
              "transient_xyz_encoding_1.0.weight": (256, 111), "transient_xyz_encoding_5.0.weight": (256, 367),
              "transient_xyz_encoding_final.bias": (256,), "transient_sigma.bias": (1,),
              "transient_rgb.0.weight": (3, 256), "transient_flow_fw.0.weight": (3, 256),
              "transient_flow_bw.0.bias": (3,)}
    for k, shp in expect.items():
        assert tuple(sd[k].shape) == shp, k
    coarse = NeRF("coarse", encode_transient=True, in_channels_t=48, output_flow=True)
    assert sum(v.numel() for v in coarse.state_dict().values()) == 1_217_032
    assert not coarse.has_flow_heads                                             # nerf.py:110: fine only
    assert NeRF("coarse", encode_appearance=True).encode_appearance is False     # nerf.py:67


def test_head_masks_and_channel_counts_follow_the_forward_modes():
    from nsff_pl_b200.models.nerf import NeRF
    from nsff_pl_b200 import _cabi
    m = NeRF.out_mask
    assert NeRF.out_channels(m(False, True, True, ["fw", "bw", "disocc"])) == 14
    assert NeRF.out_channels(m(False, True, True, [])) == 8
    assert NeRF.out_channels(m(False, True, False, [])) == 4
    assert NeRF.out_channels(m(False, False, True, ["bw"])) == 7
    assert NeRF.out_channels(m(True, True, True, [])) == 2
    assert NeRF.out_channels(m(True, True, False, [])) == 1
    assert m(True, True, True, ["fw"]) & _cabi.OUT_FLOW_FW == 0           # sigma_only ignores flow heads
    fine = NeRF("fine", encode_transient=True, in_channels_t=48, output_flow=True)
    assert 2 * (fine.macs_per_point(m(False, True, True, ["fw", "bw"])) +
                2 * fine.macs_per_point(m(False, False, True, ["fw"]))) == 4_759_040


def test_cpu_tensors_and_missing_backward_raise():
    from nsff_pl_b200.models.nerf import NeRF, PosEmbedding
    from nsff_pl_b200.models.rendering import render_rays, sample_pdf
    with pytest.raises(NotImplementedError):
        PosEmbedding(9, 10)(torch.zeros(4, 3))
    with pytest.raises(NotImplementedError):
        sample_pdf(torch.rand(2, 5), torch.rand(2, 4), 3)
    net = NeRF("fine")
    emb = {"xyz": PosEmbedding(9, 10), "dir": PosEmbedding(3, 4)}
    with pytest.raises(NotImplementedError):
        render_rays({"fine": net}, emb, torch.zeros(2, 6), None, 0, N_samples=4)
    with pytest.raises(NotImplementedError):          # NeRF.forward on an embedded row records no graph
        net(torch.zeros(2, 90))


def test_pos_embedding_attributes():
    from nsff_pl_b200.models.nerf import PosEmbedding
    e = PosEmbedding(9, 10)
    assert torch.equal(e.freqs, 2.0 ** torch.arange(10)) and e.is_octaves(10)
    assert not PosEmbedding(3, 4, logscale=False).is_octaves(4)
    assert "freqs" not in e.state_dict()              # plain attribute like the reference (nerf.py:12-15)


def test_install_as_reference_modules():
    import sys
    import nsff_pl_b200
    saved = {k: sys.modules.get(k) for k in ("models", "models.nerf", "models.rendering")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        nsff_pl_b200.install_as_reference_modules()
        from models.rendering import render_rays, sample_pdf      # noqa: F401  (what train.py/eval.py import)
        from models.nerf import PosEmbedding, NeRF                # noqa: F401
        assert render_rays.__module__ == "nsff_pl_b200.models.rendering"
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_shard_bounds_cover_the_frame_in_order():
    from nsff_pl_b200.parallel import shard_bounds
    for n, world in ((147456, 8), (10, 3), (7, 8), (0, 2)):
        blocks = [shard_bounds(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_pack_unpack_pixels_roundtrip():
    from nsff_pl_b200.parallel import PIXEL_KEYS, pack_pixels, unpack_pixels
    res = {"rgb_fine": torch.rand(5, 3), "depth_fine": torch.rand(5), "transient_flow_fw": torch.rand(5, 3),
           "transient_flow_bw": torch.rand(5, 3), "weights_fine": torch.rand(5, 8)}
    buf = pack_pixels(res)
    assert buf.shape == (5, 10)
    back = unpack_pixels(buf, res)
    for k in PIXEL_KEYS:
        assert torch.equal(back[k], res[k])


def test_backward_buffer_sizes_are_host_side_pure_functions():
    """Size queries of the tensor-core backward need no GPU: tile-image bytes, the activation save buffer written by
    nsff_field_forward_save, the W / W^T block pack and the scratch (recompute mode: the activation images; saved mode: the
    dZ image of every layer, which the chained backward keeps for the weight-gradient GEMMs)."""
    import ctypes as C
    from nsff_pl_b200 import _cabi
    from nsff_pl_b200.models.nerf import NeRF
    L = _cabi.lib()
    assert L.nsff_img_bytes(1, 1) == 128 * 128 and L.nsff_img_bytes(129, 4) == 2 * 4 * 128 * 128
    assert L.nsff_img_bytes(0, 4) == 0
    fine = NeRF("fine", encode_transient=True, in_channels_t=48, output_flow=True)
    desc = fine.field_desc()
    B = 1000                                                   # 8 tiles
    img, bits = 8 * 4 * 128 * 128, 8 * 8 * 128 * 4
    both = _cabi.OUT_STATIC | _cabi.OUT_TRANSIENT | _cabi.OUT_FLOW_FW | _cabi.OUT_FLOW_BW
    # static: 8 trunk layers + final + dir (bits for the 9 ReLU layers); transient: 8 trunk + final (8 bit planes)
    assert L.nsff_field_save_bytes(C.byref(desc), B, both) == (10 + 9) * img + (9 + 8) * bits
    assert L.nsff_field_save_bytes(C.byref(desc), B, _cabi.OUT_TRANSIENT | _cabi.OUT_FLOW_FW) == 9 * img + 8 * bits
    assert L.nsff_field_save_bytes(C.byref(desc), B, _cabi.OUT_STATIC | _cabi.OUT_SIGMA_ONLY) == 0
    pack = L.nsff_field_bwd_tc_pack_bytes(C.byref(desc))
    assert pack > 0 and pack % (64 * 128) == 0                 # whole 64-row x 64-column bf16 blocks
    full = L.nsff_field_bwd_tc_scratch_bytes(C.byref(desc), B, 0)
    lean = L.nsff_field_bwd_tc_scratch_bytes(C.byref(desc), B, 1)
    assert full >= (10 + 9) * img and lean >= (10 + 9) * img and abs(full - lean) < 4 * img
    small = NeRF("fine", W=64, encode_transient=True, in_channels_t=48, output_flow=True)
    assert L.nsff_field_bwd_tc_pack_bytes(C.byref(small.field_desc())) == 0      # tcgen05 path: 8 x 256 only


def test_tile_image_helpers_refuse_cpu_tensors():
    """nsff_pl_b200.tile_image wraps CUDA-only kernels: CPU tensors raise like every other entry point (no fallback)."""
    from nsff_pl_b200 import tile_image as TI
    with pytest.raises(NotImplementedError):
        TI.image_from_rows(torch.zeros(4, 64))
    with pytest.raises(NotImplementedError):
        TI.pack_weight(torch.zeros(256, 64))
