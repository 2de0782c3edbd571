"""Unit tests of the two tcgen05 GEMM kernels behind the tensor-core backward (csrc/nsff_field_bwd_tc.cu), through
the C ABI, against fp32 torch matmuls of the same bf16-rounded operands."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).float()


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


@pytest.mark.parametrize("n_rows", [128, 300, 5000])
def test_image_roundtrip(n_rows):
    from nsff_pl_b200 import tile_image as TI
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n_rows, 200, device="cuda", generator=g)
    img = TI.image_from_rows(x)
    back = TI.rows_from_image(img, n_rows, 200)
    assert torch.equal(back, _bf(x))


@pytest.mark.parametrize("n_rows,K,N", [(300, 256, 256), (5000, 111, 256), (40000, 367, 256), (1000, 256, 128)])
def test_gemm_kmajor_all_epilogues(n_rows, K, N):
    from nsff_pl_b200 import tile_image as TI
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(n_rows, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    a_img = TI.image_from_rows(x)
    chunks = (K + 63) // 64
    blocks = TI.pack_weight(w, N=N)
    ref = _bf(x) @ _bf(w).T + b
    lin = TI.rows_from_image(TI.gemm(a_img, chunks, blocks, N, n_rows, bias=b, epi=TI.EPI_LINEAR), n_rows, N)
    assert _rel(lin, ref) < 1e-2, _rel(lin, ref)
    relu = TI.rows_from_image(TI.gemm(a_img, chunks, blocks, N, n_rows, bias=b, epi=TI.EPI_RELU), n_rows, N)
    assert _rel(relu, ref.clamp_min(0)) < 1e-2
    m = torch.randn(n_rows, N, device="cuda", generator=g).clamp_min(0)
    mk = TI.rows_from_image(TI.gemm(a_img, chunks, blocks, N, n_rows, epi=TI.EPI_RELU_MASK,
                                    mask_img=TI.image_from_rows(m)), n_rows, N)
    assert _rel(mk, (ref - b) * (_bf(m) > 0)) < 1e-2
    # bit-plane route: the relu pass records (out > 0); a masked pass through those bits equals the image-masked one,
    # and its fused column sums equal the sums of the (pre-rounding) masked result
    bits = TI.bit_plane(n_rows, "cuda")
    TI.gemm(a_img, chunks, blocks, N, n_rows, bias=b, epi=TI.EPI_RELU, relu_bits=bits)
    n_tiles = (n_rows + 127) // 128
    words = bits.view(n_tiles, 8, 128).permute(0, 2, 1).reshape(n_tiles * 128, 8)[:n_rows]       # (row, word)
    on = ((words[:, :, None] >> torch.arange(32, device="cuda")) & 1).reshape(n_rows, 256)[:, :N].bool()
    # the recorded bits are the sign of the kernel's own fp32 sums: equal to (ref > 0) except within rounding of zero
    assert bool((on == (ref > 0))[ref.abs() > 1e-3].all())
    cs = torch.zeros(N, device="cuda")
    mk2 = TI.rows_from_image(TI.gemm(a_img, chunks, blocks, N, n_rows, epi=TI.EPI_BIT_MASK, mask_bits=bits, colsum=cs),
                             n_rows, N)
    want2 = (ref - b) * on
    assert _rel(mk2, want2) < 1e-2
    assert float((cs - want2.sum(0)).abs().max()) <= 2e-3 * float(want2.abs().sum(0).max()), (cs, want2.sum(0))
    rows = torch.full((n_rows, N + 5), 7.0, device="cuda")
    TI.gemm(a_img, chunks, blocks, N, n_rows, epi=TI.EPI_ROWS, out_rows=rows, ncols=N - 3, beta=0)
    assert _rel(rows[:, :N - 3], (ref - b)[:, :N - 3]) < 1e-2 and bool((rows[:, N - 3:] == 7.0).all())
    TI.gemm(a_img, chunks, blocks, N, n_rows, epi=TI.EPI_ROWS, out_rows=rows, ncols=N - 3, beta=1)
    assert _rel(rows[:, :N - 3], 2 * (ref - b)[:, :N - 3]) < 1.5e-2
    rows4 = torch.zeros(n_rows, N, device="cuda")                       # 16-byte aligned rows: vectorised stores
    TI.gemm(a_img, chunks, blocks, N, n_rows, epi=TI.EPI_ROWS, out_rows=rows4, ncols=N - 17, beta=0)
    TI.gemm(a_img, chunks, blocks, N, n_rows, epi=TI.EPI_ROWS, out_rows=rows4, ncols=N - 17, beta=1)
    assert _rel(rows4[:, :N - 17], 2 * (ref - b)[:, :N - 17]) < 1.5e-2 and bool((rows4[:, N - 17:] == 0).all())


def test_gemm_transposed_weight_is_dgrad():
    from nsff_pl_b200 import tile_image as TI
    g = torch.Generator(device="cuda").manual_seed(3)
    dz = torch.randn(3000, 256, device="cuda", generator=g)
    w = torch.randn(256, 256, device="cuda", generator=g) / 16
    out = TI.rows_from_image(TI.gemm(TI.image_from_rows(dz), 4, TI.pack_weight(w, transpose=True), 256, 3000), 3000, 256)
    assert _rel(out, _bf(dz) @ _bf(w)) < 1e-2


@pytest.mark.parametrize("n_rows,nb", [(128, 4), (300, 1), (20000, 2), (131072, 4)])
def test_wgrad_mnmajor(n_rows, nb):
    from nsff_pl_b200 import tile_image as TI
    g = torch.Generator(device="cuda").manual_seed(4)
    a = torch.randn(n_rows, 256, device="cuda", generator=g)
    b = torch.randn(n_rows, 64 * nb, device="cuda", generator=g)
    acc = TI.wgrad(TI.image_from_rows(a), TI.image_from_rows(b), nb, n_rows)
    ref = _bf(b).T @ _bf(a)
    assert _rel(acc, ref) < 2e-3, _rel(acc, ref)
