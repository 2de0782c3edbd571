"""bf16 "tile images" - the HBM layout of the tensor-core backward (csrc/nsff_field_bwd_tc.cu) - and thin wrappers
over its two GEMM kernels.  Used by the unit tests and available to hosts that want the kernels on their own data.

An image of an (n_rows, 64*chunks) matrix is [ceil(n_rows/128) tiles][chunks][128 rows x 64 bf16, 128-byte swizzled]:
exactly the shared-memory operand layout of tcgen05.mma, so operand loads are plain bulk copies."""
import torch

from . import _cabi

EPI_RELU, EPI_LINEAR, EPI_RELU_MASK, EPI_ROWS, EPI_BIT_MASK = 0, 1, 2, 3, 4


def _st(t):
    return _cabi.stream_ptr(t.device)


def image_from_rows(x):
    """(n_rows, C) fp32 CUDA -> image with ceil(C/64) chunks (uint8 tensor)."""
    x = _cabi.dev_tensor(x.contiguous(), "x")
    n, c = x.shape
    chunks = (c + 63) // 64
    img = torch.empty(_cabi.lib().nsff_img_bytes(n, chunks), dtype=torch.uint8, device=x.device)
    _cabi.select_device(x)
    for j in range(chunks):
        _cabi.check(_cabi.lib().nsff_img_from_rows(_cabi.ptr(x), c, 64 * j, min(64, c - 64 * j), n, _cabi.ptr(img),
                                                   chunks, j, _st(x)), "nsff_img_from_rows")
    return img


def rows_from_image(img, n_rows, cols):
    chunks = (cols + 63) // 64
    out = torch.empty(n_rows, cols, dtype=torch.float32, device=img.device)
    _cabi.select_device(img)
    for j in range(chunks):
        _cabi.check(_cabi.lib().nsff_img_to_rows(_cabi.ptr(img), chunks, j, _cabi.ptr(out), cols, 64 * j,
                                                 min(64, cols - 64 * j), n_rows, _st(img)), "nsff_img_to_rows")
    return out


def pack_weight(w, N=None, transpose=False):
    """nn.Linear weight (out, in) -> ceil(K/64) blocks of N x 64 with blk[n][k] = w[n, k] (or w[k, n] if transpose)."""
    w = _cabi.dev_tensor(w.contiguous(), "w")
    rows, cols = w.shape
    n_dim, k_dim = (cols, rows) if transpose else (rows, cols)
    N = N or n_dim
    kc = (k_dim + 63) // 64
    blocks = torch.empty(kc * N * 128, dtype=torch.uint8, device=w.device)
    _cabi.select_device(w)
    for j in range(kc):
        _cabi.check(_cabi.lib().nsff_img_pack_weight(_cabi.ptr(w), cols, int(transpose), 0, 64 * j, n_dim,
                                                     min(64, k_dim - 64 * j), N,
                                                     blocks.data_ptr() + j * N * 128, _st(w)), "nsff_img_pack_weight")
    return blocks


def bit_plane(n_rows, device):
    """Empty ReLU bit plane for an (n_rows, <=256) activation: per tile [8 words][128 rows] uint32."""
    return torch.zeros(((n_rows + 127) // 128) * 8 * 128, dtype=torch.int32, device=device)


def gemm(a_img, a_chunks, blocks, N, n_rows, bias=None, epi=EPI_LINEAR, mask_img=None, out_rows=None, ncols=None,
         beta=0, relu_bits=None, mask_bits=None, colsum=None):
    """out = epi(A . B^T + bias): returns an image (epi 0..2, 4) or writes fp32 rows into out_rows (epi 3)."""
    args = _cabi.ImgGemmArgs()
    args.a_img, args.a_chunks, args.b_blocks, args.N = a_img.data_ptr(), a_chunks, blocks.data_ptr(), N
    args.bias = bias.data_ptr() if bias is not None else None
    args.epi, args.n_rows = epi, n_rows
    out = None
    if epi == EPI_ROWS:
        args.out_rows, args.ld, args.ncols, args.beta = out_rows.data_ptr(), out_rows.stride(0), ncols, beta
    else:
        out = torch.empty(_cabi.lib().nsff_img_bytes(n_rows, N // 64), dtype=torch.uint8, device=a_img.device)
        args.out_img = out.data_ptr()
        args.mask_img = mask_img.data_ptr() if mask_img is not None else None
        args.relu_bits = relu_bits.data_ptr() if relu_bits is not None else None
        args.mask_bits = mask_bits.data_ptr() if mask_bits is not None else None
        args.colsum = colsum.data_ptr() if colsum is not None else None
    _cabi.select_device(a_img)
    _cabi.check(_cabi.lib().nsff_img_gemm(args, _st(a_img)), "nsff_img_gemm")
    return out if out is not None else out_rows


def wgrad(a_img, b_img, b_chunks, n_rows):
    """(64*b_chunks, 256) fp32 = B^T . A over the rows (A: 4-chunk image)."""
    acc = torch.zeros(64 * b_chunks, 256, dtype=torch.float32, device=a_img.device)
    args = _cabi.ImgWgradArgs()
    args.a_img, args.b_img, args.b_chunks, args.n_rows, args.acc_t = (a_img.data_ptr(), b_img.data_ptr(), b_chunks,
                                                                     n_rows, acc.data_ptr())
    _cabi.select_device(a_img)
    _cabi.check(_cabi.lib().nsff_img_wgrad(args, _st(a_img)), "nsff_img_wgrad")
    return acc
