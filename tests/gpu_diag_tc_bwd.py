"""GPU diagnostic (not a test): per-kernel and per-parameter errors of the tensor-core backward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from nsff_pl_b200 import tile_image as TI  # noqa: E402


def bf(x):
    return x.to(torch.bfloat16).float()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(300, 200, device="cuda", generator=g)
    print("roundtrip exact:", torch.equal(TI.rows_from_image(TI.image_from_rows(x), 300, 200), bf(x)))
    for n_rows, K, N in [(300, 256, 256), (5000, 111, 256), (40000, 367, 256), (1000, 256, 128)]:
        x = torch.randn(n_rows, K, device="cuda", generator=g)
        w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
        b = torch.randn(N, device="cuda", generator=g)
        out = TI.rows_from_image(TI.gemm(TI.image_from_rows(x), (K + 63) // 64, TI.pack_weight(w, N=N), N, n_rows, bias=b),
                                 n_rows, N)
        torch.cuda.synchronize()
        print(f"gemm {n_rows}x{K}->{N}: rel err {rel(out, bf(x) @ bf(w).T + b):.3e}")
    for n_rows, nb in [(128, 4), (300, 1), (20000, 2), (131072, 4)]:
        a = torch.randn(n_rows, 256, device="cuda", generator=g)
        b = torch.randn(n_rows, 64 * nb, device="cuda", generator=g)
        acc = TI.wgrad(TI.image_from_rows(a), TI.image_from_rows(b), nb, n_rows)
        torch.cuda.synchronize()
        ref = bf(b).T @ bf(a)
        print(f"wgrad {n_rows} nb={nb}: rel err {rel(acc, ref):.3e}  (|ref| max {float(ref.abs().max()):.2f}, "
              f"acc max {float(acc.abs().max()):.2f})")
    # timing of the two kernels at the training shape
    n_rows = 131072
    a = TI.image_from_rows(torch.randn(n_rows, 256, device="cuda", generator=g))
    w = TI.pack_weight(torch.randn(256, 256, device="cuda", generator=g) / 16)
    for name, fn in (("gemm 256x256", lambda: TI.gemm(a, 4, w, 256, n_rows, epi=TI.EPI_RELU)),
                     ("wgrad 256x256", lambda: TI.wgrad(a, a, 4, n_rows))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name} @131072 rows: {ms * 1e3:.1f} us  ({2 * n_rows * 65536 / ms / 1e9:.0f} TFLOP/s, "
              f"{n_rows * 512 * 2 / ms / 1e6:.0f} GB/s moved)")
    if "--full" in sys.argv:
        from test_gpu_backward import _grads_with, TRAIN_KW
        ref = _grads_with("fp32", 24, 48, TRAIN_KW)
        got = _grads_with("bf16", 24, 48, TRAIN_KW)
        for k in ref:
            err = float((got[k] - ref[k]).norm() / (ref[k].norm() + 1e-30))
            print(f"{k:45s} rel {err:.3e}  |ref| {float(ref[k].norm()):.3e} |got| {float(got[k].norm()):.3e}")


if __name__ == "__main__":
    main()
