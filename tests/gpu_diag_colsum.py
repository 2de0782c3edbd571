import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from nsff_pl_b200 import tile_image as TI

def bf(x): return x.to(torch.bfloat16).float()
g = torch.Generator(device="cuda").manual_seed(0)
for n_rows in (384, 407, 5000):
    for K in (10, 256):
        x = torch.randn(n_rows, K, device="cuda", generator=g)
        w = torch.randn(256, K, device="cuda", generator=g)
        a = TI.image_from_rows(x); blocks = TI.pack_weight(w, N=256)
        ref = bf(x) @ bf(w).T
        for epi, name in ((TI.EPI_LINEAR, "lin"), (TI.EPI_RELU, "relu")):
            cs = torch.zeros(256, device="cuda")
            img = TI.gemm(a, (K + 63) // 64, blocks, 256, n_rows, epi=epi, colsum=cs)
            out = TI.rows_from_image(img, n_rows, 256)
            want = ref if epi == TI.EPI_LINEAR else ref.clamp_min(0)
            e = (cs - want.sum(0))
            per = [float(e[i*32:(i+1)*32].norm() / want.sum(0)[i*32:(i+1)*32].norm()) for i in range(8)]
            print(f"n={n_rows} K={K} {name}: colsum rel {float(e.norm()/want.sum(0).norm()):.3e}  image-sum rel "
                  f"{float((out.sum(0)-want.sum(0)).norm()/want.sum(0).norm()):.3e}  per32 {[f'{p:.1e}' for p in per]}")
