import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(__file__)); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from nsff_pl_b200.models.rendering import render_rays
from parity_util import build_fields, synthetic_rays
models, emb, _, _ = build_fields(W=256, seed=0, n_frames=30, precision="bf16")
for m in models.values(): m.requires_grad_(True)
emb["t"].weight.requires_grad_(True)
params = [p for m in models.values() for p in m.parameters()] + [emb["t"].weight]
opt = torch.optim.Adam(params, lr=5e-4)
rays, ts = synthetic_rays(1024, n_frames=30, seed=100); rays, ts = rays.cuda(), ts.cuda()
target = torch.rand(1024, 3, device="cuda")
kw = dict(N_samples=128, perturb=1.0, noise_std=1.0, N_importance=0, test_time=False, output_transient_flow=["fw", "bw", "disocc"])
for it in range(12):
    torch.cuda.synchronize(); t0 = time.time()
    opt.zero_grad(set_to_none=True)
    res = render_rays(models, emb, rays, ts, 29, **kw)
    loss = ((res["rgb_fine"] - target) ** 2).mean() + (res["disocc_fw"] * (res["rgb_fw"] - target) ** 2).mean()
    torch.cuda.synchronize(); t1 = time.time()
    loss.backward()
    torch.cuda.synchronize(); t2 = time.time()
    opt.step()
    torch.cuda.synchronize(); t3 = time.time()
    st = torch.cuda.memory_stats()
    print(f"it {it}: fwd {1e3*(t1-t0):.2f} bwd {1e3*(t2-t1):.2f} opt {1e3*(t3-t2):.2f} ms  alloc {torch.cuda.memory_allocated()/2**30:.2f} GB reserved {torch.cuda.memory_reserved()/2**30:.2f} GB  cudaMallocs {st['num_device_alloc']} frees {st['num_device_free']}")
