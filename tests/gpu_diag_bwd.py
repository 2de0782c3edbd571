"""Diagnostic: per-parameter relative gradient error of the CUDA backward vs oracle autograd (fp32 and fp64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from test_gpu_backward import _run, _loss, TRAIN_KW
from parity_util import synthetic_rays
from oracle import nsff_oracle as O

for prec, R, S in (("fp32", 12, 32), ("fp32", 48, 64)):
    models, emb, specs, enc, loss, loss_ref, res, want = _run(prec, R, S, TRAIN_KW)
    # fp64 oracle
    rays, ts = synthetic_rays(R, n_frames=8, seed=R)
    torch.set_default_dtype(torch.float64)
    specs64 = {k: O.FieldSpec(**{**sp.__dict__, "params": {n: v.detach().double().requires_grad_(True) for n, v in sp.params.items()}}) for k, sp in specs.items()}
    enc64 = O.Encoders(O.pos_freqs(9, 10).double(), O.pos_freqs(3, 4).double(), t_table=enc.t_table.detach().double().requires_grad_(True))
    want64 = O.render_rays(specs64, enc64, rays.double(), ts, 7, N_samples=S, **TRAIN_KW)
    l64 = _loss({k: v.float() if False else v for k, v in want64.items()})
    l64.backward()
    torch.set_default_dtype(torch.float32)
    print(f"=== {prec} R={R} S={S} loss {loss:.6f} ref32 {loss_ref:.6f} ref64 {float(l64):.6f}")
    for name, p in models["fine"].named_parameters():
        r32 = specs["fine"].params[name].grad.double()
        r64 = specs64["fine"].params[name].grad
        got = p.grad.detach().cpu().double()
        e_mine = float((got - r64).norm() / (r64.norm() + 1e-12))
        e_or32 = float((r32 - r64).norm() / (r64.norm() + 1e-12))
        flag = "  <<<" if e_mine > 5 * e_or32 + 1e-5 else ""
        print(f"  {name:40s} |g| {float(r64.norm()):9.3e}  mine-vs-64 {e_mine:9.3e}  oracle32-vs-64 {e_or32:9.3e}{flag}")
