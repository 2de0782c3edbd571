"""Shared helpers for the GPU parity tests, smoke() and bench.py: build the same seeded field on both
sides (nsff_pl_b200 NeRF on CUDA, oracle FieldSpec on CPU), run both render paths, compare dicts."""
import torch

from oracle import nsff_oracle as O
from scenes import energise_, synthetic_rays  # noqa: F401

N_TAU = 48


def build_fields(W=256, coarse=False, transient=True, appearance=False, viewdir=True, n_frames=30, seed=0,
                 device="cuda", precision="bf16"):
    """Returns (models, embeddings) for nsff_pl_b200 and (specs, enc) for the oracle, sharing weights."""
    from nsff_pl_b200.models.nerf import NeRF, PosEmbedding
    torch.manual_seed(seed)
    models = {"fine": NeRF("fine", D=8, W=W, skips=[4], in_channels_xyz=63, in_channels_dir=27, use_viewdir=viewdir,
                           encode_appearance=appearance, in_channels_a=48, encode_transient=transient,
                           in_channels_t=N_TAU, output_flow=transient, flow_scale=0.2)}
    if coarse:
        models["coarse"] = NeRF("coarse", D=8, W=W, skips=[4], in_channels_xyz=63, in_channels_dir=27,
                                use_viewdir=viewdir, encode_transient=transient, in_channels_t=N_TAU)
    emb = {"xyz": PosEmbedding(9, 10), "dir": PosEmbedding(3, 4)}
    if transient:
        emb["t"] = torch.nn.Embedding(n_frames, N_TAU)
    if appearance:
        emb["a"] = torch.nn.Embedding(n_frames, 48)
    for m in models.values():
        energise_(m)
        m.requires_grad_(False)
        m.precision = precision
    specs = {k: O.FieldSpec.from_module(m) for k, m in models.items()}
    enc = O.Encoders(O.pos_freqs(9, 10), O.pos_freqs(3, 4),
                     t_table=emb["t"].weight.detach().clone() if transient else None,
                     a_table=emb["a"].weight.detach().clone() if appearance else None)
    if device is not None:
        for m in models.values():
            m.to(device)
        for k in ("t", "a"):
            if k in emb:
                emb[k].to(device).requires_grad_(False)
    return models, emb, specs, enc


def max_abs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max()) if a.numel() else 0.0


def psnr(a, b):
    mse = float(((a.detach().double().cpu() - b.detach().double().cpu()) ** 2).mean())
    return 99.0 if mse == 0 else -10.0 * torch.log10(torch.tensor(mse)).item()


def compare_dicts(got, want, tol, default_tol=None, check_order=True):
    """tol: {key or key-substring: tolerance (None = do not bound)}; returns {key: max_abs_err}; raises AssertionError on violation."""
    if check_order:
        assert list(got.keys()) == list(want.keys()), f"result keys differ:\n got {list(got)}\nwant {list(want)}"
    errs = {}
    bad = []
    for k, w in want.items():
        g = got[k]
        assert tuple(g.shape) == tuple(w.shape), f"{k}: shape {tuple(g.shape)} != {tuple(w.shape)}"
        assert torch.isfinite(g).all(), f"{k}: non-finite values"
        e = max_abs(g, w)
        errs[k] = e
        t = default_tol
        best = -1
        for sub, v in tol.items():          # exact name wins, else the longest matching substring
            score = 10_000 if sub == k else (len(sub) if sub in k else -1)
            if score > best:
                best, t = score, v
        scale = max(1.0, float(w.abs().max())) if w.numel() else 1.0     # tolerances are relative to the
        if t is not None and e > t * scale:                                 # tensor's magnitude when > 1
            bad.append(f"{k}: max abs err {e:.3e} > {t:.1e} * {scale:.2f}")
    assert not bad, "\n".join(bad)
    return errs
