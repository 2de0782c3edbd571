"""Multi-GPU plumbing for the render path: rays are independent, weights are replicated,
so a frame shards into contiguous ray blocks (one per rank, pixel order preserved) and the only collective
is ONE all-gather of the packed per-ray outputs.  One process per GPU, torch.distributed (NCCL on GPUs;
gloo in the CPU tests).  Nothing here touches the kernels.
"""
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist

PIXEL_KEYS = ("rgb_fine", "depth_fine", "transient_flow_fw", "transient_flow_bw")


def shard_bounds(n_rays: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one ray and concatenate in order."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_rays, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_pixels(results: Dict[str, torch.Tensor], keys: Sequence[str] = PIXEL_KEYS) -> torch.Tensor:
    """(R, C) fp32 buffer of the per-ray outputs named in `keys` (each (R,) or (R,c))."""
    cols = [results[k].reshape(results[k].shape[0], -1) for k in keys]
    return torch.cat(cols, 1).contiguous()


def unpack_pixels(buf: torch.Tensor, like: Dict[str, torch.Tensor], keys: Sequence[str] = PIXEL_KEYS):
    out, c0 = {}, 0
    for k in keys:
        shp = like[k].shape[1:]
        n = 1
        for s in shp:
            n *= s
        out[k] = buf[:, c0:c0 + n].reshape(buf.shape[0], *shp)
        c0 += n
    return out


def all_gather_pixels(local: torch.Tensor, n_rays_total: int, group=None) -> torch.Tensor:
    """Gathers the ranks' (R_rank, C) blocks into the full (n_rays_total, C) frame buffer on every rank.
    Equal shards use a single all_gather_into_tensor; ragged shards pad to the largest block."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    C = local.shape[1]
    sizes = [shard_bounds(n_rays_total, r, world) for r in range(world)]
    lens = [hi - lo for lo, hi in sizes]
    if local.shape[0] != lens[rank]:
        raise ValueError(f"rank {rank}: expected {lens[rank]} rays, got {local.shape[0]}")
    if len(set(lens)) == 1:
        full = torch.empty(n_rays_total, C, device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(full, local.contiguous(), group=group)
        return full
    m = max(lens)
    padded = torch.zeros(m, C, device=local.device, dtype=local.dtype)
    padded[:local.shape[0]] = local
    gathered = torch.empty(world * m, C, device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    return torch.cat([gathered[r * m:r * m + lens[r]] for r in range(world)], 0)


def render_frame_sharded(render_fn, rays: torch.Tensor, ts, keys: Sequence[str] = PIXEL_KEYS, group=None):
    """Config 4: every rank renders its contiguous ray block with `render_fn(rays_block, ts_block) -> dict`
    and all ranks end up with the whole frame's per-ray outputs."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(rays.shape[0], rank, world)
    res = render_fn(rays[lo:hi].contiguous(), None if ts is None else ts[lo:hi].contiguous())
    full = all_gather_pixels(pack_pixels(res, keys), rays.shape[0], group)
    return unpack_pixels(full, res, keys)


def all_reduce_gradients(params: Sequence[torch.Tensor], group=None) -> int:
    """Data-parallel training (the reference's DDP, train.py:294-301): average the gradients of
    `params` over the ranks with ONE all-reduce of a flat fp32 bucket (2.4 M parameters = 9.8 MB for fine + coarse +
    time codes).  Parameters without a gradient on this rank contribute zeros - every rank must pass the same list.
    Returns the number of elements reduced."""
    world = dist.get_world_size(group)
    params = [p for p in params if p.requires_grad]
    if world == 1 or not params:
        return 0
    ref = next((p.grad for p in params if p.grad is not None), None)
    dev = ref.device if ref is not None else params[0].device
    flat = torch.zeros(sum(p.numel() for p in params), device=dev, dtype=torch.float32)
    o = 0
    for p in params:
        n = p.numel()
        if p.grad is not None:
            flat[o:o + n].copy_(p.grad.reshape(-1))
        o += n
    dist.all_reduce(flat, group=group)
    flat /= world
    o = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[o:o + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[o:o + n].view_as(p.grad))
        o += n
    return int(flat.numel())
