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


class PixelGather:
    """Pre-allocated, double-buffered all-gather of a frame's packed per-ray outputs (config 4: ONE frame sharded
    over the ranks).  Each rank writes its pixel columns straight into its own rows of the full-frame buffer (no
    torch.cat, no per-step allocation) and the collective runs in place on a side stream, so the all-gather of frame i
    overlaps the field kernel of frame i+1; `wait()` makes the current stream wait for a frame's gather.

        g = PixelGather(n_rays_total, device)
        for frame ...:
            res = render_rays(...rays[lo:hi]...)
            buf = g.submit(res)          # returns immediately; buf is valid after g.wait(buf)
        g.wait(buf)
    """

    def __init__(self, n_rays_total: int, device, keys: Sequence[str] = PIXEL_KEYS, n_cols: int = 10, group=None,
                 depth: int = 2):
        self.group, self.keys, self.n_total = group, tuple(keys), int(n_rays_total)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.lo, self.hi = shard_bounds(self.n_total, self.rank, self.world)
        lens = [b - a for a, b in (shard_bounds(self.n_total, r, self.world) for r in range(self.world))]
        if len(set(lens)) != 1:
            raise ValueError("PixelGather needs equal ray blocks (n_rays_total divisible by the world size); "
                             "use all_gather_pixels() for ragged shards")
        self.bufs = [torch.empty(self.n_total, n_cols, device=device, dtype=torch.float32) for _ in range(depth)]
        self.done = [None] * depth
        self.i = 0
        self.cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None

    def submit(self, results: Dict[str, torch.Tensor]) -> torch.Tensor:
        slot = self.i % len(self.bufs)
        self.i += 1
        full = self.bufs[slot]
        if self.cuda and self.done[slot] is not None:
            torch.cuda.current_stream().wait_event(self.done[slot])       # the slot's previous gather has finished
        mine, c0 = full[self.lo:self.hi], 0
        for k in self.keys:                                                # pixel columns, in place
            v = results[k].reshape(results[k].shape[0], -1)
            mine[:, c0:c0 + v.shape[1]].copy_(v)
            c0 += v.shape[1]
        if not self.cuda:
            dist.all_gather_into_tensor(full, mine, group=self.group)
            return full
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            dist.all_gather_into_tensor(full, mine, group=self.group)     # in place: `mine` is this rank's block of `full`
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.done[slot] = ev
        return full

    def wait(self, full: torch.Tensor = None):
        """Current stream waits for the gather of `full` (default: every gather in flight)."""
        if not self.cuda:
            return
        for b, ev in zip(self.bufs, self.done):
            if ev is not None and (full is None or b is full):
                torch.cuda.current_stream().wait_event(ev)


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
    `params` over the ranks with ONE all-reduce per flat gradient buffer (the per-model buffer the backward kernels fill,
    reduced in place) plus ONE flat fp32 bucket for whatever lives elsewhere (2.4 M parameters = 9.8 MB for fine + coarse +
    time codes).  Parameters without a gradient on this rank contribute zeros - every rank must pass the same list.
    Returns the number of elements reduced."""
    world = dist.get_world_size(group)
    params = [p for p in params if p.requires_grad]
    if world == 1 or not params:
        return 0
    # Gradients that already tile one storage densely (the per-model gradient buffer of models/autograd.py) are reduced
    # in place through one flat view of that range - no bucket copies; everything else goes through the flat bucket.
    n_reduced = 0
    rest = params
    if all(p.grad is not None for p in params):
        by_storage = {}
        for p in params:
            g = p.grad
            if g.is_contiguous() and g.dtype == torch.float32:
                by_storage.setdefault((g.untyped_storage().data_ptr(), g.device), []).append(p)
        rest = []
        taken = set()
        for (_, dev), ps in by_storage.items():
            lo = min(p.grad.storage_offset() for p in ps)
            hi = max(p.grad.storage_offset() + p.grad.numel() for p in ps)
            if len(ps) < 2 or sum(p.grad.numel() for p in ps) < 0.9 * (hi - lo):
                continue
            flat = torch.empty(0, dtype=torch.float32, device=dev).set_(ps[0].grad.untyped_storage(), lo, (hi - lo,))
            dist.all_reduce(flat, group=group)
            flat /= world
            n_reduced += int(flat.numel())
            taken.update(id(p) for p in ps)
        rest = [p for p in params if id(p) not in taken]
        if not rest:
            return n_reduced
    params = rest
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
    return n_reduced + int(flat.numel())
