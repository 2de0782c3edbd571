"""CPU, world_size 2, gloo: the N>1 host path.  Each rank renders its contiguous ray block (with the oracle -
there is no GPU here), all-gathers the packed pixels through nsff_pl_b200.parallel, and the gathered frame
must equal the single-process render bit for bit (rays are independent)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays, ret):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nsff_pl_b200 import parallel
        from oracle import nsff_oracle as O
        from parity_util import build_fields, synthetic_rays
        _, _, specs, enc = build_fields(W=256, seed=4, n_frames=8, device=None)
        rays, ts = synthetic_rays(n_rays, n_frames=8, seed=2)
        kw = dict(N_samples=8, test_time=True, output_transient_flow=["fw", "bw"])

        def render(r, t):
            with torch.no_grad():
                return O.render_rays(specs, enc, r, t, 7, **kw)

        frame = parallel.render_frame_sharded(render, rays, ts)
        if rank == 0:
            full = render(rays, ts)
            ok = all(torch.equal(frame[k], full[k]) for k in parallel.PIXEL_KEYS)
            shapes = {k: tuple(frame[k].shape) for k in parallel.PIXEL_KEYS}
            ret["ok"], ret["shapes"] = ok, shapes
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [16, 13])       # equal shards (one all_gather) and ragged shards (padded)
def test_sharded_render_equals_single_process(n_rays):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n_rays, ret), nprocs=2, join=True)
    assert ret["ok"] is True
    assert ret["shapes"]["rgb_fine"] == (n_rays, 3) and ret["shapes"]["depth_fine"] == (n_rays,)


def _grad_worker(rank, world, port, ret):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nsff_pl_b200 import parallel
        torch.manual_seed(0)                                     # replicated parameters
        params = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7)),
                  torch.nn.Parameter(torch.randn(2, 2), requires_grad=False), torch.nn.Parameter(torch.randn(4))]
        g = torch.Generator().manual_seed(100 + rank)            # per-rank ray batch -> per-rank gradients
        params[0].grad = torch.randn(5, 3, generator=g)
        params[1].grad = torch.randn(7, generator=g)
        if rank == 0:                                            # a head that only rank 0 exercised
            params[3].grad = torch.randn(4, generator=g)
        n = parallel.all_reduce_gradients(params)
        ret[rank] = (n, [None if p.grad is None else p.grad.clone() for p in params])
    finally:
        dist.destroy_process_group()


def test_gradient_all_reduce_averages_over_ranks():
    """Training N>1: one flat all-reduce, gradients averaged, frozen parameters untouched,
    a parameter that received a gradient on one rank only is averaged with zeros on the others."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grad_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    (n0, g0), (n1, g1) = ret[0], ret[1]
    assert n0 == n1 == 15 + 7 + 4
    want0, want1, want3 = [], [], []
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        want0.append(torch.randn(5, 3, generator=g))
        want1.append(torch.randn(7, generator=g))
        want3.append(torch.randn(4, generator=g) if r == 0 else torch.zeros(4))
    for got in (g0, g1):
        assert torch.allclose(got[0], sum(want0) / world) and torch.allclose(got[1], sum(want1) / world)
        assert got[2] is None
        assert torch.allclose(got[3], sum(want3) / world)


def _arena_grad_worker(rank, world, port, ret):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nsff_pl_b200 import parallel
        params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(6))]
        g = torch.Generator().manual_seed(200 + rank)
        flat = torch.randn(16 + 8, generator=g)                   # 16-byte slots like models/autograd.py: 15 -> 16, 7 -> 8
        params[0].grad = flat[0:15].view(5, 3)
        params[1].grad = flat[16:23]
        params[2].grad = torch.randn(6, generator=g)              # a gradient that lives elsewhere (the time-code table)
        n = parallel.all_reduce_gradients(params)
        ret[rank] = (n, [p.grad.clone() for p in params], params[0].grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr())
    finally:
        dist.destroy_process_group()


def test_gradient_all_reduce_reduces_a_shared_gradient_buffer_in_place():
    """Gradients that tile one storage (the per-model buffer the backward kernels fill) are reduced through one flat view of
    it, in place; the rest goes through the bucket."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_arena_grad_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    flats, others = [], []
    for r in range(world):
        g = torch.Generator().manual_seed(200 + r)
        flats.append(torch.randn(24, generator=g))
        others.append(torch.randn(6, generator=g))
    want = sum(flats) / world
    for r in range(world):
        n, got, in_place = ret[r]
        assert n == 23 + 6 and in_place
        assert torch.allclose(got[0].reshape(-1), want[0:15]) and torch.allclose(got[1], want[16:23])
        assert torch.allclose(got[2], sum(others) / world)


def _gather_worker(rank, world, port, ret):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nsff_pl_b200 import parallel
        n = 12
        g = parallel.PixelGather(n, "cpu")
        ok = True
        for frame in range(3):                       # double-buffered: three frames reuse slot 0
            full_res = {"rgb_fine": torch.arange(n * 3, dtype=torch.float32).view(n, 3) + 100 * frame,
                        "depth_fine": torch.arange(n, dtype=torch.float32) * 0.5 + frame,
                        "transient_flow_fw": torch.ones(n, 3) * frame, "transient_flow_bw": -torch.ones(n, 3) * frame}
            lo, hi = parallel.shard_bounds(n, rank, world)
            buf = g.submit({k: v[lo:hi] for k, v in full_res.items()})
            g.wait(buf)
            back = parallel.unpack_pixels(buf, full_res)
            ok = ok and all(torch.equal(back[k], full_res[k]) for k in parallel.PIXEL_KEYS)
        if rank == 0:
            ret["ok"] = ok
            try:
                parallel.PixelGather(13, "cpu")
                ret["ragged_raises"] = False
            except ValueError:
                ret["ragged_raises"] = True
    finally:
        dist.destroy_process_group()


def test_pixel_gather_in_place_double_buffered():
    """config 4 plumbing: each rank writes its rows of the pre-allocated frame buffer, in-place all-gather."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["ok"] is True and ret["ragged_raises"] is True
