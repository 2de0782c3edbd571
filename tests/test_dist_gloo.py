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
