"""CPU tests of the multi-GPU host logic: the tile partition is an exact cover, and combining the
per-rank frames with one reduce is an exact gather (run with the gloo backend, world size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("substreams", [1, 2, 4, 8, 16, 32, 64])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_owner_map_is_an_exact_cover(substreams, world):
    from wurblpt_b200 import multi
    w, h = 101, 67
    owner = multi.owner_map(w, h, substreams, world)
    assert owner.shape == (h, w)
    assert owner.min() == 0 and owner.max() == min(world, owner.max() + 1) - 1 or world == 1
    counts = np.bincount(owner.ravel(), minlength=world)
    assert counts.sum() == w * h
    if world > 1:
        # interleaved tiles: every rank gets its fair share within one tile row
        tw, th = multi.tile_shape(substreams)
        assert counts.max() - counts.min() <= 2 * tw * th * ((h + th - 1) // th)
    tw, th = multi.tile_shape(substreams)
    assert tw * th * substreams == max(32, substreams)


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from wurblpt_b200 import multi
    r, w, _ = multi.init_process_group(backend="gloo")
    assert (r, w) == (rank, world)
    width, height, sub = 64, 48, 2
    owner = torch.from_numpy(multi.owner_map(width, height, sub, world))
    rng = np.random.default_rng(7)
    full = torch.from_numpy(rng.random((height, width, 3)).astype(np.float32))
    # what rank r would have rendered: its own pixels, zero elsewhere
    frame = torch.where((owner == rank).unsqueeze(-1), full, torch.zeros_like(full))
    opts = multi.rank_opts(rank, world, substreams=sub)
    assert (opts.tile_first, opts.tile_stride) == (rank, world)
    multi.reduce_frame(frame, dst=0)
    if rank == 0:
        out.put(bool(torch.equal(frame, full)))     # bit-exact: x + 0 == x
    dist.destroy_process_group()


def test_gloo_reduce_is_an_exact_gather():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
