"""One process per GPU: tile partition and frame combination (torch.distributed plumbing).

The reference distributes 4096-pixel blocks over MPI ranks from a coordinator thread and gathers
the finished blocks on rank 0 (reference libwurblpt/mpi.hpp:42-289). Here every rank holds a
replica of the scene, renders the pixel tiles  rank, rank+N, rank+2N, ...  (inside a GPU, warps
pull their rank's tiles from a device-side atomic ticket counter), and the per-rank frames are
combined with ONE reduce over NCCL / NVLink. A pixel is written by exactly one rank and is zero
everywhere else, so the sum is an exact gather: the N-GPU frame is bit-identical to the 1-GPU frame.
"""
import os

import numpy as np

from . import capi

_TILE_W = {1: 8, 2: 4, 4: 4, 8: 2, 16: 2, 32: 1, 64: 1, 128: 1}      # capi.cu: tileWs by log2(substreams)


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend=None):
    """Initialise torch.distributed when launched by torchrun with more than one rank."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def rank_opts(rank, world, substreams=0, stream=None, traversal=0, frame_on_device=True, pipeline=0):
    """Render options that make `rank` of `world` render its interleaved share of the tiles."""
    return capi.make_opts(traversal=traversal, substreams=substreams, tile_first=rank, tile_stride=world,
                          frame_on_device=frame_on_device, stream=stream, pipeline=pipeline)


class SharedWork:
    """Dynamic work distribution over the GPUs of several processes: ONE counter in rank 0's GPU memory that the
    windows of every rank's kernels draw their blocks of work items from (CUDA IPC handle broadcast once; remote ranks
    reach it over NVLink). This is the MPICoordinator hand-out (reference mpi.hpp:74-107) without the messages: a GPU
    that is done early simply draws more. Two counters are used for alternate frames: rank 0 clears the one of the NEXT
    frame at the start of every frame, which is safe because the frame reduce separates the frames (no rank can
    start frame k+1 before rank 0 has entered the reduce of frame k, which follows the clear on its stream)."""

    def __init__(self, rank, world, device_index):
        import torch.distributed as dist
        self.rank, self.world, self.device = rank, world, device_index
        self.step = 0
        box = [None]
        if rank == 0:
            self.base, handle = capi.ipc_counter_create(device_index)
            box[0] = handle
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            self.base = capi.ipc_counter_open(device_index, box[0])

    def begin_frame(self, stream=None):
        """Counter of this frame (device pointer). Call once per frame on every rank, same number of times."""
        current = self.base + 128 * (self.step & 1)
        if self.rank == 0:
            capi.ipc_counter_clear(self.base + 128 * ((self.step + 1) & 1), stream)
        self.step += 1
        return current

    def close(self):
        capi.ipc_counter_close(self.device, self.base, self.rank != 0)
        self.base = None


def tile_shape(substreams):
    """(tile width, tile height) in pixels: a tile is 32 work items (32 / substreams pixels), or one pixel."""
    w = _TILE_W[substreams]
    return w, max(1, 32 // substreams) // w


def owner_map(width, height, substreams, world):
    """owner[y, x] = rank that renders pixel (x, y): mirror of the tiling arithmetic in wpt_render()."""
    tw, th = tile_shape(substreams)
    tiles_x = (width + tw - 1) // tw
    ys, xs = np.mgrid[0:height, 0:width]
    tile = (ys // th) * tiles_x + (xs // tw)
    return (tile % world).astype(np.int32)


def reduce_frame(frame, dst=0):
    """Combine per-rank frames (torch tensors on each rank's device) into rank `dst` with one reduce."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(frame, dst=dst, op=dist.ReduceOp.SUM)
    return frame
