"""Multi-GPU sharding of the registration.

One process per GPU.  Two ways the path shards:
  * ONE registration: the source cloud is split by contiguous point range, the
    voxel map is replicated, and the 28 partial sums (27 unique H/b terms + the
    correspondence count) are all-reduced every Gauss-Newton iteration, so every
    rank solves the identical 6x6 system and applies the identical step.
  * independent scan/map pairs: one shard of the batch per rank, no collective.
torch.distributed is plumbing only: the collective runs on the context's own
stream, on the library's device buffer (zero-copy view).
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous point range [begin, end) of `rank`; ranges tile [0, n) exactly."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(n_items: int, rank: int, world: int) -> range:
    """Indices of the independent problems (scan/map pairs) owned by `rank`."""
    b, e = shard_range(n_items, rank, world)
    return range(b, e)


class _DevView:
    """__cuda_array_interface__ view of `count` fp64 at a raw device pointer."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": (count,), "typestr": "<f8",
                                         "version": 2}


class TorchAllReduce:
    """allreduce(ptr, count, stream) callback for Map.align_cloud_sharded():
    sums the library's device buffer in place over `group` (NCCL on GPUs)."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.calls = 0

    def __call__(self, ptr: int, count: int, stream: int) -> None:
        torch, dist = self.torch, self.dist
        t = torch.as_tensor(_DevView(ptr, count), device="cuda")
        ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
        with torch.cuda.stream(ext):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        self.calls += 1


def align_sharded(gmap, cloud_shard, guess, group=None, **kw):
    """Registration of a source cloud whose point range lives on this rank."""
    import torch.distributed as dist
    cb = TorchAllReduce(group) if dist.is_initialized() and dist.get_world_size(group) > 1 else None
    return gmap.align_cloud_sharded(cloud_shard, np.asarray(guess, dtype=np.float64), cb, **kw)


def torch_all_gather_bytes(group=None):
    """all_gather callable for capi.Comm: exchanges the 64-byte IPC handles over the process
    group (works with NCCL and gloo; plumbing only, called once at set-up)."""
    import torch
    import torch.distributed as dist

    def gather(mine: bytes):
        world = dist.get_world_size(group)
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t, group=group)
        return [bytes(o.cpu().tolist()) for o in out]

    return gather


def make_comm(ctx, group=None):
    """capi.Comm over the ranks of `group` (one process per GPU)."""
    import torch.distributed as dist
    from . import capi
    if not dist.is_initialized():
        return capi.Comm(ctx, 0, 1)
    return capi.Comm(ctx, dist.get_rank(group), dist.get_world_size(group), torch_all_gather_bytes(group))
