"""Host-side partitioning for the multi-GPU paths (one process per GPU).

* pairs / frames / query row-blocks: contiguous, balanced ranges, no collective;
* bundle adjustment: observations are sharded by residual; the per-camera normal-equation blocks
  ``[n_cam x 10]`` (Hxx,Hxy,Hxz,Hyy,Hyz,Hzz,gx,gy,gz,cost) are additive over shards, so one all-reduce
  (sum) per LM iteration makes every rank take the same step.
"""
from __future__ import annotations

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Balanced contiguous range [lo, hi) of n items for `rank` (the first n % world ranks get one more)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_pairs(n_frames: int) -> list[tuple[int, int]]:
    """Unordered frame pairs (i < j) of an n-frame sequence in row-major order (BASELINE config 3)."""
    return [(i, j) for i in range(n_frames) for j in range(i + 1, n_frames)]


def shard_pairs(n_frames: int, rank: int, world: int) -> list[tuple[int, int]]:
    pairs = all_pairs(n_frames)
    lo, hi = shard_range(len(pairs), rank, world)
    return pairs[lo:hi]


def pack_blocks(H: np.ndarray, g: np.ndarray, cost: np.ndarray) -> np.ndarray:
    """[n_cam x 6], [n_cam x 3], [n_cam] -> the [n_cam x 10] layout the device all-reduces."""
    return np.concatenate([H, g, cost[:, None]], axis=1)


def unpack_blocks(blk: np.ndarray):
    return blk[:, :6], blk[:, 6:9], blk[:, 9]


def allreduce_blocks_(blk, group=None):
    """In-place sum over ranks of a block buffer (torch tensor, CPU/gloo or CUDA/nccl)."""
    dist.all_reduce(blk, op=dist.ReduceOp.SUM, group=group)
    return blk


class _DevicePtr:
    """Zero-copy view of `count` fp64 values at a raw device address, via __cuda_array_interface__."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def device_view_f64(ptr: int, count: int, device):
    return torch.as_tensor(_DevicePtr(ptr, count), device=device)


def make_nccl_allreduce(device, group=None):
    """Callback for BAProblem.set_allreduce: sums the library's device buffer over ranks in place with
    torch.distributed (NCCL).  The collective is enqueued on torch's current stream, the same stream the
    library context was created on, so it is ordered between the evaluation and the decision kernels."""

    def _cb(ptr: int, count: int):
        t = device_view_f64(ptr, count, device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    return _cb
