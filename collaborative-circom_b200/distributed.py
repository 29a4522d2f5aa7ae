"""The one collective of the multi-GPU path: an all-gather of every rank's partial MSM sums, once per proof.

One process per GPU (torch.distributed; NCCL over NVLink on the GPU box, gloo in the CPU tests).  The reference has no
counterpart -- each party runs the whole MSM on its own cores (rayon); the partition here is by index range of the bases,
`cohost_msm_shard_range`, so every rank holds the partial sums of the same MSM sites and the fold is a plain group sum.
EC addition is not an NCCL reduction op, hence gather + local fold instead of all-reduce.
"""
from __future__ import annotations

import ctypes

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """(offset, length) of the slice of an n-term MSM accumulated by `rank` -- the C++ host layer's own rule."""
    from .prover import load_host
    off, ln = ctypes.c_size_t(), ctypes.c_size_t()
    load_host().cohost_msm_shard_range(n, rank, world, ctypes.byref(off), ctypes.byref(ln))
    return off.value, ln.value


def make_all_gather(world: int, device=None):
    """Returns f(partials: np.ndarray[uint64]) -> np.ndarray[uint64] of world * len(partials), rank order.

    device: a torch device for the staging tensors (cuda:<local rank> with NCCL; None = CPU tensors for gloo)."""
    if world == 1:
        return lambda partials: partials
    import torch
    import torch.distributed as dist
    bufs = {}

    def all_gather(partials: np.ndarray) -> np.ndarray:
        n = partials.size
        if n not in bufs:
            bufs[n] = (torch.empty(n, dtype=torch.int64, device=device), torch.empty(world * n, dtype=torch.int64, device=device))
        src, dst = bufs[n]
        src.copy_(torch.from_numpy(np.ascontiguousarray(partials).view(np.int64)))
        dist.all_gather_into_tensor(dst, src)
        return dst.cpu().numpy().view(np.uint64)

    return all_gather


class _DevBuf:
    """A raw device range as a __cuda_array_interface__ object, so torch can view it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def make_p2p(device):
    """The cross-GPU leg of the REP3 network in block mode (host/network.hpp DeviceBridge): returns comm(ops) for Rep3Session, which
    issues the listed sends / receives between GPUs as ONE grouped NCCL call (batch_isend_irecv: sends and receives of a round cannot
    deadlock) and returns when the data has arrived.  ops: (dir, peer_rank, device_ptr, nbytes), dir 0 = send, 1 = receive."""
    import torch
    import torch.distributed as dist

    def comm(ops):
        with torch.cuda.device(device):
            reqs = []
            keep = []
            for d, peer, ptr, nbytes in ops:
                t = torch.as_tensor(_DevBuf(ptr, nbytes), device=device)
                keep.append(t)
                reqs.append(dist.P2POp(dist.isend if d == 0 else dist.irecv, t, peer))
            for w in dist.batch_isend_irecv(reqs):
                w.wait()
            torch.cuda.current_stream().synchronize()

    return comm
