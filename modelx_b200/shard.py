"""Chunk-range sharding of one blob across ranks (one process per GPU).

The hash path needs no collective: rank r digests a contiguous range of chunks.  The only exchange
is an all-gather of the 32-byte chunk digests; this module holds the index
arithmetic both bench.py and the tests use, plus the gather itself over torch.distributed.
"""
from __future__ import annotations

from typing import Tuple


def chunk_count(size: int, chunk: int) -> int:
    return max(1, -(-size // chunk))


def chunks_per_rank(nchunks: int, world: int) -> int:
    """Every rank is given the same slot count so the gather is one fixed-size all_gather; ranks at
    the tail may own fewer (or zero) real chunks."""
    return -(-nchunks // world)


def chunk_range(rank: int, world: int, nchunks: int) -> Tuple[int, int]:
    per = chunks_per_rank(nchunks, world)
    return min(rank * per, nchunks), min((rank + 1) * per, nchunks)


def byte_range(rank: int, world: int, size: int, chunk: int) -> Tuple[int, int]:
    c0, c1 = chunk_range(rank, world, chunk_count(size, chunk))
    return min(c0 * chunk, size), min(c1 * chunk, size)


def gather_chunk_digests(local, world: int, nchunks: int, group=None):
    """all_gather fixed-size per-rank digest buffers (uint8 tensor of chunks_per_rank*32 bytes, zero
    padded) and return the first nchunks*32 bytes: because ranges are contiguous and only the tail
    is padded, that prefix is exactly the blob's chunk-digest list.  Works on CUDA tensors (NCCL)
    and CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist

    per = chunks_per_rank(nchunks, world)
    assert local.numel() == per * 32 and local.dtype == torch.uint8
    if world == 1:
        return local[:nchunks * 32]
    out = torch.empty(world * per * 32, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out[:nchunks * 32]
