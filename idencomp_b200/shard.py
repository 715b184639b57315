"""Sharding of the block-independent hot path over ranks (one process per GPU), SURVEY.md section 8e.

Blocks of an .idn container are independent (model state resets per block, compressor_block.rs:61-62), so the path
shards with NO data-path collective: every rank compresses a contiguous range of blocks, the host side exchanges the
per-rank container sizes (an all_gather of one integer), takes an exclusive prefix sum and concatenates the bodies
in rank order.  torch.distributed is plumbing only (gloo on CPU in the tests, NCCL or gloo next to the GPU path).
"""
from __future__ import annotations

import numpy as np


def block_ranges(n_blocks: int, world: int):
    """Contiguous, balanced block ranges: rank r gets [lo, hi).  Earlier ranks take the remainder."""
    base, rem = divmod(n_blocks, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def exclusive_offsets(sizes):
    """Offsets of the per-rank bodies in the final container (exclusive prefix sum) and the total."""
    off = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(np.asarray(sizes, dtype=np.int64), out=off[1:])
    return off[:-1].tolist(), int(off[-1])


def gather_bodies(body: bytes, dist, dst: int = 0):
    """all_gather the body sizes, then gather the bodies on `dst` in rank order.  Returns (offsets, joined bytes) on
    dst and (offsets, None) elsewhere.  `dist` is an initialised torch.distributed module (any backend that moves CPU
    tensors, e.g. gloo)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([len(body)], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, mine)
    sizes = [int(s.item()) for s in sizes]
    offsets, total = exclusive_offsets(sizes)
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.uint8)
    if body:
        buf[:len(body)] = torch.frombuffer(bytearray(body), dtype=torch.uint8)
    if rank == dst:
        parts = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)]
        dist.gather(buf, parts, dst=dst)
        joined = bytearray(total)
        for r in range(world):
            joined[offsets[r]:offsets[r] + sizes[r]] = parts[r][:sizes[r]].numpy().tobytes()
        return offsets, bytes(joined)
    dist.gather(buf, None, dst=dst)
    return offsets, None


def compress_sharded(compress_range, preamble: bytes, n_blocks: int, dist, dst: int = 0):
    """File-level compress over all ranks.  `compress_range(lo, hi)` returns the container bytes (block headers +
    slices, no preamble, no terminator) of blocks [lo, hi); `preamble` is header + metadata (identical on every rank:
    the model subset is chosen once, on the first block, before sharding).  Returns the .idn bytes on `dst`."""
    lo, hi = block_ranges(n_blocks, dist.get_world_size())[dist.get_rank()]
    body = compress_range(lo, hi) if hi > lo else b""
    _, joined = gather_bodies(body, dist, dst)
    if joined is None:
        return None
    return preamble + joined + b"\x00" * 8  # empty terminator block (idn/compressor.rs:579)
