"""Multi-GPU decomposition: one process per GPU, launched by torchrun.

Both paths shard by independent units with no data exchange on the hot path:
  LZ4  : contiguous ranges of blocks          (the reference's thread-per-block, P-LZ4:724-749)
  JPEG : contiguous ranges of 8x8 group rows  (the reference's thread-per-group, P-JPG:1297-1302)
Each rank encodes its units into its own device buffer; the only collective is ONE all-gather of the
per-rank compressed byte totals, whose exclusive scan gives every rank its base offset in the global
stream (what the reference's serial write_output, LZ4.c:427-441, does implicitly).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    first: int   # first unit owned by this rank
    count: int   # number of units owned


def shard_units(n_units: int, rank: int, world: int) -> Shard:
    """Rank g owns units [g*ceil(n/G), min(n, (g+1)*ceil(n/G)))."""
    if world < 1 or not (0 <= rank < world) or n_units < 0:
        raise ValueError("bad shard arguments")
    per = (n_units + world - 1) // world if n_units else 0
    first = min(n_units, rank * per)
    last = min(n_units, first + per)
    return Shard(rank, world, first, last - first)


def lz4_shard_bytes(n: int, block_len: int, rank: int, world: int):
    """Byte range of the input owned by `rank`, and its Shard of blocks."""
    nblocks = (n + block_len - 1) // block_len
    sh = shard_units(nblocks, rank, world)
    lo = sh.first * block_len
    hi = min(n, (sh.first + sh.count) * block_len)
    return lo, hi, sh


def jpeg_shard_groups(w: int, h: int, rank: int, world: int) -> Shard:
    """Shard of 8x8 groups by whole group rows (any group-row boundary is a valid cut: groups are independent)."""
    bpr = (w + 7) // 8
    total = (w * h + 63) // 64  # the reference processes ceil(w*h/64) groups, JPEG.c:1131
    rows = (total + bpr - 1) // bpr
    sh = shard_units(rows, rank, world)
    first = min(total, sh.first * bpr)
    last = min(total, (sh.first + sh.count) * bpr)
    return Shard(rank, world, first, last - first)


def exclusive_bases(totals):
    """Exclusive scan of per-rank totals -> per-rank base offsets, and the grand total."""
    bases, run = [], 0
    for t in totals:
        bases.append(run)
        run += int(t)
    return bases, run


def gather_totals(local_total: int, device=None, group=None):
    """The single collective of the path: all-gather of one int64 per rank (NCCL on GPU, gloo on CPU).

    Returns (bases list, grand_total).  With no process group initialised it degenerates to one rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [0], int(local_total)
    world = dist.get_world_size(group)
    mine = torch.tensor([int(local_total)], dtype=torch.int64, device=device if device is not None else "cpu")
    allv = torch.empty(world, dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    return exclusive_bases(allv.tolist())


def gather_totals_device(d_total, group=None):
    """Device-side form of gather_totals for the timed path: `d_total` is a 1-element int64 CUDA tensor written by
    the encode kernel; returns the world-size int64 tensor of all ranks' totals without a host round trip.
    Issued on torch's current stream (callers make that the context stream so it is ordered after the kernel)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    allv = torch.empty(world, dtype=torch.int64, device=d_total.device)
    dist.all_gather_into_tensor(allv, d_total.contiguous(), group=group)
    return allv
