"""Voice sharding across ranks and the single per-update collective.

Voices are independent until they `+=` into the device mix buffers, and the
post-process (HRTF decoder / B-Format decode) is linear, so every rank mixes its own
voices all the way to RealOut and ONE sum-reduce of the [real_channels][1024] block
per update combines them.  This module holds the rank arithmetic and the collective
so the same code runs under NCCL (bench.py, GPUs) and gloo (CPU tests)."""
from __future__ import annotations


def shard_range(total_voices: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition: (first, count) of the voices owned by `rank`."""
    base, rem = divmod(total_voices, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def owner_of(voice: int, total_voices: int, world: int) -> int:
    base, rem = divmod(total_voices, world)
    cut = rem * (base + 1)
    if voice < cut:
        return voice // (base + 1)
    return rem + (voice - cut) // max(base, 1)


def reduce_real_out(block, dst: int = 0):
    """Sum-reduces one rank-local RealOut block (a torch tensor, CUDA for NCCL or CPU
    for gloo) onto rank `dst`; in place.  A no-op for world size 1."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(block, dst=dst, op=dist.ReduceOp.SUM)
    return block


def slot_owner(slot: int, world: int) -> int:
    """Effect slots are independent of each other: slot s runs on rank s mod G."""
    return slot % world


def allreduce_wet(wet):
    """Sums the slots' Wet buffers of all ranks in place (torch tensor viewing the device's
    wet storage between b200mix_render_begin and b200mix_render_end): effects consume the
    summed send input.  Under NCCL call it with the mixer's stream current."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(wet, op=dist.ReduceOp.SUM)
    return wet
