"""Voice sharding across ranks — launcher-side plumbing only.

The exchange itself lives in the library (include/b200mix.h, "voice-sharded device sets"):
b200mix_render reduce-scatters the slots' Wet buffers and reduces RealOut onto rank 0 on the
device's own stream.  What is left for the host is (1) dealing voices and slots over the ranks
and (2) carrying 64-byte CUDA IPC handles (or the 128-byte NCCL id) between the processes
once at start-up.  A C++ host does (2) with whatever it has (MPI, a socket); the Python
launchers here (bench.py, tools/) use a torch.distributed group."""
from __future__ import annotations

import ctypes as C

HANDLE_BYTES = 64       # B200MIX_SHARD_HANDLE_BYTES
NCCL_ID_BYTES = 128     # B200MIX_NCCL_ID_BYTES


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


def slot_owner(slot: int, world: int) -> int:
    """The library's ownership rule: slot s is processed by rank s mod world, which is the
    rank that installs its effect and receives the summed send input."""
    return slot % world


def owned_slots(num_slots: int, world: int, rank: int) -> list[int]:
    return list(range(rank, num_slots, world))


def connect(lib, dev, rank: int, world: int, transport: str = "p2p", group=None) -> None:
    """Joins device `dev` (a b200mix_device*) of ctypes library `lib` to the sharded set.
    transport "p2p": b200mix_shard_init -> all-gather of the IPC handles over `group` ->
    b200mix_shard_connect.  transport "nccl": rank 0's b200mix_shard_nccl_id is broadcast,
    then b200mix_shard_nccl.  `group` is a torch.distributed group whose backend can move
    Python objects from host memory (gloo); None = the default group."""
    import torch.distributed as dist

    def ck(rc, what):
        if rc != 0:
            lib.b200mix_last_error.restype = C.c_char_p
            lib.b200mix_last_error.argtypes = [C.c_void_p]
            raise RuntimeError(f"{what} failed ({rc}): {(lib.b200mix_last_error(dev) or b'').decode()}")

    if transport == "p2p":
        buf = C.create_string_buffer(HANDLE_BYTES)
        ck(lib.b200mix_shard_init(dev, rank, world, buf), "b200mix_shard_init")
        blobs = [None] * world
        dist.all_gather_object(blobs, buf.raw, group=group)
        assert all(len(b) == HANDLE_BYTES for b in blobs)
        ck(lib.b200mix_shard_connect(dev, b"".join(blobs)), "b200mix_shard_connect")
    elif transport == "nccl":
        box = [None]
        if rank == 0:
            buf = C.create_string_buffer(NCCL_ID_BYTES)
            ck(lib.b200mix_shard_nccl_id(buf), "b200mix_shard_nccl_id")
            box[0] = buf.raw
        dist.broadcast_object_list(box, src=0, group=group)
        ck(lib.b200mix_shard_nccl(dev, rank, world, box[0]), "b200mix_shard_nccl")
    else:
        raise ValueError(transport)
