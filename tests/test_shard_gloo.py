"""The N>1 path on CPU: world_size-2 gloo.  Each rank mixes its shard of the voices
(with the CPU oracle standing in for the device mixer — the host-side dealing of voices and
slots is what is under test; gloo stands in for the library's own exchange, which is tested
on hardware in test_gpu_shard.py) and the reduced RealOut must equal the single-process mix of
all voices.  The handle exchange of pyb200mix.shard.connect runs here against a recording
stand-in for the library."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyb200mix import shard

HERE = os.path.dirname(os.path.abspath(__file__))


def reduce_real_out(block, dst=0):
    """gloo stand-in for the library's RealOut reduce (b200mix_render on a sharded set)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(block, dst=dst, op=dist.ReduceOp.SUM)
    return block


def allreduce_wet(wet):
    """gloo stand-in for the library's wet exchange: every owner sees the summed send input."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(wet, op=dist.ReduceOp.SUM)
    return wet


def test_shard_ranges_partition_the_voices():
    for total, world in [(4096, 8), (10, 4), (3, 8), (65536, 8), (1, 1)]:
        seen = []
        for r in range(world):
            first, count = shard.shard_range(total, world, r)
            seen += list(range(first, first + count))
            for v in range(first, first + count):
                assert shard.owner_of(v, total, world) == r
        assert seen == list(range(total))


def _mix(voices, total, updates):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "openal-soft_b200"))
    from helpers import mixlib, synth
    from helpers.mixlib import MixDevice
    from pyb200mix import abi, scene
    rng = np.random.default_rng(11)
    desc = synth.hrtf_desc(total, 64)
    params, coeffs, dry = synth.voice_set(rng, total, 64)
    dev = MixDevice(mixlib.oracle(), desc)
    dev.set_hrtf_decoder(*synth.decoder(np.random.default_rng(7)))
    for i in voices:
        dev.buffer_data(i, abi.FMT_I16, scene.voice_buffer_fast(i, 6000))
    for p in params:
        p.loop_end = 6000
        p.position %= 3000
    dev.voices_update([params[i] for i in voices], coeffs[voices], dry[voices], None)
    out = np.stack([dev.render() for _ in range(updates)])
    dev.close()
    return out


def _worker(rank, world, port, total, updates, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard.shard_range(total, world, rank)
    out = _mix(list(range(first, first + count)), total, updates)
    block = torch.from_numpy(out.copy())
    reduce_real_out(block, dst=0)
    if rank == 0:
        ret.put(block.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_reduce_equals_single_process_mix():
    total, updates, world = 24, 3, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, updates, ret)) for r in range(world)]
    for p in procs:
        p.start()
    reduced = ret.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    full = _mix(list(range(total)), total, updates)
    err = np.abs(reduced.astype(np.float64) - full)
    assert np.abs(full).max() > 1e-3
    scale = max(1.0, float(np.abs(full).max()))   # fp32 re-association of the two partial sums
    assert err.max() <= 5e-6 * scale and np.sqrt((err ** 2).mean()) <= 5e-7 * scale


# ---- effect slots: wet all-reduce between the two halves of the update --------------------
def _mix_slots(voices, total, updates, rank, world):
    """Voices `voices` on this rank; 2 reverb slots, each installed only on its owner rank;
    the wet buffers are summed across ranks between render_begin and render_end."""
    import ctypes as C
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "openal-soft_b200"))
    from helpers import golden, mixlib, synth
    from helpers.mixlib import MixDevice
    from pyb200mix import abi, scene
    rng = np.random.default_rng(21)
    desc = synth.hrtf_desc(total, 64)
    desc.num_sends = 1
    desc.wet_channels = 4
    desc.max_slots = 2
    params, coeffs, dry = synth.voice_set(rng, total, 64)
    send = (rng.standard_normal((total, 1, 4)) * 0.3).astype(np.float32)
    for k, p in enumerate(params):
        p.send_slot[0] = k % 2
        p.loop_end = 6000
        p.position %= 3000
    fxs = [golden.load("hrtf_bsinc24_reverb_v6"), golden.load("hrtf_spline_reverb_dens0_mod_v4")]
    dev = MixDevice(mixlib.oracle(), desc)
    dev.set_hrtf_decoder(*synth.decoder(np.random.default_rng(7)))
    for i in voices:
        dev.buffer_data(i, abi.FMT_I16, scene.voice_buffer_fast(i, 6000))
    for s_, fx in enumerate(fxs):
        if shard.slot_owner(s_, world) == rank:
            dev.slot_reverb(s_, abi.reverb_params_from(fx["reverb_params"].tobytes()),
                            fx["reverb_gains"])
    dev.voices_update([params[i] for i in voices], coeffs[voices], dry[voices], send[voices])
    outs = []
    for _ in range(updates):
        ptr, cnt = dev.render_begin()
        wet = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(cnt,))
        allreduce_wet(torch.from_numpy(wet))        # in place on the oracle's storage
        outs.append(dev.render_end())
    dev.close()
    return np.stack(outs)


def _slot_worker(rank, world, port, total, updates, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard.shard_range(total, world, rank)
    out = _mix_slots(list(range(first, first + count)), total, updates, rank, world)
    block = torch.from_numpy(out.copy())
    reduce_real_out(block, dst=0)
    if rank == 0:
        ret.put(block.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_slot_ownership_equals_single_process_mix():
    total, updates, world = 16, 4, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_slot_worker, args=(r, world, port, total, updates, ret))
             for r in range(world)]
    for p in procs:
        p.start()
    reduced = ret.get(timeout=150)
    for p in procs:
        p.join(timeout=30)
    single = _mix_slots(list(range(total)), total, updates, 0, 1)
    scale = float(np.abs(single).max())
    assert scale > 1e-3
    assert np.abs(reduced - single).max() <= 5e-6 * max(scale, 1.0)


# ---- pyb200mix.shard.connect: the start-up exchange, against a recording stand-in -----------
class _FakeLib:
    """Records the b200mix_shard_* calls connect() makes (no GPU here)."""

    def __init__(self, rank):
        self.rank, self.calls = rank, []

    def b200mix_shard_init(self, dev, rank, world, buf):
        buf.raw = bytes([0x40 + rank]) * 64
        self.calls.append(("init", rank, world))
        return 0

    def b200mix_shard_connect(self, dev, blob):
        self.calls.append(("connect", bytes(blob)))
        return 0

    def b200mix_shard_nccl_id(self, buf):
        buf.raw = bytes(range(128))
        self.calls.append(("nccl_id",))
        return 0

    def b200mix_shard_nccl(self, dev, rank, world, ident):
        self.calls.append(("nccl", rank, world, bytes(ident)))
        return 0


def _connect_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _FakeLib(rank)
    shard.connect(lib, None, rank, world, "p2p")
    shard.connect(lib, None, rank, world, "nccl")
    ret.put((rank, lib.calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_connect_gathers_handles_in_rank_order_and_broadcasts_the_nccl_id():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_connect_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(ret.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    want_blob = bytes([0x40]) * 64 + bytes([0x41]) * 64
    for r in range(world):
        calls = got[r]
        assert calls[0] == ("init", r, world)
        assert calls[1] == ("connect", want_blob)          # every rank sees the rank-ordered table
        assert calls[-1] == ("nccl", r, world, bytes(range(128)))
    assert ("nccl_id",) in got[0] and ("nccl_id",) not in got[1]


def test_slot_ownership_rule():
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            mine = shard.owned_slots(13, world, r)
            assert all(shard.slot_owner(s, world) == r for s in mine)
            seen += mine
        assert sorted(seen) == list(range(13))
