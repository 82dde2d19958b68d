#!/usr/bin/env python3
"""Secondary measurement (not the driver's bench line): BASELINE config 3 shape —
V HRTF voices (bsinc24) each sending to one of S aux slots carrying an EAX reverb
(parameter block = the reference's, from tests/golden) or a convolution (IR taps).
Prints per-update device time and the per-kernel split from the library's counters.
usage: bench_effects.py [--voices 16384] [--slots 32] [--effect reverb|conv] [--taps 96000]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "openal-soft_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from pyb200mix import abi, scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voices", type=int, default=16384)
    ap.add_argument("--slots", type=int, default=32)
    ap.add_argument("--effect", default="reverb")
    ap.add_argument("--taps", type=int, default=96000)
    ap.add_argument("--steps", type=int, default=16)
    args = ap.parse_args()
    import torch
    lib = bench.load_product()
    lib.b200mix_slot_reverb.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(abi.ReverbParams)]
    lib.b200mix_slot_convolution.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.b200mix_slot_output_gains.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    nv, ns = args.voices, args.slots
    desc = abi.DeviceDesc()
    desc.struct_size = C.sizeof(abi.DeviceDesc)
    desc.cuda_device = 0
    desc.sample_rate = 48000
    desc.dry_channels, desc.real_channels, desc.ir_size = 4, 2, 64
    desc.post_process = abi.POST_HRTF
    desc.real_left, desc.real_right = 0, 1
    desc.max_voices = desc.max_buffers = nv
    desc.num_sends, desc.wet_channels, desc.max_slots = 1, 4, ns
    h = C.c_void_p()
    assert lib.b200mix_create(C.byref(desc), C.byref(h)) == 0, lib.b200mix_last_error(None)
    rng = np.random.default_rng(7)
    dec = (rng.standard_normal((4, 91, 2)) * 0.05).astype(np.float32)
    hf = np.array([2.0, 1.1547005, 1.1547005, 1.1547005], dtype=np.float32)
    sc = np.full(4, -0.9123257, dtype=np.float32)
    lib.b200mix_set_hrtf_decoder(h, 4, 91, dec.ctypes.data, hf.ctypes.data, sc.ctypes.data)
    if args.effect == "reverb":
        fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "hrtf_bsinc24_reverb_v6.npz")))
        rp = abi.reverb_params_from(fx["reverb_params"].tobytes())
        rp.struct_size = C.sizeof(abi.ReverbParams)
        g = np.ascontiguousarray(fx["reverb_gains"], dtype=np.float32)
        for s in range(ns):
            assert lib.b200mix_slot_reverb(h, s, C.byref(rp)) == 0, lib.b200mix_last_error(h)
            lib.b200mix_slot_output_gains(h, s, 8, g.ctypes.data)
    else:
        for s in range(ns):
            ir = (np.random.default_rng(0xC0FFEE ^ s).standard_normal((1, args.taps))
                  * np.exp(-np.arange(args.taps) / (args.taps / 6.0)) * 0.02).astype(np.float32)
            g = np.array([[0.5, 0.0, 0.0, 0.8]], dtype=np.float32)
            assert lib.b200mix_slot_convolution(h, s, 1, args.taps, ir.ctypes.data) == 0, lib.b200mix_last_error(h)
            lib.b200mix_slot_output_gains(h, s, 1, g.ctypes.data)
    hrtf = bench.load_hrtf(lib)
    for k in range(nv):
        pcm = scene.voice_buffer_fast(k)
        lib.b200mix_buffer_data(h, k, abi.FMT_I16, 1, pcm.shape[0], pcm.ctypes.data, pcm.nbytes)
    params, coeffs, pitches = bench.synth_voices(range(nv), nv, lib, hrtf)
    send = np.zeros((nv, 1, 4), dtype=np.float32)
    send[:, 0, :] = np.array([0.5, 0.2, -0.1, 0.3], dtype=np.float32) * scene.voice_gain(nv)
    for k in range(nv):
        params[k].send_slot[0] = k % ns
    assert lib.b200mix_voices_update(h, nv, params, coeffs.ctypes.data, None, send.ctypes.data) == 0, \
        lib.b200mix_last_error(h)
    stream = torch.cuda.ExternalStream(lib.b200mix_stream(h))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = C.c_void_p()
    for _ in range(4):
        assert lib.b200mix_render_device(h, 1024, C.byref(out)) == 0, lib.b200mix_last_error(h)
    torch.cuda.synchronize()
    lib.b200mix_profile(h, 1)
    ms, mix = [], []
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        lib.b200mix_render_device(h, 1024, C.byref(out))
        e1.record(stream)
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
        mix.append(lib.b200mix_last_mix_kernel_ms(h))
    print(json.dumps({"voices": nv, "slots": ns, "effect": args.effect,
                      "taps": args.taps if args.effect != "reverb" else None,
                      "ms_per_update": float(np.mean(ms)), "mix_kernel_ms": float(np.mean(mix)),
                      "effects_and_post_ms": float(np.mean(ms) - np.mean(mix)),
                      "rt_voices": nv * (1000.0 * 1024 / 48000) / float(np.mean(ms))}))
    lib.b200mix_destroy(h)


if __name__ == "__main__":
    main()
