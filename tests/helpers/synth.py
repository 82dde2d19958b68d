"""Synthetic post-ALU voice descriptors (no reference needed): random decaying HRIRs,
delays, gains and steps.  Used for oracle-vs-CUDA parity at sizes where the
reference's own parameter stage is not available (GPU box) or too slow."""
import ctypes as C

import numpy as np

from pyb200mix import abi, scene


def hrtf_desc(max_voices, ir=64, dry_channels=4):
    d = abi.DeviceDesc()
    d.struct_size = C.sizeof(abi.DeviceDesc)
    d.cuda_device = -1
    d.sample_rate = 48000
    d.dry_channels = dry_channels
    d.real_channels = 2
    d.wet_channels = 0
    d.num_sends = 0
    d.ir_size = ir
    d.post_process = abi.POST_HRTF
    d.real_left = 0
    d.real_right = 1
    d.max_voices = max_voices
    d.max_buffers = max_voices
    d.max_slots = 0
    return d


def stereo_desc(max_voices, dry_channels=3):
    d = hrtf_desc(max_voices, ir=0, dry_channels=dry_channels)
    d.post_process = abi.POST_AMBIDEC
    return d


def decoder(rng, channels=4, ir=91):
    coeffs = (rng.standard_normal((channels, ir, 2)) * np.exp(-np.arange(ir) / 12.0)[None, :, None]
              * 0.2).astype(np.float32)
    hf = np.array([2.0] + [1.1547005] * (channels - 1), dtype=np.float32)[:channels]
    sc = np.full(channels, -0.9123257, dtype=np.float32)
    return coeffs, hf, sc


def voice_set(rng, n, ir, resampler=abi.RS_BSINC24, hrtf=True, dry_channels=4, looping=True,
              frames=scene.BUFFER_FRAMES, pitch_lo=0.5, pitch_hi=2.0):
    """Returns (params list, coeffs [n][ir][2], dry [n][cd])."""
    params = []
    coeffs = (rng.standard_normal((n, max(ir, 1), 2)) * np.exp(-np.arange(max(ir, 1)) / 10.0)[None, :, None]
              ).astype(np.float32)
    dry = (rng.standard_normal((n, dry_channels)) * 0.3).astype(np.float32)
    g = scene.voice_gain(n)
    for i in range(n):
        p = abi.VoiceParams()
        p.voice = i
        p.flags = abi.VF_PLAYING | abi.VF_STATIC | abi.VF_RESET
        if looping:
            p.flags |= abi.VF_LOOPING
        if hrtf:
            p.flags |= abi.VF_HRTF
        p.buffer = i
        p.resampler = resampler if not isinstance(resampler, (list, tuple)) else resampler[i % len(resampler)]
        p.position = int(rng.integers(0, frames // 2))
        p.position_frac = int(rng.integers(0, 65536))
        p.loop_start = 0
        p.loop_end = frames
        pitch = 1.0 if i % 16 == 15 else float(rng.uniform(pitch_lo, pitch_hi))
        if i % 16 == 15:
            p.position_frac = 0
        p.step = max(1, min(int(pitch * 65536.0), 10 << 16))
        p.hrtf_delay[0] = int(rng.integers(0, 64))
        p.hrtf_delay[1] = int(rng.integers(0, 64))
        p.hrtf_gain = g * float(rng.uniform(0.5, 1.0))
        for s in range(abi.MAX_SENDS):
            p.send_slot[s] = abi.NO_SLOT
        params.append(p)
    return params, coeffs, dry
