"""Builds the same synthetic scene on the reference (through its public AL API) and
on a b200mix-surface implementation (oracle or CUDA product), feeding the latter
with the reference's own post-ALU voice parameters (the drop-in seam)."""
import ctypes as C
import numpy as np
from . import refal
from .mixlib import MixDevice
from pyb200mix import abi, scene


def make_ref_scene(num_voices, hrtf, resampler, attrs=None, pitch_fn=scene.voice_pitch,
                   looping=True, frames=scene.BUFFER_FRAMES, max_sources=None, fmt="i16"):
    a = {refal.ALC_HRTF_SOFT: 1 if hrtf else 0,
         refal.ALC_MONO_SOURCES: max_sources or max(num_voices, 1)}
    if attrs:
        a.update(attrs)
    ref = refal.RefDevice(a)
    ref.out_channels = ref.desc.real_channels
    pcms = []
    for i in range(num_voices):
        pcm = scene.voice_buffer_fmt(i, frames, fmt)
        pcms.append(pcm)
        ref.add_voice(pcm, scene.BUFFER_RATE, pitch_fn(i), scene.voice_position(i),
                      scene.voice_gain(num_voices), resampler, looping=looping,
                      fmt=scene.FORMATS[fmt][1])
    return ref, pcms


def mirror_device(mixlib, ref, max_voices, pcms, fmt=abi.FMT_I16):
    """Creates the b200mix-surface device that mirrors a reference device."""
    desc = abi.DeviceDesc()
    C.memmove(C.byref(desc), C.byref(ref.desc), C.sizeof(desc))
    desc.max_voices = max_voices
    desc.max_buffers = max(len(pcms), 1)
    nslots, wet = ref.slot_info()
    desc.max_slots = nslots
    desc.wet_channels = wet[0] if nslots else 0
    dev = MixDevice(mixlib, desc)
    if desc.post_process == abi.POST_HRTF:
        dev.set_hrtf_decoder(*ref.hrtf_decoder())
    elif desc.post_process == abi.POST_AMBIDEC:
        dev.set_ambi_decoder(*ref.ambi_decoder())
    for i, pcm in enumerate(pcms):
        dev.buffer_data(i, fmt, pcm)
    return dev


def feed_params(dev, ref, first, nv):
    """Copies the reference's current post-ALU voice targets into dev.
    Voice slot k of the reference plays source k (sources are started in order)."""
    wet = dev.desc.wet_channels
    n, params, coeffs, dry, send, state = ref.snapshot(wet_channels=wet)
    plist = []
    for k in range(nv):
        p = params[k]
        q = abi.VoiceParams()
        C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
        q.buffer = k
        if first:
            q.flags |= abi.VF_RESET
            q.position = 0
            q.position_frac = 0
        plist.append(q)
    dev.voices_update(plist, coeffs[:nv], dry[:nv], send[:nv] if wet else None)
    return state


def err_stats(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.sqrt((d ** 2).mean())), float(np.abs(d).max())
