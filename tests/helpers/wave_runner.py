#!/usr/bin/env python3
"""Plays a seeded scene through ONE OpenAL library's own PLAYBACK backend — the reference's Wave File
Writer (alc/backends/wave.cpp), a real backend with its mixer thread — for a little over half a
second and leaves the .wav.  The device is paused while the sources are set up and started, so
everything begins on the first update after alcDeviceResumeSOFT: the file's content after its
leading silence is deterministic.  tests/test_seam_cpu.py compares the stock reference with the
patched library (ALSOFT_B200MIX=1): backends call DeviceBase::renderSamples, which is where the seam
sits, so a playback device mixes on the GPU exactly like a loopback device.

usage: wave_runner.py <libopenal path> <out.wav>"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "openal-soft_b200"))
from pyb200mix import scene  # noqa: E402

lib, wav = sys.argv[1], sys.argv[2]
conf = wav + ".conf"
open(conf, "w").write(f"[general]\ndrivers=wave\nfrequency=48000\nperiod_size=1024\nperiods=2\nchannels=stereo\nsample-type=float32\nstereo-encoding=hrtf\n[wave]\nfile={wav}\n")
os.environ["ALSOFT_CONF"] = conf
os.environ.setdefault("ALSOFT_LOGLEVEL", "1")
al = C.CDLL(lib, mode=C.RTLD_GLOBAL)
al.alcOpenDevice.restype = C.c_void_p; al.alcOpenDevice.argtypes=[C.c_char_p]
al.alcCreateContext.restype = C.c_void_p; al.alcCreateContext.argtypes=[C.c_void_p, C.c_void_p]
al.alcMakeContextCurrent.argtypes=[C.c_void_p]; al.alcDestroyContext.argtypes=[C.c_void_p]; al.alcCloseDevice.argtypes=[C.c_void_p]
al.alcDevicePauseSOFT.argtypes=[C.c_void_p]; al.alcDeviceResumeSOFT.argtypes=[C.c_void_p]
al.alGenBuffers.argtypes=[C.c_int, C.POINTER(C.c_uint)]; al.alGenSources.argtypes=[C.c_int, C.POINTER(C.c_uint)]
al.alBufferData.argtypes=[C.c_uint, C.c_int, C.c_void_p, C.c_int, C.c_int]
al.alSourcei.argtypes=[C.c_uint, C.c_int, C.c_int]; al.alSourcef.argtypes=[C.c_uint, C.c_int, C.c_float]
al.alSource3f.argtypes=[C.c_uint, C.c_int, C.c_float, C.c_float, C.c_float]; al.alSourcePlayv.argtypes=[C.c_int, C.POINTER(C.c_uint)]
dev = al.alcOpenDevice(None); assert dev
ctx = al.alcCreateContext(dev, None); assert ctx
al.alcMakeContextCurrent(ctx)
al.alcDevicePauseSOFT(dev)
V=16; keep=[]; src=(C.c_uint*V)()
for i in range(V):
    b,s=C.c_uint(0),C.c_uint(0)
    pcm=np.ascontiguousarray(scene.voice_buffer_fast(i, scene.BUFFER_FRAMES)); keep.append(pcm)
    al.alGenBuffers(1,C.byref(b)); al.alBufferData(b,0x1101,pcm.ctypes.data,pcm.nbytes,48000)
    al.alGenSources(1,C.byref(s)); al.alSourcei(s,0x1009,b.value); al.alSourcei(s,0x1007,1)
    al.alSourcef(s,0x1003,scene.voice_pitch(i)); al.alSourcef(s,0x100A,scene.voice_gain(V))
    al.alSource3f(s,0x1004,*[float(x) for x in scene.voice_position(i)]); src[i]=s.value
al.alSourcePlayv(V,src)
al.alcDeviceResumeSOFT(dev)
time.sleep(0.6)
al.alcDevicePauseSOFT(dev)
al.alcMakeContextCurrent(None); al.alcDestroyContext(ctx); al.alcCloseDevice(dev)
