#!/usr/bin/env python3
"""Drives ONE OpenAL library (a path given on the command line) through the public AL/ALC API
only — loopback device, buffers, sources, alcRenderSamplesSOFT — on a seeded scene, and writes
the rendered updates to an .npz.  tests/test_gpu_dropin.py runs it once on the stock compiled
reference and once on the patched libopenal_b200.so (ALSOFT_B200MIX=1): the application code is
identical, only the mixer behind alcRenderSamplesSOFT differs.

usage: al_runner.py <libopenal path> <out.npz> <voices> <updates> <hrtf 0|1> [resampler] [fx]

fx = "none" (default) | "reverb" (one EAX reverb slot, every source sends to it) | "mix" (EAX
reverb + echo + an equalizer slot that feeds the reverb slot via AL_EFFECTSLOT_TARGET_SOFT;
properties, slot gain and an effect type change while playing) | "filt" (direct low-pass /
band-pass filters that change and detach while playing) | "mixfilt" (both, plus send filters)
| "stream" (alSourceQueueBuffers: queues that run out, looping queues, buffers queued and
unqueued while playing) | "stereo" (AL_FORMAT_STEREO16 sources next to mono ones) | "conv"
(two convolution slots: a mono float32 impulse response at 44.1 kHz — resampled by the library — and
a stereo 16-bit one at the device rate; slot gain changes while playing) | "reset" (reverb scene;
alcResetDeviceSOFT toggles HRTF while the sources play) | "hoa" (second- / third-order B-Format beds,
AL_SOFT_bformat_hoa, ACN or FuMa, on the first-order device) | "hoadev2" / "hoadev3" (an ALC_BFORMAT3D_SOFT
device of order 2 / 3; beds of that order and the next, rotated by the application: AmbiRotator) |
"bformat" (first-order B-Format
sources, AL_FORMAT_BFORMAT3D_16, whose orientation the application turns) | "rebuf" (a buffer is
deleted and another one of the same size created — usually at the same address — and played) |
"misc" (pause / resume, seeking a playing source, pitch and gain changes, a moving listener,
looping switched off while playing) | "misc2" (a second context on the same device, sources that
share one buffer, a send that moves to another slot and is removed, deferred updates through
alcSuspendContext / alcProcessContext, all sources stopped and others started on the freed voices)
| "misc3" (streaming sources paused, resumed and sought; a queue that underruns, is refilled and
played again; a stereo and a B-Format source with a filtered reverb send; the slot's effect set to
null and back) | "allfx" (one slot per remaining EFX effect — vocal morpher, frequency shifter,
autowah, distortion, compressor, ring modulator, flanger — with property changes while playing) |
"pshift" (two pitch-shifter slots, up and down, re-tuned while playing) | "i16" (16-bit output: the host's limiter, dither and Write<i16> run
on the block the mixer delivered; the .npz then holds the samples scaled to +-1) | "quad", "x51",
"mono", "uhj", "uhj512", "tsme", "stab51", "bs2b" (other outputs: quad / 5.1 / mono speakers, UHJ-
encoded stereo with the IIR or the 512-tap FIR encoder, TSME, 5.1 with the front stabilizer, stereo
with BS2B crossfeed — the last four through the reference's own configuration file) | "ragged"
(reverb scene rendered in updates of 1024, 100, 7, 640, 1, 333 … frames) | "formats" (one source per
buffer storage format: 8-bit, 16-bit, 32-bit integer, float32, double, mu-law, A-law, IMA4 and
MS-ADPCM mono and stereo, and quad / 5.1 16-bit sources) | "fuzzN" (the "mixfilt" scene with
streaming sources, driven by a seeded random sequence of API calls — play / stop / pause / rewind,
seeks, pitch / gain / position / looping changes, filters and sends attached and removed, queues
unqueued and refilled, slot gains and effect properties, sources deleted and created; from N = 100
also: a slot's effect replaced by another type, slot targets, deferred updates, resampler changes,
stereo and B-Format sources created along the way, streams fed new buffers, buffers swapped on
stopped sources; N = 200..299 also rendered in ragged update
sizes; N = 300..399 with a slot turned into a convolution reverb and alcResetDeviceSOFT toggling HRTF; from N = 400 the
extended set on a scene without filters until the sequence attaches one) | "ctx" (300 sources: a second context created while the first plays, the first
one's voice array growing past 256, the second context destroyed while its sources play) | "short" (sources of 40 ... 1500 frames that end
inside the update they start in, restarted every other update) | "direct" (a stereo source
with AL_DIRECT_CHANNELS_SOFT: not wired into the seam — the device must disconnect, not crash)"""
import ctypes as C
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "openal-soft_b200"))
from pyb200mix import scene  # noqa: E402

ALC_FREQUENCY, ALC_MONO_SOURCES = 0x1007, 0x1010
ALC_FORMAT_CHANNELS_SOFT, ALC_FORMAT_TYPE_SOFT = 0x1990, 0x1991
ALC_STEREO_SOFT, ALC_FLOAT_SOFT, ALC_HRTF_SOFT, ALC_SHORT_SOFT = 0x1501, 0x1406, 0x1992, 0x1402
ALC_MONO_SOFT, ALC_QUAD_SOFT, ALC_5POINT1_SOFT = 0x1500, 0x1503, 0x1504
ALC_OUTPUT_MODE_SOFT, ALC_STEREO_UHJ_SOFT = 0x19AC, 0x19AF
AL_BUFFER, AL_LOOPING, AL_PITCH, AL_GAIN, AL_POSITION = 0x1009, 0x1007, 0x1003, 0x100A, 0x1004
AL_SOURCE_STATE, AL_PLAYING, AL_STOPPED, AL_SAMPLE_OFFSET = 0x1010, 0x1012, 0x1014, 0x1025
AL_FORMAT_MONO16, AL_SOURCE_RESAMPLER_SOFT = 0x1101, 0x1212
AL_FORMAT_STEREO16, AL_BUFFERS_PROCESSED, AL_BUFFERS_QUEUED = 0x1103, 0x1016, 0x1015
AL_FORMAT_BFORMAT3D_16, AL_ORIENTATION = 0x20032, 0x100F
AL_VELOCITY = 0x1006
AL_UNPACK_AMBISONIC_ORDER_SOFT, AL_AMBISONIC_LAYOUT_SOFT, AL_AMBISONIC_SCALING_SOFT = 0x199D, 0x1997, 0x1998
AL_ACN_SOFT, AL_SN3D_SOFT, AL_N3D_SOFT = 1, 1, 2
AL_DIRECT_CHANNELS_SOFT, ALC_CONNECTED = 0x1033, 0x313
AL_EFFECT_NULL = 0x0000
AL_EFFECT_DISTORTION, AL_EFFECT_FLANGER, AL_EFFECT_FREQUENCY_SHIFTER, AL_EFFECT_VOCAL_MORPHER = 0x0003, 0x0005, 0x0006, 0x0007
AL_EFFECT_RING_MODULATOR, AL_EFFECT_AUTOWAH, AL_EFFECT_COMPRESSOR = 0x0009, 0x000A, 0x000B
AL_AUXILIARY_SEND_FILTER, AL_FILTER_NULL = 0x20006, 0
AL_EFFECT_TYPE, AL_EFFECT_EAXREVERB, AL_EFFECT_ECHO, AL_EFFECT_EQUALIZER, AL_EFFECT_CHORUS = 0x8001, 0x8000, 0x0004, 0x000C, 0x0001
AL_EFFECTSLOT_EFFECT, AL_EFFECTSLOT_GAIN, AL_EFFECTSLOT_TARGET_SOFT = 0x0001, 0x0002, 0x199C
AL_EAXREVERB_DECAY_TIME, AL_EAXREVERB_REFLECTIONS_GAIN = 0x0006, 0x0009
AL_ECHO_DELAY, AL_ECHO_FEEDBACK = 0x0001, 0x0004
AL_EQUALIZER_LOW_GAIN, AL_EQUALIZER_MID1_GAIN = 0x0001, 0x0003
AL_CHORUS_RATE = 0x0003
AL_EFFECT_CONVOLUTION_SOFT, AL_FORMAT_MONO_FLOAT32 = 0xA000, 0x10010
AL_DIRECT_FILTER, AL_FILTER_TYPE, AL_FILTER_LOWPASS, AL_FILTER_BANDPASS = 0x20005, 0x8001, 0x0001, 0x0003
AL_LOWPASS_GAIN, AL_LOWPASS_GAINHF, AL_BANDPASS_GAIN, AL_BANDPASS_GAINLF, AL_BANDPASS_GAINHF = 1, 2, 1, 2, 3


def format_buffer(i, frames):
    """voice i's buffer in the i-th storage format (core/fmt_traits.h; al/buffer.cpp:640-720)."""
    k = i % 13
    pcm = scene.voice_buffer_fast(i, frames)
    if k == 0:
        return np.ascontiguousarray(scene.voice_buffer_fmt(i, frames, "u8")), 0x1100
    if k == 1:
        return np.ascontiguousarray(pcm), AL_FORMAT_MONO16
    if k == 2:
        return np.ascontiguousarray(pcm.astype(np.float32) / np.float32(32768.0)), 0x10010
    if k == 3:
        return np.ascontiguousarray(pcm.astype(np.float64) / 32768.0 * 0.9), 0x10012
    if k == 4:
        return np.ascontiguousarray(scene.voice_buffer_fmt(i, frames, "mulaw")), 0x10014
    if k == 5:
        return np.ascontiguousarray(scene.voice_buffer_fmt(i, frames, "alaw")), 0x10016
    if k == 6:
        return np.ascontiguousarray(scene.adpcm_blocks(i, "ima4", min(frames, 6500) // 65)), 0x1300
    if k == 7:
        return np.ascontiguousarray(scene.adpcm_blocks(i, "msadpcm", frames // 64)), 0x1302
    if k == 8:
        return np.ascontiguousarray(pcm.astype(np.int32) * 65536 + 12345), 0x19DB
    rng = np.random.default_rng(0xF0 + i)
    if k == 9:
        # stereo IMA4: per channel a 4-byte header (predictor, step index), then 4-byte groups per channel
        blocks = min(frames, 6500) // 65
        out = rng.choice(np.array([0x00, 0x11, 0x19, 0x91, 0x08, 0x80, 0x21, 0x12], dtype=np.uint8), size=(blocks, 72))
        hdr = scene.voice_buffer_fast(i, blocks * 2).astype(np.int64) & 0xFFFF
        for c in range(2):
            out[:, 4 * c] = hdr[c::2] & 0xFF
            out[:, 4 * c + 1] = hdr[c::2] >> 8
            out[:, 4 * c + 2] = 20 + c
            out[:, 4 * c + 3] = 0
        return np.ascontiguousarray(out.reshape(-1)), 0x1301
    if k == 10:
        # stereo MS-ADPCM: predictor index, scale and two history samples per channel, then nibbles
        blocks = frames // 64
        out = rng.choice(np.array([0x00, 0x11, 0x1F, 0xF1, 0x0F, 0xF0, 0x21, 0xEF], dtype=np.uint8), size=(blocks, 76))
        hist = scene.voice_buffer_fast(i, blocks * 4).astype(np.int64) & 0xFFFF
        for c in range(2):
            out[:, c] = (i + c) % 7
            out[:, 2 + 2 * c], out[:, 3 + 2 * c] = 40 + c, 0
            out[:, 6 + 2 * c], out[:, 7 + 2 * c] = hist[c::4] & 0xFF, hist[c::4] >> 8
            out[:, 10 + 2 * c], out[:, 11 + 2 * c] = hist[2 + c::4] & 0xFF, hist[2 + c::4] >> 8
        return np.ascontiguousarray(out.reshape(-1)), 0x1303
    nch, fmt = (4, 0x1205) if k == 11 else (6, 0x120B)
    chans = [scene.voice_buffer_fast(i + c, frames) for c in range(nch)]
    return np.ascontiguousarray(np.stack(chans, axis=1).reshape(-1)), fmt


class _NoCalls:
    """stands in for the library when one fuzz action is masked out (AL_RUNNER_FUZZ_SKIP="update:action,...")"""
    def __getattr__(self, name):
        return lambda *a: 0


FUZZ_EXT = {"on": False, "suspended": None, "extra": [], "keep": [], "retarget": False, "family3": False,
            "dev": None, "attrs": None, "hrtf": 0, "irbuf": 0}


def fuzz_actions(real_al, rng, sources, V, slots, streams, filters, bufids, u=0, ctx=None):
    """A handful of random API calls between two updates (errors an application could provoke —
    a seek beyond the end, looping a playing queue — are part of the sequence; both libraries see
    the same calls)."""
    stream_ids = {i for i, _ in streams}
    skip = os.environ.get("AL_RUNNER_FUZZ_SKIP", "").split(",")
    ext = FUZZ_EXT["on"]
    if FUZZ_EXT["suspended"]:
        real_al.alcProcessContext(ctx)                       # the batch deferred since the last update becomes visible
        FUZZ_EXT["suspended"] = None
    if ext and not FUZZ_EXT["extra"]:
        # other source formats for the sources created along the way: stereo and first-order B-Format
        for nch, fmt in ((2, AL_FORMAT_STEREO16), (4, AL_FORMAT_BFORMAT3D_16), (2, AL_FORMAT_STEREO16)):
            chans = [scene.voice_buffer_fast(40 + nch + c + len(FUZZ_EXT["extra"]), 9000) for c in range(nch)]
            pcm = np.ascontiguousarray(np.stack(chans, axis=1).reshape(-1))
            FUZZ_EXT["keep"].append(pcm)
            b = C.c_uint(0)
            real_al.alGenBuffers(1, C.byref(b))
            real_al.alBufferData(b, fmt, pcm.ctypes.data, pcm.nbytes, 48000)
            FUZZ_EXT["extra"].append(b.value)
    for k in range(int(rng.integers(3, 9))):
        al = _NoCalls() if f"{u}:{k}" in skip else real_al
        only = os.environ.get("AL_RUNNER_FUZZ_ONLY")
        if only and f"{u}:{k}" not in only.split(","):
            al = _NoCalls()
        i = int(rng.integers(0, V))
        s = sources[i]
        op = int(rng.integers(0, (26 if FUZZ_EXT["family3"] else 24) if ext else 16))
        if os.environ.get("AL_RUNNER_FUZZ_LOG"):
            st = C.c_int(0)
            real_al.alGetSourcei(s, AL_SOURCE_STATE, C.byref(st))
            print(f"fuzz: source {i} (state {st.value:#x}, {'stream' if i in stream_ids else 'static'}) op {op}", file=sys.stderr)
        if op == 0:
            al.alSourcePlay(s)
        elif op == 1:
            al.alSourceStop(s)
        elif op == 2:
            al.alSourcePause(s)
        elif op == 3:
            al.alSourceRewind(s)
        elif op == 4:
            al.alSourcef(s, AL_PITCH, float(rng.uniform(0.3, 3.0)))
        elif op == 5:
            al.alSourcef(s, AL_GAIN, float(rng.uniform(0.0, 0.3)))
        elif op == 6:
            al.alSource3f(s, AL_POSITION, *[float(x) for x in rng.uniform(-3.0, 3.0, 3)])
        elif op == 7:
            al.alSourcei(s, AL_SAMPLE_OFFSET, int(rng.integers(0, 7000)))
        elif op == 8:
            al.alSourcei(s, AL_LOOPING, int(rng.integers(0, 2)))
        elif op == 9:
            al.alSourcei(s, AL_DIRECT_FILTER, int(rng.choice([AL_FILTER_NULL, filters[0], filters[1]])))
        elif op == 10:
            k = int(rng.integers(0, len(slots) + 1))
            al.alSource3i(s, AL_AUXILIARY_SEND_FILTER, slots[k][0] if k < len(slots) else 0, int(rng.integers(0, 2)),
                          int(rng.choice([AL_FILTER_NULL, filters[0]])))
        elif op == 11 and i in stream_ids:
            done = C.c_int(0)
            al.alGetSourcei(s, AL_BUFFERS_PROCESSED, C.byref(done))
            if done.value > 0:
                n = int(rng.integers(1, done.value + 1))
                got = (C.c_uint * n)()
                al.alSourceUnqueueBuffers(s, n, got)
                if rng.integers(0, 2):
                    al.alSourceQueueBuffers(s, n, got)
        elif op == 12:
            k = int(rng.integers(0, len(slots)))
            al.alAuxiliaryEffectSlotf(slots[k][0], AL_EFFECTSLOT_GAIN, float(rng.uniform(0.1, 1.0)))
        elif op == 13:
            al.alEffectf(slots[0][1], AL_EAXREVERB_DECAY_TIME, float(rng.uniform(0.4, 4.0)))
            al.alEffectf(slots[0][1], AL_EAXREVERB_REFLECTIONS_GAIN, float(rng.uniform(0.0, 1.0)))
            al.alAuxiliaryEffectSloti(slots[0][0], AL_EFFECTSLOT_EFFECT, slots[0][1])
        elif op == 14:
            al.alListener3f(AL_POSITION, *[float(x) for x in rng.uniform(-1.0, 1.0, 3)])
            al.alListenerf(AL_GAIN, float(rng.uniform(0.5, 1.0)))
        elif op == 15 and i not in stream_ids and bufids[i]:
            # the source goes away; a new one takes its place and its buffer
            old = C.c_uint(s)
            al.alDeleteSources(1, C.byref(old))
            new = C.c_uint(0)
            al.alGenSources(1, C.byref(new))
            al.alSourcei(new, AL_BUFFER, bufids[i])
            al.alSourcei(new, AL_LOOPING, int(rng.integers(0, 2)))
            al.alSourcef(new, AL_GAIN, 0.15)
            al.alSource3f(new, AL_POSITION, *[float(x) for x in rng.uniform(-2.0, 2.0, 3)])
            if os.environ.get("DBG_PITCH"):  # (debug knobs used while bisecting)
                al.alSourcef(new, AL_PITCH, float(os.environ["DBG_PITCH"]))
            if os.environ.get("DBG_RS"):
                al.alSourcei(new, AL_SOURCE_RESAMPLER_SOFT, int(os.environ["DBG_RS"]))
            sources[i] = new.value
            al.alSourcePlay(new)
        elif op == 16:
            # slot 1's effect becomes another one (a new EffectState), or none
            et = int(rng.choice([AL_EFFECT_ECHO, AL_EFFECT_CHORUS, AL_EFFECT_NULL, AL_EFFECT_RING_MODULATOR, AL_EFFECT_EQUALIZER,
                                 AL_EFFECT_DISTORTION, AL_EFFECT_COMPRESSOR, AL_EFFECT_FLANGER, AL_EFFECT_AUTOWAH]))
            if FUZZ_EXT["family3"] and et in (AL_EFFECT_CHORUS, AL_EFFECT_FLANGER, AL_EFFECT_RING_MODULATOR, AL_EFFECT_COMPRESSOR,
                                              AL_EFFECT_ECHO):
                # deviceUpdate leaves these effects' oscillator phase / envelope / feedback filter history
                # alone (e.g. ChorusState::mLfoOffset, alc/effects/chorus.cpp:136-168; EchoState::mFilter,
                # alc/effects/echo.cpp:76-89), so in the reference they run on across a device reset while a
                # re-created mixer starts them clean: DESIGN.md "known divergences"
                et = AL_EFFECT_EQUALIZER
            if os.environ.get("AL_RUNNER_FUZZ_LOG"):
                print(f"fuzz:   effect type {et:#x}", file=sys.stderr)
            al.alEffecti(slots[1][1], AL_EFFECT_TYPE, et)
            al.alAuxiliaryEffectSloti(slots[1][0], AL_EFFECTSLOT_EFFECT, slots[1][1])
        elif op == 17:
            # (only where the Dry mix and the slots' Wet mixes map their channels alike — HRTF output.
            # Elsewhere a per-channel effect that moves to a mix with another channel map while it plays
            # carries its scalar mCurrentGain along in the reference; library and oracle keep gains per
            # output channel and fade the moved channels: DESIGN.md "known divergences")
            tgt = int(rng.choice([0, slots[0][0]]))
            if FUZZ_EXT["retarget"]:
                al.alAuxiliaryEffectSloti(slots[2][0], AL_EFFECTSLOT_TARGET_SOFT, tgt)
        elif op == 18 and ctx is not None:
            if al is real_al:
                real_al.alcSuspendContext(ctx)
                FUZZ_EXT["suspended"] = True
        elif op == 19:
            al.alSourcei(s, AL_SOURCE_RESAMPLER_SOFT, int(rng.integers(0, 8)))
        elif op == 20 and i not in stream_ids and bufids[i]:
            # the source goes away; a stereo or B-Format one takes its place
            old = C.c_uint(s)
            al.alDeleteSources(1, C.byref(old))
            new = C.c_uint(0)
            al.alGenSources(1, C.byref(new))
            al.alSourcei(new, AL_BUFFER, FUZZ_EXT["extra"][int(rng.integers(0, len(FUZZ_EXT["extra"])))])
            al.alSourcei(new, AL_LOOPING, int(rng.integers(0, 2)))
            al.alSourcef(new, AL_GAIN, 0.1)
            al.alSourcef(new, AL_PITCH, float(rng.uniform(0.5, 2.0)))
            al.alSource3i(new, AL_AUXILIARY_SEND_FILTER, slots[0][0], 0, AL_FILTER_NULL)
            if al is real_al:
                sources[i] = new.value
            al.alSourcePlay(new)
        elif op == 21 and i in stream_ids:
            # the application feeds the stream one more buffer
            nb = C.c_uint(0)
            al.alGenBuffers(1, C.byref(nb))
            part = np.ascontiguousarray(scene.voice_buffer_fast(60 + i + u, int(rng.integers(800, 3000))))
            FUZZ_EXT["keep"].append(part)
            al.alBufferData(nb, AL_FORMAT_MONO16, part.ctypes.data, part.nbytes, 48000)
            al.alSourceQueueBuffers(s, 1, C.byref(nb))
        elif op == 22 and i not in stream_ids:
            # a stopped source takes another source's buffer (AL_INVALID_OPERATION while it plays)
            al.alSourcei(s, AL_BUFFER, bufids[int(rng.integers(0, V))] or bufids[1])
        elif op == 23:
            al.alSourcei(s, 0x202, int(rng.integers(0, 2)))                 # AL_SOURCE_RELATIVE
            al.alSourcef(s, 0x1021, float(rng.uniform(0.0, 2.0)))           # AL_ROLLOFF_FACTOR
        elif op == 24:
            # slot 1 becomes a convolution reverb (the buffer goes onto the slot before the effect does)
            if not FUZZ_EXT["irbuf"] and al is real_al:
                g = np.random.default_rng(0x1B)
                ir = (g.standard_normal(1800) * np.exp(-np.arange(1800) / 400.0) * 0.2).astype(np.float32)
                FUZZ_EXT["keep"].append(ir)
                b = C.c_uint(0)
                real_al.alGenBuffers(1, C.byref(b))
                real_al.alBufferData(b, AL_FORMAT_MONO_FLOAT32, ir.ctypes.data, ir.nbytes, 44100)
                FUZZ_EXT["irbuf"] = b.value
            al.alEffecti(slots[1][1], AL_EFFECT_TYPE, AL_EFFECT_CONVOLUTION_SOFT)
            al.alAuxiliaryEffectSloti(slots[1][0], AL_BUFFER, FUZZ_EXT["irbuf"])
            al.alAuxiliaryEffectSloti(slots[1][0], AL_EFFECTSLOT_EFFECT, slots[1][1])
        elif op == 25 and int(rng.integers(0, 3)) == 0:
            # the application switches the output mode (HRTF <-> plain stereo) with everything in flight
            # (not while updates are deferred: the reference would then mix voices whose parameters
            # still describe the previous configuration — nothing to compare against)
            if al is real_al and not FUZZ_EXT["suspended"]:
                FUZZ_EXT["hrtf"] ^= 1
                attrs2 = list(FUZZ_EXT["attrs"])
                attrs2[attrs2.index(ALC_HRTF_SOFT) + 1] = FUZZ_EXT["hrtf"]
                assert real_al.alcResetDeviceSOFT(FUZZ_EXT["dev"], (C.c_int * len(attrs2))(*attrs2))
                FUZZ_EXT["retarget"] = bool(FUZZ_EXT["hrtf"])
        al.alGetError()


def main():
    lib, out_path, V, U, hrtf = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    resampler = int(sys.argv[6]) if len(sys.argv) > 6 else 7          # bsinc24
    fx = sys.argv[7] if len(sys.argv) > 7 else "none"
    conf = os.path.join(os.path.dirname(out_path), f"alsoft_{os.getpid()}.conf")
    conf_lines = {"tsme": "[general]\nstereo-encoding=tsme\n", "stab51": "[general]\nfront-stablizer=true\n",
                  "bs2b": "[general]\ncf_level=4\n", "uhj512": "[general]\n[uhj]\nencode-filter=fir512\n"}
    open(conf, "w").write(conf_lines.get(fx, "[general]\n"))
    os.environ["ALSOFT_CONF"] = conf
    os.environ.setdefault("ALSOFT_LOGLEVEL", "1")
    al = C.CDLL(lib, mode=C.RTLD_GLOBAL)
    al.alcLoopbackOpenDeviceSOFT.restype = C.c_void_p
    al.alcLoopbackOpenDeviceSOFT.argtypes = [C.c_char_p]
    al.alcCreateContext.restype = C.c_void_p
    al.alcCreateContext.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    al.alcMakeContextCurrent.argtypes = [C.c_void_p]
    al.alcDestroyContext.argtypes = [C.c_void_p]
    al.alcCloseDevice.argtypes = [C.c_void_p]
    al.alcSuspendContext.argtypes = [C.c_void_p]
    al.alcProcessContext.argtypes = [C.c_void_p]
    al.alSourceStopv.argtypes = [C.c_int, C.POINTER(C.c_uint)]
    al.alcRenderSamplesSOFT.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    al.alcResetDeviceSOFT.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    al.alcGetIntegerv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    al.alGenBuffers.argtypes = [C.c_int, C.POINTER(C.c_uint)]
    al.alDeleteBuffers.argtypes = [C.c_int, C.POINTER(C.c_uint)]
    al.alGenSources.argtypes = [C.c_int, C.POINTER(C.c_uint)]
    al.alDeleteSources.argtypes = [C.c_int, C.POINTER(C.c_uint)]
    al.alBufferData.argtypes = [C.c_uint, C.c_int, C.c_void_p, C.c_int, C.c_int]
    al.alSourcei.argtypes = [C.c_uint, C.c_int, C.c_int]
    al.alSourcef.argtypes = [C.c_uint, C.c_int, C.c_float]
    al.alSource3f.argtypes = [C.c_uint, C.c_int, C.c_float, C.c_float, C.c_float]
    al.alSourcefv.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_float)]
    al.alSourcePlayv.argtypes = [C.c_int, C.POINTER(C.c_uint)]
    al.alSourceStop.argtypes = [C.c_uint]
    al.alSourcePlay.argtypes = [C.c_uint]
    al.alSourcePause.argtypes = [C.c_uint]
    al.alSourceRewind.argtypes = [C.c_uint]
    al.alListener3f.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
    al.alListenerf.argtypes = [C.c_int, C.c_float]
    al.alGetSourcei.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_int)]
    al.alSourceQueueBuffers.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_uint)]
    al.alSourceUnqueueBuffers.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_uint)]
    al.alSource3i.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int]
    al.alGenEffects.argtypes = [C.c_int, C.POINTER(C.c_uint)]
    al.alEffecti.argtypes = [C.c_uint, C.c_int, C.c_int]
    al.alEffectf.argtypes = [C.c_uint, C.c_int, C.c_float]
    al.alGenAuxiliaryEffectSlots.argtypes = [C.c_int, C.POINTER(C.c_uint)]
    al.alAuxiliaryEffectSloti.argtypes = [C.c_uint, C.c_int, C.c_int]
    al.alAuxiliaryEffectSlotf.argtypes = [C.c_uint, C.c_int, C.c_float]
    al.alGenFilters.argtypes = [C.c_int, C.POINTER(C.c_uint)]
    al.alFilteri.argtypes = [C.c_uint, C.c_int, C.c_int]
    al.alFilterf.argtypes = [C.c_uint, C.c_int, C.c_float]
    dev = al.alcLoopbackOpenDeviceSOFT(None)
    assert dev
    out16 = fx == "i16"
    layout, nout = {"quad": (ALC_QUAD_SOFT, 4), "x51": (ALC_5POINT1_SOFT, 6), "stab51": (ALC_5POINT1_SOFT, 6),
                    "mono": (ALC_MONO_SOFT, 1), "hoadev2": (0x1507, 9), "hoadev3": (0x1507, 16)}.get(fx, (ALC_STEREO_SOFT, 2))
    attrs = [ALC_FORMAT_CHANNELS_SOFT, layout, ALC_FORMAT_TYPE_SOFT, ALC_SHORT_SOFT if out16 else ALC_FLOAT_SOFT, ALC_FREQUENCY, 48000,
             ALC_MONO_SOURCES, max(V, 1), ALC_HRTF_SOFT, hrtf] + ([ALC_OUTPUT_MODE_SOFT, ALC_STEREO_UHJ_SOFT] if fx in ("uhj", "uhj512") else []) + [0]
    if fx.startswith("hoadev"):
        # ALC_BFORMAT3D_SOFT output (RealOut is the Dry mix) of order 2 / 3: ACN, order 2 in N3D, order 3 in SN3D
        attrs = attrs[:-1] + [0x1997, 1, 0x1998, 2 if fx == "hoadev2" else 1, 0x1999, int(fx[-1]), 0]
    ctx = al.alcCreateContext(dev, (C.c_int * len(attrs))(*attrs))
    assert ctx
    al.alcMakeContextCurrent(ctx)
    keep, sources = [], (C.c_uint * V)()

    def make_slot(al_type, gain=1.0, fprops=None, iprops=None):
        e, sl = C.c_uint(0), C.c_uint(0)
        al.alGenEffects(1, C.byref(e))
        al.alEffecti(e, AL_EFFECT_TYPE, al_type)
        for k, v in (fprops or {}).items():
            al.alEffectf(e, k, float(v))
        for k, v in (iprops or {}).items():
            al.alEffecti(e, k, int(v))
        al.alGenAuxiliaryEffectSlots(1, C.byref(sl))
        al.alAuxiliaryEffectSlotf(sl, AL_EFFECTSLOT_GAIN, gain)
        al.alAuxiliaryEffectSloti(sl, AL_EFFECTSLOT_EFFECT, e.value)
        return sl.value, e.value

    slots, streams, bufids = [], [], []
    reset = fx == "reset"
    ragged = fx == "ragged"
    fuzz = None
    if fx.startswith("fuzz"):
        fuzz = np.random.default_rng(0xF22 + int(fx[4:] or 0))
        FUZZ_EXT["on"] = int(fx[4:] or 0) >= 100          # seeds from 100: the extended set of calls
        FUZZ_EXT["retarget"] = bool(hrtf)
        ragged = 200 <= int(fx[4:] or 0) < 300             # seeds 200..299: ... rendered in ragged update sizes
        FUZZ_EXT["family3"] = 300 <= int(fx[4:] or 0) < 400   # seeds 300..399: ... plus convolution slots and device resets
        FUZZ_EXT.update(dev=dev, attrs=attrs, hrtf=hrtf)
        # seeds from 400: the extended set on a scene that starts WITHOUT filters — the first one is attached by the sequence
        fx = "mix" if int(fx[4:] or 0) >= 400 else "mixfilt"
    if reset or (ragged and fuzz is None):
        fx = "reverb"
    filt = fx in ("filt", "mixfilt")
    sendfilter = C.c_uint(0)
    if fx == "misc3":
        al.alGenFilters(1, C.byref(sendfilter))
        al.alFilteri(sendfilter, AL_FILTER_TYPE, AL_FILTER_LOWPASS)
        al.alFilterf(sendfilter, AL_LOWPASS_GAIN, 0.8)
        al.alFilterf(sendfilter, AL_LOWPASS_GAINHF, 0.4)
    if fx == "mixfilt":
        fx = "mix"
    lowpass, bandpass = C.c_uint(0), C.c_uint(0)
    if filt or fuzz is not None:
        al.alGenFilters(1, C.byref(lowpass))
        al.alFilteri(lowpass, AL_FILTER_TYPE, AL_FILTER_LOWPASS)
        al.alFilterf(lowpass, AL_LOWPASS_GAIN, 0.9)
        al.alFilterf(lowpass, AL_LOWPASS_GAINHF, 0.25)
        al.alGenFilters(1, C.byref(bandpass))
        al.alFilteri(bandpass, AL_FILTER_TYPE, AL_FILTER_BANDPASS)
        al.alFilterf(bandpass, AL_BANDPASS_GAIN, 0.8)
        al.alFilterf(bandpass, AL_BANDPASS_GAINLF, 0.3)
        al.alFilterf(bandpass, AL_BANDPASS_GAINHF, 0.5)
    if fx == "conv":
        def conv_slot(ir, fmt, rate, gain):
            b, e, sl = C.c_uint(0), C.c_uint(0), C.c_uint(0)
            ir = np.ascontiguousarray(ir)
            keep.append(ir)
            al.alGenBuffers(1, C.byref(b))
            al.alBufferData(b, fmt, ir.ctypes.data, ir.nbytes, rate)
            al.alGenEffects(1, C.byref(e))
            al.alEffecti(e, AL_EFFECT_TYPE, AL_EFFECT_CONVOLUTION_SOFT)
            al.alGenAuxiliaryEffectSlots(1, C.byref(sl))
            al.alAuxiliaryEffectSloti(sl, AL_BUFFER, b.value)
            al.alAuxiliaryEffectSlotf(sl, AL_EFFECTSLOT_GAIN, gain)
            al.alAuxiliaryEffectSloti(sl, AL_EFFECTSLOT_EFFECT, e.value)
            return sl.value, e.value
        rng = np.random.default_rng(0xC0FFEE)
        t = np.arange(2200)
        mono = (rng.standard_normal(2200) * np.exp(-t / 500.0) * 0.2).astype(np.float32)
        t2 = np.arange(1500)
        st = (rng.standard_normal((1500, 2)) * np.exp(-t2 / 300.0)[:, None] * 0.2 * 32767).astype(np.int16)
        slots.append(conv_slot(mono, AL_FORMAT_MONO_FLOAT32, 44100, 0.8))
        slots.append(conv_slot(st, AL_FORMAT_STEREO16, 48000, 0.6))
    if fx == "allfx":
        slots.append(make_slot(AL_EFFECT_VOCAL_MORPHER, 0.9, {0x0006: 2.0}))                       # rate
        slots.append(make_slot(AL_EFFECT_FREQUENCY_SHIFTER, 0.8, {0x0001: 220.0}))                 # frequency
        slots.append(make_slot(AL_EFFECT_AUTOWAH, 0.7, {0x0003: 100.0, 0x0004: 50.0}))             # resonance, peak gain
        slots.append(make_slot(AL_EFFECT_DISTORTION, 0.6, {0x0001: 0.3, 0x0002: 0.2}))             # edge, gain
        slots.append(make_slot(AL_EFFECT_COMPRESSOR, 0.9))
        slots.append(make_slot(AL_EFFECT_RING_MODULATOR, 0.8, {0x0001: 300.0}))                    # frequency
        slots.append(make_slot(AL_EFFECT_FLANGER, 0.8, {0x0003: 0.4, 0x0005: -0.6}))               # rate, feedback
    if fx == "pshift":
        slots.append(make_slot(0x0008, 0.9))                                                       # AL_EFFECT_PITCH_SHIFTER: an octave up
        slots.append(make_slot(0x0008, 0.8, None, {0x0001: -7, 0x0002: 20}))                       # a fifth down + 20 cents
    if fx in ("reverb", "mix", "misc3"):
        slots.append(make_slot(AL_EFFECT_EAXREVERB, 0.9))
    if fx == "mix":
        slots.append(make_slot(AL_EFFECT_ECHO, 0.7, {AL_ECHO_DELAY: 0.031, AL_ECHO_FEEDBACK: 0.4}))
        slots.append(make_slot(AL_EFFECT_EQUALIZER, 0.8, {AL_EQUALIZER_LOW_GAIN: 0.5, AL_EQUALIZER_MID1_GAIN: 2.0}))
        al.alAuxiliaryEffectSloti(slots[2][0], AL_EFFECTSLOT_TARGET_SOFT, slots[0][0])
    if fuzz is not None and FUZZ_EXT["family3"]:
        # (see fuzz_actions, op 16: no echo in sequences with device resets)
        al.alEffecti(slots[1][1], AL_EFFECT_TYPE, AL_EFFECT_EQUALIZER)
        al.alAuxiliaryEffectSloti(slots[1][0], AL_EFFECTSLOT_EFFECT, slots[1][1])
    for i in range(V):
        b, s = C.c_uint(0), C.c_uint(0)
        # every fourth voice is a short one-shot (runs out, fades, stops by itself)
        oneshot = i % 4 == 3
        pcm = np.ascontiguousarray(scene.voice_buffer_fast(i, 3000 + 37 * i if oneshot else scene.BUFFER_FRAMES))
        if fx == "short":
            # every source ends inside the update it starts in (or the next one): 40 ... 1500 frames, no loop
            oneshot = True
            pcm = np.ascontiguousarray(scene.voice_buffer_fast(i, 40 + 61 * i))
        fmt = AL_FORMAT_MONO16
        if fx == "misc3" and i == 20:
            other = scene.voice_buffer_fast(i + 1, len(pcm))
            pcm = np.ascontiguousarray(np.stack([pcm, other], axis=1).reshape(-1))
            fmt = AL_FORMAT_STEREO16
        if fx == "misc3" and i == 22:
            chans = [pcm] + [scene.voice_buffer_fast(i + k, len(pcm)) for k in (1, 2, 3)]
            pcm = np.ascontiguousarray(np.stack(chans, axis=1).reshape(-1))
            fmt = AL_FORMAT_BFORMAT3D_16
        if (fx == "stereo" and i % 2 == 0) or (fx == "direct" and i == 0):
            # a stereo buffer: left = this voice's waveform, right = the next one's (interleaved)
            other = scene.voice_buffer_fast(i + 1, len(pcm))
            pcm = np.ascontiguousarray(np.stack([pcm, other], axis=1).reshape(-1))
            fmt = AL_FORMAT_STEREO16
        if fx == "bformat" and i % 2 == 0:
            # W X Y Z = four different waveforms
            chans = [pcm] + [scene.voice_buffer_fast(i + k, len(pcm)) for k in (1, 2, 3)]
            pcm = np.ascontiguousarray(np.stack(chans, axis=1).reshape(-1))
            fmt = AL_FORMAT_BFORMAT3D_16
        hoa_order = 0
        if fx in ("hoa", "hoadev2", "hoadev3") and i % 2 == 0:
            # AL_SOFT_bformat_hoa: second- and third-order beds (9 / 16 channels), ACN or FuMa layout
            # (on an ambisonic device: beds of the device's order and of the next one — lower-order
            # beds would be up-sampled, VoiceFlag::IsAmbisonic)
            hoa_order = (2 + (i // 2) % 2) if fx == "hoa" else (int(fx[-1]) + (i // 2) % 2)
            nchan = (hoa_order + 1) ** 2
            chans = [pcm] + [scene.voice_buffer_fast(i + k, len(pcm)) for k in range(1, nchan)]
            pcm = np.ascontiguousarray(np.stack(chans, axis=1).reshape(-1))
            fmt = AL_FORMAT_BFORMAT3D_16
        if fx == "formats":
            pcm, fmt = format_buffer(i, len(pcm))
        keep.append(pcm)
        al.alGenSources(1, C.byref(s))
        if (fx == "stream" and i % 3 != 2) or (fx == "misc3" and i < 6) or (fuzz is not None and i % 4 == 0):
            # a streaming source: three queued buffers of different lengths; every third source loops its queue
            qb = (C.c_uint * 3)()
            al.alGenBuffers(3, qb)
            off = 0
            for k, ln in enumerate((2000 + 13 * i, 1500, 3100)):
                part = np.ascontiguousarray(pcm[off:off + ln])
                keep.append(part)
                al.alBufferData(qb[k], fmt, part.ctypes.data, part.nbytes, 48000)
                off += ln
            al.alSourceQueueBuffers(s, 3, qb)
            al.alSourcei(s, AL_LOOPING, 1 if i % 3 == 1 else 0)
            streams.append((i, qb))
        else:
            al.alGenBuffers(1, C.byref(b))
            if hoa_order:
                al.alBufferi(b, AL_UNPACK_AMBISONIC_ORDER_SOFT, hoa_order)
                if i % 4 == 0 or hoa_order > 3:            # FuMa stops at third order
                    al.alBufferi(b, AL_AMBISONIC_LAYOUT_SOFT, AL_ACN_SOFT)
                    al.alBufferi(b, AL_AMBISONIC_SCALING_SOFT, AL_N3D_SOFT if i % 8 == 0 else AL_SN3D_SOFT)
            al.alBufferData(b, fmt, pcm.ctypes.data, pcm.nbytes, 48000)
            assert al.alGetError() == 0
            al.alSourcei(s, AL_BUFFER, b.value)
            al.alSourcei(s, AL_LOOPING, 0 if oneshot else 1)
        al.alSourcef(s, AL_PITCH, scene.voice_pitch(i))
        al.alSourcef(s, AL_GAIN, scene.voice_gain(V))
        al.alSource3f(s, AL_POSITION, *[float(x) for x in scene.voice_position(i)])
        al.alSourcei(s, AL_SOURCE_RESAMPLER_SOFT, resampler)
        if fx == "reverb":
            al.alSource3i(s, AL_AUXILIARY_SEND_FILTER, slots[0][0], 0, AL_FILTER_NULL)
        elif fx == "misc3" and i in (1, 20, 22):
            al.alSource3i(s, AL_AUXILIARY_SEND_FILTER, slots[0][0], 0, sendfilter.value)
        elif fx == "conv":
            al.alSource3i(s, AL_AUXILIARY_SEND_FILTER, slots[i % 2][0], 0, AL_FILTER_NULL)
        elif fx in ("allfx", "pshift"):
            al.alSource3i(s, AL_AUXILIARY_SEND_FILTER, slots[i % len(slots)][0], 0, AL_FILTER_NULL)
        elif fx == "mix":
            # send 0: reverb or the equalizer that feeds it; send 1: the echo for every third source
            al.alSource3i(s, AL_AUXILIARY_SEND_FILTER, slots[2][0] if i % 4 == 1 else slots[0][0], 0,
                          lowpass.value if (filt and i % 2 == 0) else AL_FILTER_NULL)
            if i % 3 == 0:
                al.alSource3i(s, AL_AUXILIARY_SEND_FILTER, slots[1][0], 1, AL_FILTER_NULL)
        if filt and i % 5 == 0:
            al.alSourcei(s, AL_DIRECT_FILTER, bandpass.value)
        elif filt and i % 2 == 1:
            al.alSourcei(s, AL_DIRECT_FILTER, lowpass.value)
        if fx == "direct" and i == 0:
            al.alSourcei(s, AL_DIRECT_CHANNELS_SOFT, 1)
        sources[i] = s.value
        bufids.append(b.value)
    err = al.alGetError()
    assert err == 0, hex(err)
    ctx2, sources2 = None, None
    if fx == "misc2":
        slots.append(make_slot(AL_EFFECT_EAXREVERB, 0.8))
        slots.append(make_slot(AL_EFFECT_ECHO, 0.7))
        for i in range(V):
            al.alSource3i(sources[i], AL_AUXILIARY_SEND_FILTER, slots[i % 2][0], 0, AL_FILTER_NULL)
        # a second context on the same device: four sources that all play source 1's buffer
        ctx2 = al.alcCreateContext(dev, (C.c_int * len(attrs))(*attrs))
        assert ctx2
        al.alcMakeContextCurrent(ctx2)
        sources2 = (C.c_uint * 4)()
        al.alGenSources(4, sources2)
        al.alListenerf(AL_GAIN, 0.5)
        for k in range(4):
            al.alSourcei(sources2[k], AL_BUFFER, bufids[1])
            al.alSourcei(sources2[k], AL_LOOPING, 1)
            al.alSourcef(sources2[k], AL_PITCH, 0.8 + 0.15 * k)
            al.alSource3f(sources2[k], AL_POSITION, 1.0 - k, 0.5, -1.0)
        al.alSourcePlayv(4, sources2)
        assert al.alGetError() == 0
        al.alcMakeContextCurrent(ctx)
    al.alSourcePlayv(min(V, 200) if fx == "ctx" else V, sources)
    outs, states, offsets = [], [], []
    for u in range(U):
        # the application moves a quarter of its sources, stops one and restarts another
        for i in range(u % 4, V, 4):
            x, y, z = scene.voice_position(i)
            ang = 0.3 * u + 0.1 * i
            cs, sn = math.cos(ang), math.sin(ang)
            al.alSource3f(sources[i], AL_POSITION, float(x * cs - z * sn), float(y), float(x * sn + z * cs))
        if u == 3 and V > 2:
            al.alSourceStop(sources[2])
        if u == 5 and V > 2:
            al.alSourcePlay(sources[2])
        if filt and u == 3:
            # the filter object changes; sources pick it up when it is attached again
            al.alFilterf(lowpass, AL_LOWPASS_GAINHF, 0.7)
            for i in range(1, V, 2):
                if i % 5:
                    al.alSourcei(sources[i], AL_DIRECT_FILTER, lowpass.value)
        if filt and u == 5:
            al.alSourcei(sources[0], AL_DIRECT_FILTER, AL_FILTER_NULL)
            if V > 7:
                al.alSourcei(sources[7], AL_DIRECT_FILTER, bandpass.value)
        if fx in ("bformat", "hoa", "hoadev2", "hoadev3"):
            # the sound field of every B-Format source turns a little each update
            for i in range(0, V, 2):
                ang = 0.4 * u + 0.2 * i
                ori = (C.c_float * 6)(math.sin(ang), 0.0, -math.cos(ang), 0.0, 1.0, 0.0)
                al.alSourcefv(sources[i], AL_ORIENTATION, ori)
        if fx == "short" and u % 2 == 0 and u:
            al.alSourcePlayv(V, sources)
        if fuzz is not None:
            if os.environ.get("AL_RUNNER_FUZZ_LOG"):
                print(f"fuzz: update {u}", file=sys.stderr)
            fuzz_actions(al, fuzz, sources, V, slots, streams, (lowpass.value, bandpass.value), bufids, u, ctx)
        if fx == "ctx" and V > 260:
            if u == 1:
                # a second context appears while the first one plays: its voices follow the first one's
                ctx2 = al.alcCreateContext(dev, (C.c_int * len(attrs))(*attrs))
                assert ctx2
                al.alcMakeContextCurrent(ctx2)
                sources2 = (C.c_uint * 4)()
                al.alGenSources(4, sources2)
                for k in range(4):
                    al.alSourcei(sources2[k], AL_BUFFER, bufids[2 * k])
                    al.alSourcei(sources2[k], AL_LOOPING, 1)
                    al.alSourcef(sources2[k], AL_PITCH, 0.7 + 0.2 * k)
                    al.alSourcef(sources2[k], AL_GAIN, 0.2)
                    al.alSource3f(sources2[k], AL_POSITION, 1.0 - k, 0.5, -1.0)
                al.alSourcePlayv(4, sources2)
                al.alcMakeContextCurrent(ctx)
            if u == 3:
                # the first context outgrows its 256 voices (ContextBase::allocVoices): the second one's move down the list
                rest = (C.c_uint * (V - 200))(*[sources[i] for i in range(200, V)])
                al.alSourcePlayv(V - 200, rest)
            if u == 5:
                # the second context goes away with its sources still playing
                al.alcDestroyContext(ctx2)
                ctx2 = None
        if fx == "rebuf" and u in (2, 4) and V > 6:
            # source 5: stop, swap its buffer for a NEW one of the same size and other content
            k = 5
            al.alSourceStop(sources[k])
            al.alSourcei(sources[k], AL_BUFFER, 0)
            old = C.c_uint(bufids[k])
            al.alDeleteBuffers(1, C.byref(old))
            nb = C.c_uint(0)
            al.alGenBuffers(1, C.byref(nb))
            pcm2 = np.ascontiguousarray(scene.voice_buffer_fast(k + 7 * u, scene.BUFFER_FRAMES))
            keep.append(pcm2)
            al.alBufferData(nb, AL_FORMAT_MONO16, pcm2.ctypes.data, pcm2.nbytes, 48000)
            bufids[k] = nb.value
            al.alSourcei(sources[k], AL_BUFFER, nb.value)
            al.alSourcePlay(sources[k])
        if fx == "misc" and V > 10:
            if u == 1:
                al.alSourcePause(sources[0])
                al.alSourcef(sources[1], AL_PITCH, 1.7)
                al.alSourcef(sources[4], AL_GAIN, 0.05)
            if u == 2:
                al.alSourcei(sources[5], AL_SAMPLE_OFFSET, 30000)        # seek while playing
                al.alListener3f(AL_POSITION, 0.5, 0.0, -0.25)
                al.alSource3f(sources[6], AL_VELOCITY, 30.0, 0.0, 0.0)   # doppler
            if u == 3:
                al.alSourcePlay(sources[0])                              # resume
                al.alListenerf(AL_GAIN, 0.6)
                al.alSourcei(sources[8], AL_LOOPING, 0)
            if u == 4:
                al.alSourceRewind(sources[9])
                al.alSourcef(sources[1], AL_PITCH, 0.61)
            if u == 5:
                al.alSourcePlay(sources[9])
                al.alSourcePlay(sources[10])                             # restart a playing source
        if fx == "misc2" and V > 12:
            if u == 1:
                # a batch of changes becomes visible at once
                al.alcSuspendContext(ctx)
                for i in range(0, 8):
                    al.alSource3f(sources[i], AL_POSITION, 0.3 * i - 1.0, 0.2, -1.5)
                    al.alSourcef(sources[i], AL_GAIN, 0.02 + 0.01 * i)
                al.alcProcessContext(ctx)
            if u == 2:
                al.alSource3i(sources[3], AL_AUXILIARY_SEND_FILTER, slots[0][0], 0, AL_FILTER_NULL)   # echo -> reverb
                al.alSource3i(sources[4], AL_AUXILIARY_SEND_FILTER, 0, 0, AL_FILTER_NULL)             # send removed
            if u == 3:
                al.alcMakeContextCurrent(ctx2)
                al.alSourceStop(sources2[0])
                al.alListener3f(AL_POSITION, 0.0, 0.0, 1.0)
                al.alcMakeContextCurrent(ctx)
            if u == 4:
                al.alSourceStopv(12, sources)                      # the first dozen stop ...
            if u == 5:
                half = (C.c_uint * 6)(*[sources[i] for i in (11, 9, 7, 5, 3, 1)])
                al.alSourcePlayv(6, half)                          # ... and some come back, on other voices
        if fx == "misc3" and V > 22:
            if u == 1:
                al.alSourcePause(sources[0])                               # a streaming source pauses
                al.alSourcei(sources[1], AL_SAMPLE_OFFSET, 2500)           # seek inside a queue (2nd item)
            if u == 2:
                al.alEffecti(slots[0][1], AL_EFFECT_TYPE, AL_EFFECT_NULL)  # the slot's effect goes away ...
                al.alAuxiliaryEffectSloti(slots[0][0], AL_EFFECTSLOT_EFFECT, slots[0][1])
            if u == 3:
                al.alSourcePlay(sources[0])                                # ... the stream resumes
            if u == 4:
                al.alEffecti(slots[0][1], AL_EFFECT_TYPE, AL_EFFECT_EAXREVERB)   # ... and the reverb is back
                al.alAuxiliaryEffectSloti(slots[0][0], AL_EFFECTSLOT_EFFECT, slots[0][1])
            if u == 5:
                al.alSourceStop(sources[2])                                # a playing stream is stopped ...
                al.alSourcePlay(sources[4])                                # ... and another restarted from its head
            if u == 6:
                # source 3's queue (non-looping) has run out: take the buffers back, refill, play again
                done = C.c_int(0)
                al.alGetSourcei(sources[3], AL_BUFFERS_PROCESSED, C.byref(done))
                if done.value > 0:
                    got = (C.c_uint * done.value)()
                    al.alSourceUnqueueBuffers(sources[3], done.value, got)
                    al.alSourceQueueBuffers(sources[3], done.value, got)
                    al.alSourcePlay(sources[3])
        if reset and u == 4:
            # the application switches the output mode while everything plays
            attrs2 = list(attrs)
            attrs2[attrs2.index(ALC_HRTF_SOFT) + 1] = 0 if hrtf else 1
            assert al.alcResetDeviceSOFT(dev, (C.c_int * len(attrs2))(*attrs2))
        if streams and u == 1:
            # the application keeps a stream fed: one more buffer on the first streaming source
            i0, _ = streams[0]
            extra = C.c_uint(0)
            al.alGenBuffers(1, C.byref(extra))
            part = np.ascontiguousarray(scene.voice_buffer_fast(i0 + 5, 2500))
            keep.append(part)
            al.alBufferData(extra, AL_FORMAT_MONO16, part.ctypes.data, part.nbytes, 48000)
            al.alSourceQueueBuffers(sources[i0], 1, C.byref(extra))
        if streams and u == 3:
            # ... and takes back what has been played
            i0, _ = streams[0]
            done = C.c_int(0)
            al.alGetSourcei(sources[i0], AL_BUFFERS_PROCESSED, C.byref(done))
            if done.value > 0:
                got = (C.c_uint * done.value)()
                al.alSourceUnqueueBuffers(sources[i0], done.value, got)
        if fx == "conv" and u == 3:
            al.alAuxiliaryEffectSlotf(slots[0][0], AL_EFFECTSLOT_GAIN, 0.3)
        if fx == "allfx" and u == 3:
            for k, (par, val, isint) in enumerate(((0x0001, 2, True), (0x0002, 1, True), (0x0001, 0.01, False), (0x0001, 0.7, False),
                                                   (0x0001, 0, True), (0x0003, 2, True), (0x0001, 0, True))):
                (al.alEffecti if isint else al.alEffectf)(slots[k][1], par, val)
                al.alAuxiliaryEffectSloti(slots[k][0], AL_EFFECTSLOT_EFFECT, slots[k][1])
        if fx == "pshift" and u == 3:
            al.alEffecti(slots[0][1], 0x0001, 5)                                                   # coarse tune: a fourth up
            al.alAuxiliaryEffectSloti(slots[0][0], AL_EFFECTSLOT_EFFECT, slots[0][1])
            al.alEffecti(slots[1][1], 0x0001, -12)
            al.alEffecti(slots[1][1], 0x0002, 0)
            al.alAuxiliaryEffectSloti(slots[1][0], AL_EFFECTSLOT_EFFECT, slots[1][1])
        if slots and fx not in ("conv", "allfx", "pshift") and u == 2:
            # a property that needs the reverb's other pipeline (full update), then one that does not
            al.alEffectf(slots[0][1], AL_EAXREVERB_DECAY_TIME, 2.9)
            al.alAuxiliaryEffectSloti(slots[0][0], AL_EFFECTSLOT_EFFECT, slots[0][1])
        if slots and fx not in ("conv", "allfx", "pshift") and u == 4:
            al.alEffectf(slots[0][1], AL_EAXREVERB_REFLECTIONS_GAIN, 0.3)
            al.alAuxiliaryEffectSloti(slots[0][0], AL_EFFECTSLOT_EFFECT, slots[0][1])
        if fx == "mix" and u == 3:
            al.alEffectf(slots[1][1], AL_ECHO_DELAY, 0.012)
            al.alAuxiliaryEffectSloti(slots[1][0], AL_EFFECTSLOT_EFFECT, slots[1][1])
            al.alAuxiliaryEffectSlotf(slots[2][0], AL_EFFECTSLOT_GAIN, 0.4)
        if fx == "mix" and u == 6 and not FUZZ_EXT["family3"]:
            # the echo slot becomes a chorus: a new EffectState (deviceUpdate)
            al.alEffecti(slots[1][1], AL_EFFECT_TYPE, AL_EFFECT_CHORUS)
            al.alEffectf(slots[1][1], AL_CHORUS_RATE, 2.2)
            al.alAuxiliaryEffectSloti(slots[1][0], AL_EFFECTSLOT_EFFECT, slots[1][1])
        frames = (1024, 100, 7, 640, 1, 333, 1024, 480, 480, 512)[u % 10] if ragged else 1024
        buf = np.zeros((1024, nout), dtype=np.int16 if out16 else np.float32)
        al.alcRenderSamplesSOFT(dev, buf.ctypes.data, frames)
        outs.append((buf.astype(np.float32) / 32768.0).T.copy() if out16 else buf.T.copy())
        st, off = [], []
        for i in range(min(V, 64)):
            v = C.c_int(0)
            al.alGetSourcei(sources[i], AL_SOURCE_STATE, C.byref(v))
            st.append(v.value)
            al.alGetSourcei(sources[i], AL_SAMPLE_OFFSET, C.byref(v))
            off.append(v.value)
            al.alGetSourcei(sources[i], AL_BUFFERS_PROCESSED, C.byref(v))
            st.append(v.value)
        states.append(st)
        offsets.append(off)
    hv = C.c_int(0)
    al.alcGetIntegerv(dev, 0x1993, 1, C.byref(hv))             # ALC_HRTF_STATUS_SOFT
    cv = C.c_int(1)
    al.alcGetIntegerv(dev, ALC_CONNECTED, 1, C.byref(cv))
    np.savez(out_path, out=np.stack(outs), states=np.array(states), offsets=np.array(offsets), hrtf_status=hv.value,
             connected=cv.value)
    al.alcMakeContextCurrent(None)
    if ctx2:
        al.alcDestroyContext(ctx2)
    al.alcDestroyContext(ctx)
    al.alcCloseDevice(dev)
    os.remove(conf)


if __name__ == "__main__":
    main()
