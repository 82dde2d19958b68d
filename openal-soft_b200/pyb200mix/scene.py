"""Seeded synthetic voices of SURVEY.md §8(d) — identical inputs for the reference
arm, the oracle and the CUDA mixer.  Pure numpy; no compute from oracle/."""
import math
import numpy as np

BUFFER_FRAMES = 48000
BUFFER_RATE = 48000


def voice_pitch(i: int) -> float:
    """p_i in [0.5, 2.0) from a multiplicative hash; 1/16 of voices at exactly 1.0
    (hits the resampler bypass, core/voice.cpp:764-766)."""
    if i % 16 == 15:
        return 1.0
    h = (i * 2654435761) & 0xFFFFFFFF
    return 0.5 + 1.5 * (h / 4294967296.0)


def voice_position(i: int, radius: float = 2.0):
    """Golden-angle azimuth, 13 elevation rings; OpenAL coords (-Z forward, +Y up)."""
    az = math.radians((i * 137.508) % 360.0)
    ev = math.asin(((i % 13) - 6) / 7.0)
    x = radius * math.cos(ev) * math.sin(az)
    y = radius * math.sin(ev)
    z = -radius * math.cos(ev) * math.cos(az)
    return (x, y, z)


def voice_buffer_i16(i: int, frames: int = BUFFER_FRAMES) -> np.ndarray:
    """Even i: sine 110*2^((i mod 60)/12) Hz, amplitude 0.25.  Odd i: xorshift32 white
    noise in +-0.25, seed 0x9E3779B9 ^ i."""
    if i % 2 == 0:
        f = 110.0 * 2.0 ** ((i % 60) / 12.0)
        t = np.arange(frames, dtype=np.float64)
        s = 0.25 * np.sin(2.0 * np.pi * f * t / BUFFER_RATE)
        return np.round(s * 32767.0).astype(np.int16)
    # vectorised xorshift32 is sequential; generate with a small python-free trick:
    # iterate in numpy over uint32 scalars in chunks (frames is small).
    out = np.empty(frames, dtype=np.int16)
    x = np.uint32((0x9E3779B9 ^ i) & 0xFFFFFFFF)
    if x == 0:
        x = np.uint32(1)
    xs = int(x)
    vals = np.empty(frames, dtype=np.uint32)
    for k in range(frames):
        xs ^= (xs << 13) & 0xFFFFFFFF
        xs ^= xs >> 17
        xs ^= (xs << 5) & 0xFFFFFFFF
        vals[k] = xs
    u = vals.astype(np.float64) / 4294967296.0  # [0,1)
    out[:] = np.round((u * 2.0 - 1.0) * 0.25 * 32767.0).astype(np.int16)
    return out


_NOISE_CACHE = {}


def voice_buffer_fast(i: int, frames: int = BUFFER_FRAMES) -> np.ndarray:
    """Same as voice_buffer_i16 but noise buffers are produced by a vectorised
    generator (numpy PCG seeded with i) — used at bench scale where 10^5+ python
    xorshift loops are too slow.  NOT bit-identical to voice_buffer_i16 for odd i;
    both arms of a comparison must use the same generator."""
    if i % 2 == 0:
        return voice_buffer_i16(i, frames)
    rng = np.random.Generator(np.random.PCG64(0x9E3779B9 ^ i))
    u = rng.random(frames)
    return np.round((u * 2.0 - 1.0) * 0.25 * 32767.0).astype(np.int16)


def voice_gain(num_voices: int) -> float:
    return 1.0 / math.sqrt(num_voices)


# sample formats for the format-coverage scenes: (abi FMT_*, AL format enum)
FORMATS = {"i16": (1, 0x1101), "u8": (0, 0x1100), "f32": (3, 0x10010), "mulaw": (5, 0x10014),
           "alaw": (6, 0x10016)}


def voice_buffer_fmt(i: int, frames: int, fmt: str) -> np.ndarray:
    """The voice's synthetic signal in another storage format (core/fmt_traits.h)."""
    pcm = voice_buffer_i16(i, frames)
    if fmt == "i16":
        return pcm
    if fmt == "u8":
        return ((pcm.astype(np.int32) >> 8) + 128).astype(np.uint8)
    if fmt == "f32":
        return (pcm.astype(np.float32) / np.float32(32768.0)).astype(np.float32)
    if fmt in ("mulaw", "alaw"):
        # any byte is a valid G.711 code: use a deterministic byte pattern of the signal
        return ((pcm.astype(np.int32) >> 7) & 0xFF).astype(np.uint8)
    raise ValueError(fmt)
