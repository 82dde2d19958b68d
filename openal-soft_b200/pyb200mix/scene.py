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
           "alaw": (6, 0x10016), "ima4": (7, 0x1300), "msadpcm": (8, 0x1302)}
ADPCM_BLOCK = {"ima4": (65, 36), "msadpcm": (64, 38)}     # AL default block: (samples, bytes) mono

_IMA_STEP = [7, 8, 9, 10, 11, 12, 13, 14, 16, 17, 19, 21, 23, 25, 28, 31, 34, 37, 41, 45, 50, 55, 60, 66, 73, 80,
             88, 97, 107, 118, 130, 143, 157, 173, 190, 209, 230, 253, 279, 307, 337, 371, 408, 449, 494, 544,
             598, 658, 724, 796, 876, 963, 1060, 1166, 1282, 1411, 1552, 1707, 1878, 2066, 2272, 2499, 2749,
             3024, 3327, 3660, 4026, 4428, 4871, 5358, 5894, 6484, 7132, 7845, 8630, 9493, 10442, 11487, 12635,
             13899, 15289, 16818, 18500, 20350, 22358, 24633, 27086, 29794, 32767]
_IMA_ADJ = [-1, -1, -1, -1, 2, 4, 6, 8]


def adpcm_blocks(i: int, kind: str, blocks: int) -> np.ndarray:
    """Mono block-compressed test data for voice i (bytes as alBufferData takes them).
    ima4: the voice's synthetic signal through a plain IMA encoder (65-sample blocks).
    msadpcm: well-formed 64-sample blocks (predictor, scale, two history samples) whose
    nibbles come from a seeded generator — every byte pattern is a valid stream."""
    spb, nbytes = ADPCM_BLOCK[kind]
    out = np.zeros((blocks, nbytes), dtype=np.uint8)
    if kind == "ima4":
        pcm = voice_buffer_i16(i, blocks * spb).astype(np.int64)
        idx = 0
        for b in range(blocks):
            blk = pcm[b * spb:(b + 1) * spb]
            pred = int(blk[0])
            out[b, 0], out[b, 1] = pred & 0xFF, (pred >> 8) & 0xFF
            out[b, 2], out[b, 3] = idx & 0xFF, 0
            for n in range(spb - 1):
                step = _IMA_STEP[idx]
                diff = int(blk[n + 1]) - pred
                code = 8 if diff < 0 else 0
                code |= min(7, (abs(diff) * 4) // step)
                delta = (2 * (code & 7) + 1) * step // 8
                pred = max(-32768, min(32767, pred - delta if code & 8 else pred + delta))
                idx = max(0, min(88, idx + _IMA_ADJ[code & 7]))
                out[b, 4 + (n >> 1)] |= code << ((n & 1) * 4)
        return out.reshape(-1)
    rng = np.random.default_rng(0xAD9C + i)
    pcm = voice_buffer_i16(i, blocks * 2)
    for b in range(blocks):
        out[b, 0] = (i + b) % 7
        out[b, 1], out[b, 2] = 48, 0                     # scale 48
        h0, h1 = int(pcm[2 * b + 1]) & 0xFFFF, int(pcm[2 * b]) & 0xFFFF
        out[b, 3], out[b, 4] = h0 & 0xFF, h0 >> 8
        out[b, 5], out[b, 6] = h1 & 0xFF, h1 >> 8
        # small nibbles (-2..2) keep the random walk away from the rails most of the time
        nib = rng.choice(np.array([0, 1, 2, 15, 14, 0, 1, 15], dtype=np.uint8), size=(nbytes - 7) * 2)
        out[b, 7:] = (nib[0::2] << 4) | nib[1::2]
    return out.reshape(-1)



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
