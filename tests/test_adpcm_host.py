"""The product's block decoders (openal-soft_b200/csrc/adpcm.cpp: host code that runs once inside
b200mix_buffer_data_adpcm) against the oracle's restatement of LoadSamples<IMA4Data> /
LoadSamples<MSADPCMData> (core/voice.cpp:289-484), mono and stereo, default and other block sizes.

The oracle has no "give me the decoded buffer" call, so it plays the buffer: point resampler, step
1.0, one dry channel at gain 1 — the dry bus then holds sample/32768 exactly.  The oracle's decoders
themselves are pinned to the reference by the hrtf_spline_adpcm golden and by the "formats" scene of
tests/test_seam_cpu.py (mono and stereo blocks through the live reference)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import mixlib, synth
from helpers.mixlib import MixDevice
from pyb200mix import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "openal-soft_b200", "csrc")

WRAPPER = r"""
#include "adpcm.hpp"
extern "C" void decode(int ms, const uint8_t *src, uint32_t channels, uint32_t spb, size_t blocks, int16_t *dst)
{ if(ms) b200mix::DecodeMSADPCM(src, channels, spb, blocks, dst); else b200mix::DecodeIMA4(src, channels, spb, blocks, dst); }
extern "C" size_t block_bytes(int ms, uint32_t channels, uint32_t spb) { return b200mix::AdpcmBlockBytes(ms != 0, channels, spb); }
extern "C" int block_valid(int ms, uint32_t spb) { return b200mix::AdpcmBlockValid(ms != 0, spb) ? 1 : 0; }
"""


@pytest.fixture(scope="module")
def decoder(tmp_path_factory):
    d = tmp_path_factory.mktemp("adpcm")
    src = d / "wrap.cpp"
    src.write_text(WRAPPER)
    so = d / "libadpcm_host.so"
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", CSRC, str(src),
                    os.path.join(CSRC, "adpcm.cpp"), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.decode.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_void_p]
    lib.decode.restype = None
    lib.block_bytes.argtypes = [C.c_int, C.c_uint32, C.c_uint32]
    lib.block_bytes.restype = C.c_size_t
    lib.block_valid.argtypes = [C.c_int, C.c_uint32]
    return lib


def _oracle_pcm(kind, channels, spb, blocks, data):
    """The oracle's decoded samples, read back off the dry bus."""
    frames = spb * blocks
    assert frames <= abi.LINE
    desc = synth.stereo_desc(channels, dry_channels=3)
    dev = MixDevice(mixlib.oracle(), desc)
    dev.set_ambi_decoder(np.eye(3, 2, dtype=np.float32), None, 0.0)
    dev.buffer_data_adpcm(0, kind, spb, blocks, data, channels=channels)
    out = np.zeros((frames, channels), dtype=np.float32)
    for c in range(channels):
        p = abi.VoiceParams()
        p.voice = 0
        p.flags = abi.VF_PLAYING | abi.VF_STATIC | abi.VF_RESET | abi.vf_channel(c)
        p.buffer = 0
        p.resampler = abi.RS_POINT
        p.position, p.position_frac, p.step = 0, 0, 65536
        p.loop_start, p.loop_end = 0, frames
        for s in range(abi.MAX_SENDS):
            p.send_slot[s] = abi.NO_SLOT
        dev.voices_update([p], None, np.array([[1.0, 0.0, 0.0]], dtype=np.float32), None)
        dev.render(frames)
        out[:, c] = dev.dry()[0, :frames]
    dev.close()
    return out


@pytest.mark.parametrize("kind,channels,spb,blocks", [
    (abi.FMT_IMA4, 1, 65, 15), (abi.FMT_IMA4, 2, 65, 15), (abi.FMT_IMA4, 1, 9, 40), (abi.FMT_IMA4, 2, 129, 7),
    (abi.FMT_IMA4, 1, 1, 50),
    (abi.FMT_MSADPCM, 1, 64, 16), (abi.FMT_MSADPCM, 2, 64, 16), (abi.FMT_MSADPCM, 1, 2, 60),
    (abi.FMT_MSADPCM, 2, 128, 8), (abi.FMT_MSADPCM, 1, 30, 30)])
def test_host_block_decoders_match_the_oracle(decoder, kind, channels, spb, blocks):
    ms = int(kind == abi.FMT_MSADPCM)
    assert decoder.block_valid(ms, spb)
    nbytes = decoder.block_bytes(ms, channels, spb)
    assert nbytes == (((spb - 2) // 2 + 7) if ms else ((spb - 1) // 2 + 4)) * channels
    rng = np.random.default_rng(1000 * kind + 10 * spb + channels)
    # every byte pattern is a stream the reference decodes: out-of-range predictor / step indices are
    # clamped, sums saturate at the int16 rails
    data = rng.integers(0, 256, size=blocks * nbytes, dtype=np.uint8)
    got = np.zeros((blocks * spb, channels), dtype=np.int16)
    decoder.decode(ms, data.ctypes.data, channels, spb, blocks, got.ctypes.data)
    want = _oracle_pcm(kind, channels, spb, blocks, data)
    assert np.abs(want).max() > 0.1
    assert np.array_equal(got.astype(np.float32) / np.float32(32768.0), want)


def test_block_size_rules(decoder):
    # al/buffer.cpp:270-300: IMA4 blocks hold 1 + 8k samples, MSADPCM blocks an even count >= 2
    assert [s for s in range(0, 40) if decoder.block_valid(0, s)] == [1, 9, 17, 25, 33]
    assert [s for s in range(0, 9) if decoder.block_valid(1, s)] == [2, 4, 6, 8]
