"""The product's pitch-shifter frame arithmetic (openal-soft_b200/csrc/pshift.hpp — the source the
GPU kernel runs one warp per channel) executed serially on the host (oracle/libpshift_emul.so)
against the oracle's independent restatement of PshifterState::process (oracle/efx_oracle.cpp,
which the reference-rendered fixtures efx_pshifter_* pin): gather form vs scatter form, table
twiddles vs direct ones, ragged update sizes, every shift direction."""
import ctypes as C
import os

import numpy as np
import pytest

from pyb200mix import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _libs():
    emul = C.CDLL(os.path.join(ROOT, "oracle", "libpshift_emul.so"))
    emul.pshift_emul_create.restype = C.c_void_p
    emul.pshift_emul_free.argtypes = [C.c_void_p]
    emul.pshift_emul_set.argtypes = [C.c_void_p, C.c_uint32]
    emul.pshift_emul_process.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    ora = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    ora.oefx_create.restype = C.c_void_p
    ora.oefx_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    ora.oefx_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    ora.oefx_process.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    ora.oefx_free.argtypes = [C.c_void_p]
    return emul, ora


def _props(coarse, fine):
    p = abi.efx_defaults(abi.EFFECT_PSHIFTER)
    p.pshifter.coarse_tune, p.pshifter.fine_tune = coarse, fine
    return p


def _shift_i(coarse, fine):
    pitch = np.float32(2.0) ** np.float32(np.float32(coarse * 100 + fine) / np.float32(1200.0))
    return int(np.rint(np.float32(min(max(pitch, np.float32(0.5)), np.float32(2.0))) * np.float32(65536.0)))


@pytest.mark.parametrize("channels", [1, 4, 9])
def test_product_frame_arithmetic_matches_the_oracle(channels):
    emul, ora = _libs()
    rng = np.random.default_rng(1234 + channels)
    idx = (C.c_uint32 * channels)(*range(channels))
    scale = (C.c_float * channels)(*([1.0] * channels))
    t = abi.EfxTarget()
    t.struct_size = C.sizeof(abi.EfxTarget)
    t.sample_rate, t.slot_gain = 48000, 1.0
    t.out_channels, t.out_scale, t.out_index = channels, C.cast(scale, C.c_void_p), C.cast(idx, C.c_void_p)
    t.wet_channels, t.wet_index = channels, C.cast(idx, C.c_void_p)
    t.real_center = t.real_lfe = 0xffffffff
    t.device_ambi_order = {1: 0, 4: 1, 9: 2}[channels]
    tunes = [(12, 0), (7, 30), (0, 0), (-5, -20), (-12, 0), (3, -50), (-1, 17)]
    rc = C.c_int(0)
    p = _props(*tunes[0])
    oe = ora.oefx_create(C.byref(p), C.byref(t), C.byref(rc))
    assert oe and rc.value == 0
    pe = emul.pshift_emul_create()
    emul.pshift_emul_set(pe, _shift_i(*tunes[0]))
    sizes = [1024, 1024, 517, 3, 128, 1000, 1024, 255, 1, 1024, 640, 1024, 1024, 77, 1024, 1024]
    phase = 0
    worst = 0.0
    peak = 0.0
    for u, n in enumerate(sizes):
        if u and u % 2 == 0:
            tune = tunes[(u // 2) % len(tunes)]
            p = _props(*tune)
            assert ora.oefx_update(oe, C.byref(p), C.byref(t)) == 0
            emul.pshift_emul_set(pe, _shift_i(*tune))
        x = np.zeros((channels, 1024), np.float32)
        if u:      # the first update is silent: the oracle's output gains fade in from zero during it
            k = np.arange(phase, phase + n)
            w = 0.3 * np.sin(2 * np.pi * 440.0 / 48000 * k) + 0.2 * np.sin(2 * np.pi * 3217.0 / 48000 * k + 1.0)
            w = w + 0.1 * rng.standard_normal(n)
            for c in range(channels):
                x[c, :n] = (w * np.cos(0.7 * c) + 0.05 * rng.standard_normal(n) * (c > 0)).astype(np.float32)
            phase += n
        out_o = np.zeros((channels, 1024), np.float32)
        out_e = np.zeros((channels, 1024), np.float32)
        ora.oefx_process(oe, n, x.ctypes.data, channels, out_o.ctypes.data, channels)
        emul.pshift_emul_process(pe, n, channels, x.ctypes.data, out_e.ctypes.data)
        worst = max(worst, float(np.abs(out_o[:, :n] - out_e[:, :n]).max()))
        peak = max(peak, float(np.abs(out_o[:, :n]).max()))
    ora.oefx_free(oe)
    emul.pshift_emul_free(pe)
    assert peak > 0.1                      # the shifter produced sound
    print("worst", worst, "peak", peak)
    assert worst <= 2e-6, worst
