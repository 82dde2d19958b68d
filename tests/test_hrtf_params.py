"""Host-side parameter helper: the MHR loader + HrtfStore::getCoeffs restatement in the
product (openal-soft_b200/csrc/hrtf_store.cpp) against the HRIRs the compiled reference
computed for the same source directions (stored in the golden fixtures)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from helpers import golden, mixlib
from pyb200mix import scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MHR = os.path.join(ROOT, "openal-soft_b200", "data", "Default HRTF.mhr")


def _lib():
    lib = C.CDLL(mixlib.PRODUCT_SO)
    lib.b200mix_hrtf_load.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.b200mix_hrtf_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 3
    lib.b200mix_hrtf_get_coeffs.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                            C.c_void_p, C.POINTER(C.c_uint32)]
    lib.b200mix_hrtf_free.argtypes = [C.c_void_p]
    return lib


@pytest.mark.skipif(not os.path.exists(MHR), reason="HRTF data set not staged (run build())")
def test_get_coeffs_matches_reference_alu():
    lib = _lib()
    data = open(MHR, "rb").read()
    h = C.c_void_p()
    assert lib.b200mix_hrtf_load(data, len(data), C.byref(h)) == 0
    rate, irs, cnt = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib.b200mix_hrtf_info(h, C.byref(rate), C.byref(irs), C.byref(cnt))
    assert (rate.value, irs.value, cnt.value) == (48000, 64, 1982)
    from pyb200mix import abi
    for name in ("hrtf_bsinc24_v8", "hrtf_fastbsinc12_v6", "hrtf_bsinc48_v4"):
        fx = golden.load(name)
        V = int(fx["meta"][0])
        params = (abi.VoiceParams * V).from_buffer_copy(fx["params"].tobytes())
        for i in range(V):
            x, y, z = scene.voice_position(i)
            d = math.sqrt(x * x + y * y + z * z)
            # CalcHrtfPanning, alc/alu.cpp:1210-1216
            ev = np.arcsin(np.float32(max(-1.0, min(1.0, y / d))))
            az = np.arctan2(np.float32(x / d), np.float32(-z / d))
            out = np.zeros((64, 2), dtype=np.float32)
            dl = (C.c_uint32 * 2)()
            assert lib.b200mix_hrtf_get_coeffs(h, float(ev), float(az), float(d), 0.0,
                                               out.ctypes.data, dl) == 0
            assert np.abs(out - fx["coeffs"][i]).max() <= 2e-6, (name, i)
            assert list(dl) == list(params[i].hrtf_delay), (name, i)
    lib.b200mix_hrtf_free(h)


def test_loader_rejects_garbage():
    lib = _lib()
    h = C.c_void_p()
    junk = b"MinPHR03" + bytes(40)
    assert lib.b200mix_hrtf_load(junk, len(junk), C.byref(h)) != 0
    assert lib.b200mix_hrtf_load(b"RIFFxxxxWAVEfmt xxxx", 20, C.byref(h)) != 0


@pytest.mark.ref
@pytest.mark.skipif(not os.path.exists(MHR), reason="HRTF data set not staged (run build())")
def test_build_decoder_matches_the_references_direct_hrtf_state():
    """b200mix_hrtf_build_decoder against the DirectHrtfState of a live HRTF device of the compiled
    reference (default data set, hrtf-mode full): coefficients, HF scales, splitter coefficient and
    the decoder's IR length, bit for bit."""
    from helpers import refal
    lib = _lib()
    lib.b200mix_hrtf_build_decoder.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p,
                                               C.c_void_p, C.POINTER(C.c_float)]
    data = open(MHR, "rb").read()
    h = C.c_void_p()
    assert lib.b200mix_hrtf_load(data, len(data), C.byref(h)) == 0
    ref = refal.RefDevice({refal.ALC_HRTF_SOFT: 1})
    try:
        want_c, want_hf, want_sc = ref.hrtf_decoder()
        irs = C.c_uint32(0)
        got = np.zeros(4 * 128 * 2, dtype=np.float32)
        hf = np.zeros(4, dtype=np.float32)
        sc = C.c_float(0.0)
        assert lib.b200mix_hrtf_build_decoder(h, 1, ref.desc.ir_size, C.byref(irs), got.ctypes.data, hf.ctypes.data,
                                              C.byref(sc)) == 4
        assert irs.value == want_c.shape[1], (irs.value, want_c.shape)
        got = got[:4 * irs.value * 2].reshape(4, irs.value, 2)
        assert np.array_equal(got.view(np.uint32), want_c.view(np.uint32)), np.abs(got - want_c).max()
        assert np.array_equal(hf.view(np.uint32), want_hf.view(np.uint32)), (hf, want_hf)
        assert np.float32(sc.value).view(np.uint32) == want_sc[0].view(np.uint32)
        assert lib.b200mix_hrtf_build_decoder(h, 3, 0, C.byref(irs), got.ctypes.data, hf.ctypes.data, C.byref(sc)) < 0
    finally:
        ref.close()
        lib.b200mix_hrtf_free(h)
