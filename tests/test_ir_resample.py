"""Host-side IR rate conversion of the product (b200mix_resample_ir, no GPU involved) against the
reference's PPhaseResampler through the kernel-level tap of oracle/ref_harness.cpp: bit for bit."""
import ctypes as C

import numpy as np
import pytest

from helpers import mixlib, refal

pytestmark = pytest.mark.ref


@pytest.mark.parametrize("src,dst", [(44100, 48000), (48000, 44100), (22050, 48000), (96000, 48000),
                                     (32000, 48000), (11025, 44100), (48000, 48000)])
def test_resample_ir_bit_exact(src, dst):
    prod = mixlib.product().lib
    prod.b200mix_resample_ir.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    prod.b200mix_resampled_ir_frames.argtypes = [C.c_uint32] * 3
    prod.b200mix_resampled_ir_frames.restype = C.c_int64
    _, hz = refal.libs()
    hz.refh_pphase_resample.argtypes = [C.c_uint, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
    hz.refh_pphase_resample.restype = None
    rng = np.random.default_rng(src + dst)
    for n in (1, 7, 500, 4801):
        ir = (rng.standard_normal(n) * np.exp(-np.arange(n) / max(n / 4.0, 1.0))).astype(np.float32)
        m = int(prod.b200mix_resampled_ir_frames(src, dst, n))
        assert m == (n * dst + src - 1) // src
        got = np.zeros(m, dtype=np.float32)
        assert prod.b200mix_resample_ir(src, dst, ir.ctypes.data, n, got.ctypes.data, m) == 0
        if src == dst:
            assert np.array_equal(got, ir)
            continue
        want = np.zeros(m, dtype=np.float32)
        hz.refh_pphase_resample(src, dst, ir.ctypes.data, n, want.ctypes.data, m)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (src, dst, n, np.abs(got - want).max())
        assert np.abs(want).max() > 0


def test_resample_ir_rejects_bad_arguments():
    prod = mixlib.product().lib
    prod.b200mix_resample_ir.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    x = np.zeros(4, dtype=np.float32)
    assert prod.b200mix_resample_ir(0, 48000, x.ctypes.data, 4, x.ctypes.data, 4) < 0
    assert prod.b200mix_resample_ir(44100, 48000, None, 4, x.ctypes.data, 4) < 0
