"""Pins the oracle restatement against the compiled reference (oracle/_ref) live:
coefficient tables and BsincPrepare bit-for-bit, resampler kernels bit-for-bit vs
the reference's C kernels and within rounding of its SSE kernels.
Skipped when oracle/_ref is absent (it is built by __graft_entry__.build() when
/root/reference exists, and travels to the GPU box)."""
import ctypes as C

import numpy as np
import pytest

from helpers import mixlib, refal
from pyb200mix import abi

pytestmark = pytest.mark.ref


def _oracle_lib():
    lib = mixlib.oracle().lib
    lib.oracle_get_resampler_table.restype = C.c_int64
    lib.oracle_get_resampler_table.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
    lib.oracle_get_bsinc_state.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_float),
                                           C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                           C.POINTER(C.c_uint32)]
    lib.oracle_resample.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                    C.c_uint32]
    return lib


@pytest.mark.parametrize("which", [abi.RS_BSINC12, abi.RS_BSINC24, abi.RS_BSINC48])
def test_bsinc_tables_bit_exact(which):
    _, hz = refal.libs()
    lib = _oracle_lib()
    n_ref = hz.refh_bsinc_table(which, None, 0)
    n = lib.oracle_get_resampler_table(which, None, 0)
    assert n == n_ref and n > 0
    a = np.zeros(n, dtype=np.float32)
    b = np.zeros(n, dtype=np.float32)
    hz.refh_bsinc_table(which, a.ctypes.data, n)
    lib.oracle_get_resampler_table(which, b.ctypes.data, n)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("which,idx", [(abi.RS_SPLINE, 0), (abi.RS_GAUSSIAN, 1)])
def test_cubic_tables_bit_exact(which, idx):
    _, hz = refal.libs()
    lib = _oracle_lib()
    a = np.zeros(256, dtype=np.float32)
    b = np.zeros(256, dtype=np.float32)
    hz.refh_cubic_table(idx, a.ctypes.data)
    assert lib.oracle_get_resampler_table(which, b.ctypes.data, 256) == 256
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("which", [abi.RS_BSINC12, abi.RS_BSINC24, abi.RS_BSINC48])
def test_bsinc_prepare_bit_exact(which):
    _, hz = refal.libs()
    lib = _oracle_lib()
    rng = np.random.default_rng(1)
    incs = list(rng.integers(1, 10 << 16, size=400)) + [65535, 65536, 65537, 131072, 655360]
    for inc in incs:
        r = [C.c_float(), C.c_uint32(), C.c_uint32(), C.c_uint32()]
        o = [C.c_float(), C.c_uint32(), C.c_uint32(), C.c_uint32()]
        hz.refh_bsinc_state(which, int(inc), *[C.byref(x) for x in r])
        lib.oracle_get_bsinc_state(which, int(inc), *[C.byref(x) for x in o])
        assert [x.value for x in r] == [x.value for x in o], inc


@pytest.mark.parametrize("resampler", range(10))
def test_resamplers_vs_reference_kernels(resampler):
    _, hz = refal.libs()
    lib = _oracle_lib()
    rng = np.random.default_rng(resampler)
    for inc in [1, 7000, 32768, 65535, 65536, 65537, 70000, 98304, 131072, 200000, 655360]:
        for frac in [0, 1, 2047, 2048, 40000, 65535]:
            n_out = 257
            need = ((n_out * inc + frac) >> 16) + 100
            src = rng.uniform(-1, 1, size=need).astype(np.float32)
            a = np.zeros(n_out, dtype=np.float32)
            b = np.zeros(n_out, dtype=np.float32)
            s = np.zeros(n_out, dtype=np.float32)
            hz.refh_resample(resampler, 0, inc, frac, src.ctypes.data, need, a.ctypes.data, n_out)
            hz.refh_resample(resampler, 1, inc, frac, src.ctypes.data, need, s.ctypes.data, n_out)
            lib.oracle_resample(resampler, inc, frac, src.ctypes.data, b.ctypes.data, n_out)
            assert np.array_equal(a, b), (resampler, inc, frac)
            assert np.abs(s.astype(np.float64) - b).max() <= 2e-6, (resampler, inc, frac)


def test_biquad_coeffs_bit_exact():
    """BiquadFilter::SetParams via setParamsFromSlope vs the oracle restatement."""
    _, hz = refal.libs()
    lib = mixlib.oracle().lib
    lib.oracle_biquad_coeffs.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_void_p]
    rng = np.random.default_rng(7)
    for _ in range(300):
        typ = int(rng.integers(0, 6))
        f0 = float(rng.uniform(1e-4, 0.6))
        gain = float(10.0 ** rng.uniform(-4.0, 1.0))
        slope = float(rng.uniform(0.05, 1.0))
        a = np.zeros(5, dtype=np.float32)
        b = np.zeros(5, dtype=np.float32)
        hz.refh_biquad_coeffs(typ, f0, gain, slope, a.ctypes.data)
        assert lib.oracle_biquad_coeffs(typ, f0, gain, slope, b.ctypes.data) == 0
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (typ, f0, gain, slope, a, b)


@pytest.mark.parametrize("hrtf", [1, 0])
def test_filters_ragged_updates_vs_reference_live(hrtf):
    """Pins the oracle's BiquadInterpFilter stepping across update boundaries (partial
    32-sample steps, restarts while interpolating, detach/clear) against the live reference:
    ragged update sizes, filters changed between updates through the AL filter API."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from helpers import scenes
    from pyb200mix import scene
    V = 5
    ref, pcms = scenes.make_ref_scene(V, hrtf, abi.RS_SPLINE)
    script = {0: [(0, 0.25, None), (1, 0.5, 0.3), (2, 0.1, None)],
              1: [(0, 0.9, None), (3, 0.2, 0.7)],
              2: [(0, 0.4, None), (1, 1.0, None)],      # restart while interpolating; detach
              4: [(1, 0.6, 0.6), (2, 0.8, None)],
              5: [(3, 1.0, None), (4, 0.05, 0.5)]}
    sizes = [1024, 37, 500, 1, 1000, 64, 333, 1024]

    def apply(u):
        for voice, ghf, glf in script.get(u, []):
            filt = refal.AL_FILTER_NULL if (ghf == 1.0 and glf is None) else ref.make_filter(1.0, ghf, glf)
            ref.set_direct_filter(ref.sources[voice], filt)

    apply(0)
    ref.play_all()
    dev = None
    try:
        for u, n in enumerate(sizes):
            if u:
                apply(u)
            out_ref = ref.render(n)
            if dev is None:
                dev = scenes.mirror_device(mixlib.oracle(), ref, V, pcms)
                scenes.feed_params(dev, ref, True, V)
            ents, _ = ref.voice_filters(V)
            dev.voices_filters(ents)
            out = dev.render(n)
            # the live reference runs its SSE kernels here (the oracle follows the C kernels and
            # is bit-exact with them on the golden filter scenes): rounding differences of the
            # resamplers pass through the shelving filters, hence 2e-6 instead of 1e-7
            err = np.abs(out.astype(np.float64) - out_ref).max()
            assert err <= 2e-6, (u, n, err)
            assert np.abs(out_ref).max() > 1e-4
    finally:
        if dev is not None:
            dev.close()
        ref.close()


@pytest.mark.parametrize("hrtf", [1, 0])
def test_limiter_ragged_updates_vs_reference_live(hrtf):
    """Pins the oracle's Compressor (look-ahead FIFO, sliding hold, envelope state carried
    across updates) against the live reference: float output with ALC_OUTPUT_LIMITER_SOFT, a mix
    driven past full scale, update sizes below and above the look-ahead (48) and hold (96)."""
    import ctypes as C
    from helpers import scenes
    V = 6
    ref, pcms = scenes.make_ref_scene(V, hrtf, abi.RS_SPLINE, attrs={refal.ALC_OUTPUT_LIMITER_SOFT: 1})
    ref.al.alListenerf.argtypes = [C.c_int, C.c_float]
    ref.al.alListenerf(refal.AL_GAIN, 6.0)
    sizes = [1024, 37, 500, 1, 1000, 64, 20, 20, 100, 333, 1024]
    ref.play_all()
    dev = None
    peak = 0.0
    try:
        for u, n in enumerate(sizes):
            out_ref = ref.render(n)
            if dev is None:
                dev = scenes.mirror_device(mixlib.oracle(), ref, V, pcms)
                scenes.feed_params(dev, ref, True, V)
                ld, la = ref.limiter_desc()
                assert dev.set_limiter(ld) == la == 48
            out = dev.render(n)
            err = np.abs(out.astype(np.float64) - out_ref).max()
            assert err <= 2e-6, (u, n, err)
            peak = max(peak, float(np.abs(out_ref).max()))
    finally:
        if dev is not None:
            dev.close()
        ref.close()
    # the limiter did hold the mix at full scale (the unlimited mix peaks far above 1)
    assert 0.5 < peak <= 1.0 + 1e-6, peak


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5, 6])
def test_bs2b_crossfeed_bit_exact(level):
    """The oracle's BS2B coefficients and cross-feed recurrences against the reference's
    Bs2b::bs2b_processor (kernel-level tap), chained over ragged chunks."""
    import ctypes as C
    _, hz = refal.libs()
    hz.refh_bs2b_cross_feed.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 2
    hz.refh_bs2b_cross_feed.restype = None
    lib = _oracle_lib()
    lib.oracle_bs2b_coeffs.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
    lib.oracle_bs2b_cross_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.oracle_bs2b_cross_feed.restype = None
    rng = np.random.default_rng(100 + level)
    for srate in (44100, 48000, 96000):
        coef_o = np.zeros(5, dtype=np.float32)
        assert lib.oracle_bs2b_coeffs(level, srate, coef_o.ctypes.data) == 0
        st_r = np.zeros(4, dtype=np.float32)
        st_o = np.zeros(4, dtype=np.float32)
        coef_r = np.zeros(5, dtype=np.float32)
        for n in (1024, 37, 500, 1, 129, 1000):
            l = (rng.standard_normal(n) * 0.3).astype(np.float32)
            r = (rng.standard_normal(n) * 0.3).astype(np.float32)
            lr, rr = l.copy(), r.copy()
            hz.refh_bs2b_cross_feed(level, srate, lr.ctypes.data, rr.ctypes.data, n, st_r.ctypes.data,
                                    coef_r.ctypes.data)
            assert np.array_equal(coef_o.view(np.uint32), coef_r.view(np.uint32)), (coef_o, coef_r)
            lo, ro = l.copy(), r.copy()
            lib.oracle_bs2b_cross_feed(coef_o.ctypes.data, st_o.ctypes.data, lo.ctypes.data, ro.ctypes.data, n)
            assert np.array_equal(lo.view(np.uint32), lr.view(np.uint32)), (srate, n)
            assert np.array_equal(ro.view(np.uint32), rr.view(np.uint32)), (srate, n)
            assert np.array_equal(st_o.view(np.uint32), st_r.view(np.uint32))
