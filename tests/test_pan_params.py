"""Host-side panning helpers of the product (b200mix_ambi_coeffs, b200mix_pan_gains; no GPU
involved) against the compiled reference's CalcDirectionCoeffs / ComputePanGains through the
kernel-level taps of oracle/ref_harness.cpp: bit for bit."""
import ctypes as C

import numpy as np
import pytest

from helpers import mixlib, refal

pytestmark = pytest.mark.ref


def _bind():
    prod = mixlib.product().lib
    prod.b200mix_ambi_coeffs.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
    prod.b200mix_pan_gains.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                       C.c_void_p, C.c_uint32]
    _, hz = refal.libs()
    hz.refh_calc_direction_coeffs.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
    hz.refh_calc_direction_coeffs.restype = None
    hz.refh_dry_ambi_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    hz.refh_dry_pan_gains.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    hz.refh_dry_pan_gains.restype = None
    return prod, hz


def _dirs(rng, n):
    v = rng.standard_normal((n, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float32)
    return np.concatenate([axes, v.astype(np.float32)])


def test_ambi_coeffs_bit_exact():
    prod, hz = _bind()
    rng = np.random.default_rng(71)
    spreads = [0.0, 0.0, 1e-3, 0.5, 1.0, np.pi, 5.0, 2.0 * np.pi]
    for k, d in enumerate(_dirs(rng, 1500)):
        spread = float(spreads[k % len(spreads)] if k % 3 else rng.uniform(0.0, 2.0 * np.pi))
        a = np.zeros(25, dtype=np.float32)
        b = np.zeros(25, dtype=np.float32)
        dd = np.ascontiguousarray(d)
        hz.refh_calc_direction_coeffs(dd.ctypes.data, spread, a.ctypes.data)
        assert prod.b200mix_ambi_coeffs(dd.ctypes.data, spread, b.ctypes.data) == 0
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (d, spread, a, b)


@pytest.mark.parametrize("attrs", ["stereo", "hrtf", "ambi3"])
def test_pan_gains_on_reference_devices_bit_exact(attrs):
    prod, hz = _bind()
    a = {"stereo": {refal.ALC_HRTF_SOFT: 0},
         "hrtf": {refal.ALC_HRTF_SOFT: 1},
         "ambi3": {refal.ALC_FORMAT_CHANNELS_SOFT: refal.ALC_BFORMAT3D_SOFT, refal.ALC_AMBISONIC_ORDER_SOFT: 3,
                   refal.ALC_AMBISONIC_LAYOUT_SOFT: refal.ALC_ACN_SOFT,
                   refal.ALC_AMBISONIC_SCALING_SOFT: refal.ALC_N3D_SOFT}}[attrs]
    ref = refal.RefDevice(a)
    try:
        scale = np.zeros(32, dtype=np.float32)
        index = np.zeros(32, dtype=np.uint32)
        n = hz.refh_dry_ambi_map(ref.dev, scale.ctypes.data, index.ctypes.data)
        assert n == ref.desc.dry_channels
        rng = np.random.default_rng(72)
        for d in _dirs(rng, 300):
            spread = float(rng.uniform(0.0, 3.0)) if rng.random() < 0.5 else 0.0
            gain = float(rng.uniform(0.0, 2.0))
            co = np.zeros(25, dtype=np.float32)
            dd = np.ascontiguousarray(d)
            assert prod.b200mix_ambi_coeffs(dd.ctypes.data, spread, co.ctypes.data) == 0
            g_ref = np.zeros(25, dtype=np.float32)
            hz.refh_dry_pan_gains(ref.dev, co.ctypes.data, gain, g_ref.ctypes.data)
            g = np.full(32, 7.0, dtype=np.float32)
            assert prod.b200mix_pan_gains(n, scale.ctypes.data, index.ctypes.data, co.ctypes.data, gain,
                                          g.ctypes.data, 32) == 0
            assert np.array_equal(g[:25].view(np.uint32), g_ref.view(np.uint32)), (d, spread, gain)
            assert not g[25:].any()
    finally:
        ref.close()


def test_pan_helpers_reject_bad_arguments():
    prod, _ = _bind()
    co = np.zeros(25, dtype=np.float32)
    d = np.array([0, 0, -1], dtype=np.float32)
    assert prod.b200mix_ambi_coeffs(None, 0.0, co.ctypes.data) < 0
    assert prod.b200mix_ambi_coeffs(d.ctypes.data, 0.0, None) < 0
    scale = np.ones(4, dtype=np.float32)
    bad = np.array([0, 1, 2, 25], dtype=np.uint32)
    g = np.zeros(4, dtype=np.float32)
    assert prod.b200mix_pan_gains(4, scale.ctypes.data, bad.ctypes.data, co.ctypes.data, 1.0, g.ctypes.data, 4) < 0
    assert prod.b200mix_pan_gains(5, scale.ctypes.data, bad.ctypes.data, co.ctypes.data, 1.0, g.ctypes.data, 4) < 0


@pytest.mark.parametrize("devname", ["hrtf", "stereo", "ambi3"])
def test_convolution_gains_bit_exact(devname):
    """b200mix_convolution_gains against ConvolutionState::update on live convolution slots with
    mono, stereo, quad, 5.1 and 7.1 impulse responses (oracle/ref_conv_tap.cpp reads mChans[].Target)."""
    import os
    prod, hz = _bind()
    prod.b200mix_convolution_gains.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_uint32]
    hz.refh_device_render_mode.argtypes = [C.c_void_p]
    tap = C.CDLL(os.path.join(refal.REF_DIR, "libref_conv_tap.so"))
    tap.refh_conv_gains.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    attrs = {"stereo": {refal.ALC_HRTF_SOFT: 0}, "hrtf": {refal.ALC_HRTF_SOFT: 1},
             "ambi3": {refal.ALC_FORMAT_CHANNELS_SOFT: refal.ALC_BFORMAT3D_SOFT, refal.ALC_AMBISONIC_ORDER_SOFT: 3,
                       refal.ALC_AMBISONIC_LAYOUT_SOFT: refal.ALC_ACN_SOFT,
                       refal.ALC_AMBISONIC_SCALING_SOFT: refal.ALC_N3D_SOFT}}[devname]
    # AL float32 formats: mono, stereo, quad, 5.1, 7.1 and the layout ids of the helper
    cases = [(0x10010, 1, 1), (0x10011, 2, 2), (0x1206, 4, 4), (0x120C, 6, 5), (0x1212, 8, 7)]
    rng = np.random.default_rng(5)
    for fmt, nch, layout in cases:
        from helpers import scenes
        from pyb200mix import abi
        ref, _ = scenes.make_ref_scene(1, 1 if devname == "hrtf" else 0, abi.RS_LINEAR, attrs=attrs)
        try:
            ir = (rng.standard_normal((600, nch)) * 0.05).astype(np.float32)
            slot_gain = float(np.float32(rng.uniform(0.2, 1.0)))
            slot = ref.add_convolution_slot(ir, 48000, slot_gain, fmt=fmt)
            ref.connect_send(ref.sources[0], slot)
            ref.play_all()
            ref.render(64)
            want = np.zeros((8, 25), dtype=np.float32)
            assert tap.refh_conv_gains(ref.ctx, 0, want.ctypes.data) == nch
            scale = np.zeros(32, dtype=np.float32); index = np.zeros(32, dtype=np.uint32)
            n = hz.refh_dry_ambi_map(ref.dev, scale.ctypes.data, index.ctypes.data)
            got = np.full((8, 25), 7.0, dtype=np.float32)
            rc = prod.b200mix_convolution_gains(layout, int(hz.refh_device_render_mode(ref.dev) == 1), slot_gain, n,
                                                scale.ctypes.data, index.ctypes.data, got.ctypes.data, 25)
            assert rc == nch, (devname, layout, rc)
            assert np.array_equal(got[:nch].view(np.uint32), want[:nch].view(np.uint32)), (devname, layout, got[:nch], want[:nch])
            assert np.abs(want[:nch]).max() > 0
        finally:
            ref.close()


class BuiltinDecoder(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("ambi_order", C.c_uint32), ("is_2d", C.c_uint32),
                ("dry_channels", C.c_uint32), ("real_channels", C.c_uint32), ("dual_band", C.c_uint32),
                ("map_scale", C.c_float * 5), ("map_index", C.c_uint32 * 5), ("gains_hf", C.c_float * 40),
                ("gains_lf", C.c_float * 40), ("xover_coeff", C.c_float)]


@pytest.mark.parametrize("layout,fmt", [(0, 0x1500), (1, 0x1501), (2, 0x1503), (3, 0x1504), (4, 0x1505), (5, 0x1506)])
def test_builtin_decoders_bit_exact(layout, fmt):
    """b200mix_builtin_decoder against a live reference device of that output format: channel
    counts, the Dry AmbiMap, the BFormatDec gain matrices and the crossover coefficient."""
    prod, hz = _bind()
    prod.b200mix_builtin_decoder.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(BuiltinDecoder)]
    hz.refh_device_ambi.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    hz.refh_device_ambi.restype = None
    ref = refal.RefDevice({refal.ALC_FORMAT_CHANNELS_SOFT: fmt, refal.ALC_HRTF_SOFT: 0})
    try:
        out = BuiltinDecoder()
        out.struct_size = C.sizeof(out)
        assert prod.b200mix_builtin_decoder(layout, 1, ref.desc.sample_rate, C.byref(out)) == 0
        assert (out.dry_channels, out.real_channels) == (ref.desc.dry_channels, ref.desc.real_channels)
        order, is2d, xover = C.c_uint32(0), C.c_uint32(0), C.c_float(0.0)
        hz.refh_device_ambi(ref.dev, C.byref(order), C.byref(is2d), C.byref(xover))
        assert (out.ambi_order, out.is_2d) == (order.value, is2d.value)
        scale = np.zeros(32, dtype=np.float32); index = np.zeros(32, dtype=np.uint32)
        n = hz.refh_dry_ambi_map(ref.dev, scale.ctypes.data, index.ctypes.data)
        assert list(out.map_index)[:n] == list(index[:n])
        assert np.array_equal(np.array(list(out.map_scale)[:n], dtype=np.float32).view(np.uint32), scale[:n].view(np.uint32))
        hfm, lfm, xo = ref.ambi_decoder()
        cnt = out.dry_channels * out.real_channels
        got_hf = np.array(list(out.gains_hf)[:cnt], dtype=np.float32).reshape(hfm.shape)
        assert np.array_equal(got_hf.view(np.uint32), hfm.view(np.uint32)), (got_hf, hfm)
        assert bool(out.dual_band) == (lfm is not None)
        if lfm is not None:
            got_lf = np.array(list(out.gains_lf)[:cnt], dtype=np.float32).reshape(lfm.shape)
            assert np.array_equal(got_lf.view(np.uint32), lfm.view(np.uint32))
            assert np.float32(out.xover_coeff).view(np.uint32) == np.float32(xo).view(np.uint32)
    finally:
        ref.close()
