"""Host-side reverb parameter stage of the product (b200mix_reverb_params_from_efx, no GPU
involved) against the live reference: EAX reverb properties go through the reference's AL layer
and ReverbState::update, the resulting pipeline is read back with oracle/ref_reverb_tap.cpp and
every field — line lengths, taps, filters, T60 coefficients, modulation, fade length, upmix
constants and the 8 output gain rows — must equal the helper's bit for bit."""
import ctypes as C

import numpy as np
import pytest

from helpers import mixlib, refal, scenes
from pyb200mix import abi

pytestmark = pytest.mark.ref

# AL_EAXREVERB_* ids (include/AL/efx.h) -> b200mix_efx_reverb field, default, range
PROPS = [
    (0x0001, "density", 1.0, (0.0, 1.0)), (0x0002, "diffusion", 1.0, (0.0, 1.0)),
    (0x0003, "gain", 0.32, (0.0, 1.0)), (0x0004, "gain_hf", 0.89, (0.0, 1.0)),
    (0x0005, "gain_lf", 1.0, (0.0, 1.0)), (0x0006, "decay_time", 1.49, (0.1, 20.0)),
    (0x0007, "decay_hf_ratio", 0.83, (0.1, 2.0)), (0x0008, "decay_lf_ratio", 1.0, (0.1, 2.0)),
    (0x0009, "reflections_gain", 0.05, (0.0, 3.16)), (0x000A, "reflections_delay", 0.007, (0.0, 0.3)),
    (0x000C, "late_reverb_gain", 1.26, (0.0, 10.0)), (0x000D, "late_reverb_delay", 0.011, (0.0, 0.1)),
    (0x000F, "echo_time", 0.25, (0.075, 0.25)), (0x0010, "echo_depth", 0.0, (0.0, 1.0)),
    (0x0011, "modulation_time", 0.25, (0.04, 4.0)), (0x0012, "modulation_depth", 0.0, (0.0, 1.0)),
    (0x0013, "air_absorption_gain_hf", 0.994, (0.892, 1.0)), (0x0014, "hf_reference", 5000.0, (1000.0, 20000.0)),
    (0x0015, "lf_reference", 250.0, (20.0, 1000.0)), (0x0016, "room_rolloff_factor", 0.0, (0.0, 10.0)),
]
AL_REFLECTIONS_PAN, AL_LATE_REVERB_PAN, AL_DECAY_HFLIMIT = 0x000B, 0x000E, 0x0017

DEVICES = {
    "hrtf": {refal.ALC_HRTF_SOFT: 1},
    "stereo": {refal.ALC_HRTF_SOFT: 0},
    "ambi2": {refal.ALC_FORMAT_CHANNELS_SOFT: refal.ALC_BFORMAT3D_SOFT, refal.ALC_AMBISONIC_ORDER_SOFT: 2,
              refal.ALC_AMBISONIC_LAYOUT_SOFT: refal.ALC_ACN_SOFT, refal.ALC_AMBISONIC_SCALING_SOFT: refal.ALC_N3D_SOFT},
    "ambi3": {refal.ALC_FORMAT_CHANNELS_SOFT: refal.ALC_BFORMAT3D_SOFT, refal.ALC_AMBISONIC_ORDER_SOFT: 3,
              refal.ALC_AMBISONIC_LAYOUT_SOFT: refal.ALC_ACN_SOFT, refal.ALC_AMBISONIC_SCALING_SOFT: refal.ALC_N3D_SOFT},
}


def _prop_sets(rng, count):
    yield {}, None, None, 1                       # the EFX defaults
    for k in range(count):
        vals = {}
        for pid, name, dflt, (lo, hi) in PROPS:
            if rng.random() < 0.7:
                vals[name] = float(np.float32(rng.uniform(lo, hi)))
        if k % 3 == 0:
            vals["modulation_time"] = float(np.float32(rng.uniform(0.04, 0.25)))
            vals["modulation_depth"] = float(np.float32(rng.uniform(0.1, 1.0)))
        pans = []
        for _ in range(2):
            v = rng.standard_normal(3) * (0.4 if k % 2 else 1.5)
            pans.append([float(np.float32(x)) for x in v])
        yield vals, pans[0], pans[1], int(k % 4 != 1)


@pytest.mark.parametrize("devname", sorted(DEVICES))
def test_reverb_params_from_efx_bit_exact(devname):
    prod = mixlib.product().lib
    prod.b200mix_reverb_params_from_efx.argtypes = [C.POINTER(abi.EfxReverb), C.POINTER(abi.ReverbTarget),
                                                    C.POINTER(abi.ReverbParams), C.c_void_p]
    _, hz = refal.libs()
    hz.refh_device_ambi.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    hz.refh_device_ambi.restype = None
    hz.refh_dry_ambi_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(81)
    for vals, rpan, lpan, hflimit in _prop_sets(rng, 10):
        slot_gain = float(np.float32(rng.uniform(0.2, 1.0)))
        ref, pcms = scenes.make_ref_scene(1, 1 if devname == "hrtf" else 0, abi.RS_LINEAR, attrs=DEVICES[devname])
        try:
            al = ref.al
            al.alEffectfv.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_float)]
            al.alEffecti.argtypes = [C.c_uint, C.c_int, C.c_int]
            by_name = {name: pid for pid, name, _, _ in PROPS}
            slot = ref.add_reverb_slot(props={by_name[k]: v for k, v in vals.items()}, slot_gain=slot_gain)
            e = ref._slot_effect[slot]
            if rpan is not None:
                al.alEffectfv(e, AL_REFLECTIONS_PAN, (C.c_float * 3)(*rpan))
                al.alEffectfv(e, AL_LATE_REVERB_PAN, (C.c_float * 3)(*lpan))
            al.alEffecti(e, AL_DECAY_HFLIMIT, hflimit)
            al.alAuxiliaryEffectSloti(slot, refal.AL_EFFECTSLOT_EFFECT, e)
            assert al.alGetError() == 0
            ref.connect_send(ref.sources[0], slot)
            ref.play_all()
            ref.render(64)
            want, want_gains, _ = ref.reverb_params(0)

            props = abi.EfxReverb()
            props.struct_size = C.sizeof(props)
            for pid, name, dflt, _ in PROPS:
                setattr(props, name, vals.get(name, dflt))
            for i in range(3):
                props.reflections_pan[i] = (rpan or [0.0, 0.0, 0.0])[i]
                props.late_reverb_pan[i] = (lpan or [0.0, 0.0, 0.0])[i]
            props.decay_hf_limit = hflimit
            order, is2d, xover = C.c_uint32(0), C.c_uint32(0), C.c_float(0.0)
            hz.refh_device_ambi(ref.dev, C.byref(order), C.byref(is2d), C.byref(xover))
            scale = np.zeros(32, dtype=np.float32)
            index = np.zeros(32, dtype=np.uint32)
            n = hz.refh_dry_ambi_map(ref.dev, scale.ctypes.data, index.ctypes.data)
            tgt = abi.ReverbTarget(C.sizeof(abi.ReverbTarget), ref.desc.sample_rate, order.value, is2d.value,
                                   xover.value, slot_gain, 1.0, n, scale.ctypes.data, index.ctypes.data)
            got = abi.ReverbParams()
            got_gains = np.zeros((8, n), dtype=np.float32)
            rc = prod.b200mix_reverb_params_from_efx(C.byref(props), C.byref(tgt), C.byref(got),
                                                     got_gains.ctypes.data)
            assert rc == 0, rc
            for fname, _ in abi.ReverbParams._fields_:
                a = np.frombuffer(bytes(getattr(want, fname)) if not isinstance(getattr(want, fname), (int, float))
                                  else np.array([getattr(want, fname)]).tobytes(), dtype=np.uint8)
                b = np.frombuffer(bytes(getattr(got, fname)) if not isinstance(getattr(got, fname), (int, float))
                                  else np.array([getattr(got, fname)]).tobytes(), dtype=np.uint8)
                assert np.array_equal(a, b), (devname, fname, vals, getattr(want, fname), getattr(got, fname))
            assert bytes(want) == bytes(got)
            assert np.array_equal(want_gains.view(np.uint32), got_gains.view(np.uint32)), (devname, vals)
            assert np.abs(want_gains).max() > 0
        finally:
            ref.close()


def test_full_update_test_matches_the_references_rule():
    prod = mixlib.product().lib
    prod.b200mix_reverb_full_update_needed.argtypes = [C.POINTER(abi.EfxReverb), C.POINTER(abi.EfxReverb)]

    def mk(**kw):
        p = abi.EfxReverb()
        p.struct_size = C.sizeof(p)
        for _, name, dflt, _ in PROPS:
            setattr(p, name, kw.get(name, dflt))
        p.decay_hf_limit = kw.get("decay_hf_limit", 1)
        return p
    base = mk()
    assert prod.b200mix_reverb_full_update_needed(None, C.byref(base)) == 1
    assert prod.b200mix_reverb_full_update_needed(C.byref(base), C.byref(mk())) == 0
    # gains, delays and pans are applied in place; density, diffusion, decay, modulation and the
    # reference frequencies switch pipelines (alc/effects/reverb.cpp:1243-1262)
    for name in ("gain", "gain_hf", "gain_lf", "reflections_gain", "reflections_delay", "late_reverb_gain",
                 "late_reverb_delay", "echo_time", "room_rolloff_factor"):
        assert prod.b200mix_reverb_full_update_needed(C.byref(base), C.byref(mk(**{name: 0.09}))) == 0, name
    for name, v in (("density", 0.5), ("diffusion", 0.5), ("decay_time", 2.0), ("decay_hf_ratio", 0.5),
                    ("decay_lf_ratio", 0.5), ("modulation_time", 1.0), ("modulation_depth", 0.5),
                    ("hf_reference", 4000.0), ("lf_reference", 200.0)):
        assert prod.b200mix_reverb_full_update_needed(C.byref(base), C.byref(mk(**{name: v}))) == 1, name
