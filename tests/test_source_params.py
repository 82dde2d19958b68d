"""Host-side source parameter stage of the product (b200mix_calc_source_params and the panning /
filter helpers it feeds; no GPU involved) against live voices of the compiled reference: random
listeners and sources (all distance models, cones, doppler, air absorption, radius, direct and
send filters, a reverb send with decay) are set up through the AL API, the reference's ALU computes
its voice parameters, and the helpers must reproduce them bit for bit from the same properties:
resampler step, HRIR coefficients / delays / gain or dry panning gains, send panning gains and the
direct / send filter coefficient sets."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import mixlib, refal, scenes
from pyb200mix import abi, scene
from pyb200mix.abi import (ListenerParams, ListenerProps, SourceSend, SourceProps, SourceResult, MixMap, VoiceEnv,
                            ChannelSetup, BFormatSetup)

pytestmark = pytest.mark.ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MHR = os.path.join(ROOT, "openal-soft_b200", "data", "Default HRTF.mhr")
MODELS = [0, 0xD001, 0xD002, 0xD003, 0xD004, 0xD005, 0xD006]     # AL_NONE, AL_INVERSE_DISTANCE, ...














def test_listener_params_match_calc_context_params():
    """b200mix_calc_listener_params against the ContextParams the reference derives from the same
    listener / context properties set through the AL API."""
    prod = mixlib.product().lib
    prod.b200mix_calc_listener_params.argtypes = [C.POINTER(ListenerProps), C.POINTER(ListenerParams)]
    _, hz = refal.libs()
    hz.refh_listener_params.argtypes = [C.c_void_p, C.POINTER(ListenerParams)]
    hz.refh_listener_params.restype = None
    for seed in range(8):
        rng = np.random.default_rng(700 + seed)
        ref, _ = scenes.make_ref_scene(1, 0, abi.RS_LINEAR)
        try:
            al = ref.al
            al.alListenerfv.argtypes = [C.c_int, C.POINTER(C.c_float)]
            al.alListener3f.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
            al.alListenerf.argtypes = [C.c_int, C.c_float]
            al.alDopplerFactor.argtypes = [C.c_float]
            al.alSpeedOfSound.argtypes = [C.c_float]
            al.alEnable.argtypes = [C.c_int]
            lp = ListenerProps()
            lp.struct_size = C.sizeof(lp)
            pos, vel = _f3(rng, 3.0), _f3(rng, 8.0)
            at = rng.standard_normal(3)
            up = rng.standard_normal(3) if seed % 2 else np.cross(np.cross(at, rng.standard_normal(3)), at)
            ori = [float(np.float32(x)) for x in np.concatenate([at, up])]
            for i in range(3):
                lp.position[i], lp.velocity[i], lp.orient_at[i], lp.orient_up[i] = pos[i], vel[i], ori[i], ori[3 + i]
            lp.gain = float(rng.uniform(0.1, 2.0)); lp.gain_boost = 1.0
            lp.meters_per_unit = float(rng.uniform(0.1, 4.0)); lp.air_absorption_gain_hf = 0.99426
            lp.doppler_factor = float(rng.uniform(0.0, 3.0)); lp.doppler_velocity = 1.0
            lp.speed_of_sound = float(rng.uniform(50.0, 900.0))
            lp.source_distance_model = seed % 2
            model = int(rng.integers(0, 7)); lp.distance_model = model
            al.alListener3f(0x1004, *pos); al.alListener3f(0x1006, *vel)
            al.alListenerfv(0x100F, (C.c_float * 6)(*ori))
            al.alListenerf(refal.AL_GAIN, lp.gain); al.alListenerf(0x20004, lp.meters_per_unit)
            al.alDopplerFactor(lp.doppler_factor); al.alSpeedOfSound(lp.speed_of_sound)
            al.alDistanceModel(MODELS[model])
            if seed % 2:
                al.alEnable(0x200)
            assert al.alGetError() == 0
            ref.play_all()
            ref.render(16)
            want = ListenerParams()
            hz.refh_listener_params(ref.ctx, C.byref(want))
            got = ListenerParams()
            assert prod.b200mix_calc_listener_params(C.byref(lp), C.byref(got)) == 0
            assert bytes(want) == bytes(got), (seed, list(want.matrix), list(got.matrix))
        finally:
            ref.close()


def _f3(rng, s):
    return [float(np.float32(x)) for x in rng.standard_normal(3) * s]


def _build_scene(rng, devname, V):
    attrs = {"hrtf": {refal.ALC_HRTF_SOFT: 1}, "stereo": {refal.ALC_HRTF_SOFT: 0},
             "ambi3": {refal.ALC_FORMAT_CHANNELS_SOFT: refal.ALC_BFORMAT3D_SOFT, refal.ALC_AMBISONIC_ORDER_SOFT: 3,
                       refal.ALC_AMBISONIC_LAYOUT_SOFT: refal.ALC_ACN_SOFT,
                       refal.ALC_AMBISONIC_SCALING_SOFT: refal.ALC_N3D_SOFT}}[devname]
    ref, pcms = scenes.make_ref_scene(V, 1 if devname == "hrtf" else 0, abi.RS_LINEAR, attrs=attrs)
    al = ref.al
    al.alListenerfv.argtypes = [C.c_int, C.POINTER(C.c_float)]
    al.alListener3f.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
    al.alListenerf.argtypes = [C.c_int, C.c_float]
    al.alDopplerFactor.argtypes = [C.c_float]
    al.alSpeedOfSound.argtypes = [C.c_float]
    al.alEnable.argtypes = [C.c_int]
    # listener
    lpos = _f3(rng, 2.0)
    al.alListener3f(0x1004, *lpos)
    al.alListener3f(0x1006, *_f3(rng, 5.0))
    at = rng.standard_normal(3)
    up = np.cross(np.cross(at, rng.standard_normal(3)), at)
    al.alListenerfv(0x100F, (C.c_float * 6)(*[float(x) for x in np.concatenate([at, up])]))
    al.alListenerf(refal.AL_GAIN, float(rng.uniform(0.3, 1.5)))
    al.alListenerf(0x20004, float(rng.uniform(0.2, 3.0)))              # AL_METERS_PER_UNIT
    al.alDopplerFactor(float(rng.uniform(0.0, 2.0)))
    al.alSpeedOfSound(float(rng.uniform(100.0, 600.0)))
    al.alDistanceModel(int(rng.choice(MODELS)))
    if rng.random() < 0.5:
        al.alEnable(0x200)                                             # AL_SOURCE_DISTANCE_MODEL
    slot = ref.add_reverb_slot(props={0x0006: float(rng.uniform(0.5, 5.0)), 0x0013: float(rng.uniform(0.9, 1.0)),
                                      0x0016: float(rng.uniform(0.0, 2.0))})
    for k, src in enumerate(ref.sources):
        r = rng.uniform(0.05, 30.0)
        d = rng.standard_normal(3)
        d /= np.linalg.norm(d)
        al.alSource3f(src, 0x1004, *[float(np.float32(x)) for x in d * r])
        al.alSource3f(src, 0x1006, *_f3(rng, 20.0))
        if k % 3:
            al.alSource3f(src, 0x1005, *_f3(rng, 1.0))
            inner = float(rng.uniform(0.0, 300.0))
            al.alSourcef(src, 0x1001, inner)
            al.alSourcef(src, 0x1002, float(rng.uniform(inner, 360.0)))
            al.alSourcef(src, 0x1022, float(rng.uniform(0.0, 1.0)))
            al.alSourcef(src, 0x20009, float(rng.uniform(0.0, 1.0)))
        ref_d = float(rng.uniform(0.1, 5.0))
        al.alSourcef(src, 0x1020, ref_d)
        al.alSourcef(src, 0x1023, float(rng.uniform(0.05, 60.0)))
        al.alSourcef(src, 0x1021, float(rng.uniform(0.0, 3.0)))
        al.alSourcef(src, 0x100D, float(rng.uniform(0.0, 0.2)))
        al.alSourcef(src, 0x100E, float(rng.uniform(0.5, 1.0)))
        al.alSourcef(src, refal.AL_GAIN, float(rng.uniform(0.1, 2.0)))
        al.alSourcef(src, refal.AL_PITCH, float(rng.uniform(0.3, 3.0)))
        al.alSourcei(src, 0x202, int(k % 4 == 0))
        if k % 7 == 3:
            # a source sitting exactly on the listener: the reference's no-distance panning path
            al.alSource3f(src, 0x1004, *([0.0, 0.0, 0.0] if k % 4 == 0 else lpos))
        al.alSourcei(src, 0xD000, int(rng.choice(MODELS)))             # AL_DISTANCE_MODEL (per source)
        al.alSourcef(src, 0x20007, float(rng.uniform(0.0, 10.0)))
        al.alSourcef(src, 0x20008, float(rng.uniform(0.0, 1.0)))
        al.alSourcef(src, 0xC000, float(rng.uniform(0.0, 1.0)))        # AL_DOPPLER_FACTOR (source)
        if k % 2:
            al.alSourcef(src, 0x1031, float(rng.uniform(0.0, 2.0 * r)))
        al.alSourcei(src, 0x2000A, int(rng.integers(0, 2)))
        al.alSourcei(src, 0x2000B, int(rng.integers(0, 2)))
        al.alSourcei(src, 0x2000C, int(rng.integers(0, 2)))
        if k % 2 == 0:
            ref.set_direct_filter(src, ref.make_filter(float(rng.uniform(0.2, 1.0)), float(rng.uniform(0.1, 1.0)),
                                                      float(rng.uniform(0.1, 1.0)) if k % 4 == 0 else None))
        sf = refal.AL_FILTER_NULL if k % 3 == 0 else ref.make_filter(float(rng.uniform(0.2, 1.0)),
                                                                     float(rng.uniform(0.1, 1.0)))
        if k % 5 != 4:
            ref.connect_send(src, slot, 0, sf)
    err = al.alGetError()
    assert err == 0, hex(err)
    return ref


@pytest.mark.skipif(not os.path.exists(MHR), reason="HRTF data set not staged (run build())")
@pytest.mark.parametrize("devname", ["hrtf", "stereo", "ambi3"])
def test_source_params_reproduce_the_references_voices(devname):
    prod = mixlib.product().lib
    prod.b200mix_calc_source_params.argtypes = [C.POINTER(SourceProps), C.POINTER(ListenerParams), C.c_uint32,
                                                C.c_uint32, C.c_uint32, C.POINTER(SourceResult)]
    prod.b200mix_pairwise_azimuth.argtypes = [C.c_void_p, C.c_void_p]
    prod.b200mix_ambi_coeffs.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
    prod.b200mix_pan_gains.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_uint32]
    prod.b200mix_biquad_coeffs.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_void_p]
    prod.b200mix_hrtf_load.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    prod.b200mix_hrtf_get_coeffs.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                             C.POINTER(C.c_uint32)]
    prod.b200mix_hrtf_free.argtypes = [C.c_void_p]
    _, hz = refal.libs()
    hz.refh_listener_params.argtypes = [C.c_void_p, C.POINTER(ListenerParams)]
    hz.refh_listener_params.restype = None
    hz.refh_source_props.argtypes = [C.c_void_p, C.c_int, C.POINTER(SourceProps), C.POINTER(C.c_uint32)]
    hz.refh_device_render_mode.argtypes = [C.c_void_p]
    hz.refh_dry_ambi_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    hz.refh_slot_ambi_map.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    hrtf = C.c_void_p()
    data = open(MHR, "rb").read()
    assert prod.b200mix_hrtf_load(data, len(data), C.byref(hrtf)) == 0

    V = 24
    checked = dict(cone=0, doppler=0, absorb=0, filt=0, sends=0)
    for seed in range(3):
        rng = np.random.default_rng(900 + seed)
        ref = _build_scene(rng, devname, V)
        try:
            ref.play_all()
            ref.render(64)
            nslots, wet = ref.slot_info()
            n, params, coeffs, dry, send, _ = ref.snapshot(wet_channels=wet[0])
            assert n >= V
            ents, _ = ref.voice_filters(V)
            filt = {(v, p): (a, lp, hp) for v, p, a, lp, hp in ents}
            lis = ListenerParams()
            hz.refh_listener_params(ref.ctx, C.byref(lis))
            mode = hz.refh_device_render_mode(ref.dev)          # 0 normal, 1 pairwise, 2 hrtf
            assert (mode == 2) == (devname == "hrtf")
            dscale = np.zeros(32, dtype=np.float32); dindex = np.zeros(32, dtype=np.uint32)
            nd = hz.refh_dry_ambi_map(ref.dev, dscale.ctypes.data, dindex.ctypes.data)
            wscale = np.zeros(32, dtype=np.float32); windex = np.zeros(32, dtype=np.uint32)
            nw = hz.refh_slot_ambi_map(ref.ctx, 0, wscale.ctypes.data, windex.ctypes.data)
            assert nw == wet[0]
            rate = ref.desc.sample_rate
            ns = ref.desc.num_sends
            for k in range(V):
                sp = SourceProps()
                brate = C.c_uint32(0)
                assert hz.refh_source_props(ref.ctx, k, C.byref(sp), C.byref(brate)) == 0
                res = SourceResult()
                assert prod.b200mix_calc_source_params(C.byref(sp), C.byref(lis), ns, brate.value, rate,
                                                       C.byref(res)) == 0
                assert res.step == params[k].step, (devname, seed, k, res.step, params[k].step)
                if not res.distance > 1e-6:
                    continue            # the no-distance path is covered by test_calc_voice_single_call
                pos = np.array(list(res.pos), dtype=np.float32)
                co = np.zeros(25, dtype=np.float32)
                if mode == 2:
                    out = np.zeros((ref.desc.ir_size, 2), dtype=np.float32)
                    dl = (C.c_uint32 * 2)()
                    assert prod.b200mix_hrtf_get_coeffs(hrtf, res.hrtf_elevation, res.hrtf_azimuth, res.distance,
                                                        res.spread, out.ctypes.data, dl) == 0
                    assert np.array_equal(out.view(np.uint32), coeffs[k].view(np.uint32)), (devname, seed, k)
                    assert list(dl) == list(params[k].hrtf_delay)
                    assert np.float32(res.dry_gain).view(np.uint32) == np.float32(params[k].hrtf_gain).view(np.uint32)
                else:
                    ppos = pos.copy()
                    if mode == 1:
                        assert prod.b200mix_pairwise_azimuth(pos.ctypes.data, ppos.ctypes.data) == 0
                    assert prod.b200mix_ambi_coeffs(ppos.ctypes.data, res.spread, co.ctypes.data) == 0
                    g = np.zeros(nd, dtype=np.float32)
                    assert prod.b200mix_pan_gains(nd, dscale.ctypes.data, dindex.ctypes.data, co.ctypes.data,
                                                  res.dry_gain, g.ctypes.data, nd) == 0
                    assert np.array_equal(g.view(np.uint32), dry[k].view(np.uint32)), (devname, seed, k, g, dry[k])
                # sends: always the unscaled direction (alc/alu.cpp:1218-1225,1356-1361 use the same coeffs as
                # the dry path; with pairwise scaling they share the scaled position)
                wpos = pos.copy()
                if mode == 1:
                    assert prod.b200mix_pairwise_azimuth(pos.ctypes.data, wpos.ctypes.data) == 0
                assert prod.b200mix_ambi_coeffs(wpos.ctypes.data, res.spread, co.ctypes.data) == 0
                if sp.sends[0].active:
                    g = np.zeros(nw, dtype=np.float32)
                    assert prod.b200mix_pan_gains(nw, wscale.ctypes.data, windex.ctypes.data, co.ctypes.data,
                                                  res.wet_gain[0], g.ctypes.data, nw) == 0
                    assert np.array_equal(g.view(np.uint32), send[k][0].view(np.uint32)), (devname, seed, k)
                    checked["sends"] += 1
                # filters (alc/alu.cpp:1619-1656)
                for path, (ghf, glf, hfref, lfref) in enumerate(
                        [(res.dry_gain_hf, res.dry_gain_lf, sp.direct.hf_reference, sp.direct.lf_reference)]
                        + [(res.wet_gain_hf[s], res.wet_gain_lf[s], sp.sends[s].hf_reference, sp.sends[s].lf_reference)
                           for s in range(ns)]):
                    act, lp, hp = filt[(k, path)]
                    inv = np.float32(1.0) / np.float32(rate)
                    a = np.zeros(5, dtype=np.float32); b = np.zeros(5, dtype=np.float32)
                    assert prod.b200mix_biquad_coeffs(0, float(np.float32(hfref) * inv), ghf, 1.0, a.ctypes.data) == 0
                    assert prod.b200mix_biquad_coeffs(1, float(np.float32(lfref) * inv), glf, 1.0, b.ctypes.data) == 0
                    assert bool(act) == (ghf != 1.0 or glf != 1.0), (devname, seed, k, path)
                    assert np.array_equal(a.view(np.uint32), lp.view(np.uint32)), (devname, seed, k, path, a, lp)
                    assert np.array_equal(b.view(np.uint32), hp.view(np.uint32)), (devname, seed, k, path)
                    checked["filt"] += int(bool(act))
                checked["cone"] += int(sp.inner_angle < 360.0 and any(sp.direction))
                checked["doppler"] += int(sp.doppler_factor * lis.doppler_factor > 0)
                checked["absorb"] += int(res.distance > sp.ref_distance)
        finally:
            ref.close()
    prod.b200mix_hrtf_free(hrtf)
    assert all(v > 5 for v in checked.values()), checked






@pytest.mark.skipif(not os.path.exists(MHR), reason="HRTF data set not staged (run build())")
@pytest.mark.parametrize("devname", ["hrtf", "stereo", "ambi3"])
def test_calc_voice_single_call(devname):
    """b200mix_calc_voice (the whole CalcVoiceParams of a point source in one call) against the same
    live voices: the structs it fills are what b200mix_voices_update(_dirs) / _filters take."""
    prod = mixlib.product().lib
    prod.b200mix_calc_voice.argtypes = [C.POINTER(SourceProps), C.POINTER(ListenerParams), C.POINTER(VoiceEnv),
                                        C.c_uint32, C.POINTER(abi.VoiceParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
    prod.b200mix_hrtf_load.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    prod.b200mix_hrtf_get_coeffs.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                             C.POINTER(C.c_uint32)]
    prod.b200mix_hrtf_free.argtypes = [C.c_void_p]
    _, hz = refal.libs()
    hz.refh_listener_params.argtypes = [C.c_void_p, C.POINTER(ListenerParams)]
    hz.refh_listener_params.restype = None
    hz.refh_source_props.argtypes = [C.c_void_p, C.c_int, C.POINTER(SourceProps), C.POINTER(C.c_uint32)]
    hz.refh_device_render_mode.argtypes = [C.c_void_p]
    hz.refh_dry_ambi_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    hz.refh_slot_ambi_map.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    hrtf = C.c_void_p()
    data = open(MHR, "rb").read()
    assert prod.b200mix_hrtf_load(data, len(data), C.byref(hrtf)) == 0
    V = 16
    rng = np.random.default_rng(950)
    ref = _build_scene(rng, devname, V)
    try:
        ref.play_all()
        ref.render(64)
        nslots, wet = ref.slot_info()
        n, params, coeffs, dry, send, _ = ref.snapshot(wet_channels=wet[0])
        ents, _ = ref.voice_filters(V)
        filt = {(v, p): (a, lp, hp) for v, p, a, lp, hp in ents}
        lis = ListenerParams()
        hz.refh_listener_params(ref.ctx, C.byref(lis))
        dscale = np.zeros(32, dtype=np.float32); dindex = np.zeros(32, dtype=np.uint32)
        nd = hz.refh_dry_ambi_map(ref.dev, dscale.ctypes.data, dindex.ctypes.data)
        wscale = np.zeros(32, dtype=np.float32); windex = np.zeros(32, dtype=np.uint32)
        nw = hz.refh_slot_ambi_map(ref.ctx, 0, wscale.ctypes.data, windex.ctypes.data)
        ns = ref.desc.num_sends
        env = VoiceEnv()
        env.struct_size = C.sizeof(env)
        env.device_rate, env.num_sends = ref.desc.sample_rate, ns
        env.render_mode = hz.refh_device_render_mode(ref.dev)
        env.wet_stride = nw
        env.dry = MixMap(nd, dscale.ctypes.data, dindex.ctypes.data)
        env.wet[0] = MixMap(nw, wscale.ctypes.data, windex.ctypes.data)
        zero_dist = 0
        for k in range(V):
            sp = SourceProps()
            brate = C.c_uint32(0)
            assert hz.refh_source_props(ref.ctx, k, C.byref(sp), C.byref(brate)) == 0
            vp = abi.VoiceParams()
            vp.voice = k
            d4 = np.zeros(4, dtype=np.float32)
            dg = np.full(nd, 9.0, dtype=np.float32)
            sg = np.full((ns, nw), 9.0, dtype=np.float32)
            fl = (abi.VoiceFilter * (1 + abi.MAX_SENDS))()
            rc = prod.b200mix_calc_voice(C.byref(sp), C.byref(lis), C.byref(env), brate.value, C.byref(vp),
                                         d4.ctypes.data, dg.ctypes.data, sg.ctypes.data, fl)
            assert rc == 0, rc
            assert vp.step == params[k].step
            zero_dist += int(np.isinf(d4[2]) or (env.render_mode != 2 and k % 7 == 3))
            if env.render_mode == 2:
                assert vp.flags & abi.VF_HRTF
                out = np.zeros((ref.desc.ir_size, 2), dtype=np.float32)
                dl = (C.c_uint32 * 2)()
                assert prod.b200mix_hrtf_get_coeffs(hrtf, d4[0], d4[1], d4[2], d4[3], out.ctypes.data, dl) == 0
                assert np.array_equal(out.view(np.uint32), coeffs[k].view(np.uint32))
                assert list(dl) == list(params[k].hrtf_delay)
                assert np.float32(vp.hrtf_gain).view(np.uint32) == np.float32(params[k].hrtf_gain).view(np.uint32)
            else:
                assert not (vp.flags & abi.VF_HRTF)
                assert np.array_equal(dg.view(np.uint32), dry[k].view(np.uint32)), (k, dg, dry[k])
            want_send = send[k][0] if sp.sends[0].active else np.zeros(nw, dtype=np.float32)
            assert np.array_equal(sg[0].view(np.uint32), want_send.view(np.uint32)), (k, sg[0], want_send)
            for path in range(1 + ns):
                act, lp, hp = filt[(k, path)]
                f = fl[path]
                assert (f.voice, f.path, bool(f.active)) == (k, path, bool(act))
                assert np.array_equal(np.array(list(f.lowpass), dtype=np.float32).view(np.uint32), lp.view(np.uint32))
                assert np.array_equal(np.array(list(f.highpass), dtype=np.float32).view(np.uint32), hp.view(np.uint32))
        assert zero_dist >= 2, zero_dist
    finally:
        ref.close()
        prod.b200mix_hrtf_free(hrtf)




LAYOUTS = {  # name: (AL format, channels, b200mix_channel_layout)
    "stereo": (0x1103, 2, 2), "rear": (0x1208, 2, 3), "quad": (0x1205, 4, 4),
    "x51": (0x120B, 6, 5), "x61": (0x120E, 7, 6), "x71": (0x1211, 8, 7)}


@pytest.mark.skipif(not os.path.exists(MHR), reason="HRTF data set not staged (run build())")
@pytest.mark.parametrize("devname", ["hrtf", "stereo", "ambi3"])
def test_calc_voice_channels_for_unspatialized_multichannel_sources(devname):
    """b200mix_calc_voice_channels against live stereo / rear / quad / 5.1 / 6.1 / 7.1 sources of the
    reference (CalcNonAttnVoiceParams + the no-distance panning): every mixing channel's HRIR pair
    and gain or dry panning gains, send gains, the shared step and filters, bit for bit."""
    prod = mixlib.product().lib
    prod.b200mix_calc_voice_channels.argtypes = [C.POINTER(SourceProps), C.POINTER(ListenerParams), C.POINTER(VoiceEnv),
                                                 C.c_uint32, C.POINTER(ChannelSetup), C.POINTER(C.c_uint32)] + [C.c_void_p] * 5
    prod.b200mix_hrtf_load.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    prod.b200mix_hrtf_get_coeffs.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                             C.POINTER(C.c_uint32)]
    prod.b200mix_hrtf_free.argtypes = [C.c_void_p]
    _, hz = refal.libs()
    hz.refh_listener_params.argtypes = [C.c_void_p, C.POINTER(ListenerParams)]
    hz.refh_listener_params.restype = None
    hz.refh_source_props.argtypes = [C.c_void_p, C.c_int, C.POINTER(SourceProps), C.POINTER(C.c_uint32)]
    hz.refh_device_render_mode.argtypes = [C.c_void_p]
    hz.refh_dry_ambi_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    hz.refh_slot_ambi_map.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    hrtf = C.c_void_p()
    data = open(MHR, "rb").read()
    assert prod.b200mix_hrtf_load(data, len(data), C.byref(hrtf)) == 0
    attrs = {"hrtf": {refal.ALC_HRTF_SOFT: 1}, "stereo": {refal.ALC_HRTF_SOFT: 0},
             "ambi3": {refal.ALC_FORMAT_CHANNELS_SOFT: refal.ALC_BFORMAT3D_SOFT, refal.ALC_AMBISONIC_ORDER_SOFT: 3,
                       refal.ALC_AMBISONIC_LAYOUT_SOFT: refal.ALC_ACN_SOFT,
                       refal.ALC_AMBISONIC_SCALING_SOFT: refal.ALC_N3D_SOFT}}[devname]
    a2 = dict(attrs)
    a2[refal.ALC_STEREO_SOURCES] = 16
    rng = np.random.default_rng(990)
    ref, _ = scenes.make_ref_scene(0, 1 if devname == "hrtf" else 0, abi.RS_LINEAR, attrs=a2, max_sources=1)
    try:
        al = ref.al
        al.alListenerf.argtypes = [C.c_int, C.c_float]
        al.alSourcefv.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_float)]
        al.alListenerf(refal.AL_GAIN, 0.8)
        slot = ref.add_reverb_slot()
        names = list(LAYOUTS) + ["stereo"]
        angles = {}
        for k, name in enumerate(names):
            fmt, nch, _ = LAYOUTS[name]
            pcm = (rng.standard_normal((2000, nch)) * 3000).astype(np.int16)
            ref.add_voice(np.ascontiguousarray(pcm), 44100, float(rng.uniform(0.5, 2.0)), (1.0, 2.0, 3.0),
                          float(rng.uniform(0.2, 1.5)), abi.RS_LINEAR, looping=True, fmt=fmt)
            src = ref.sources[-1]
            if k == len(names) - 1:
                ang = [float(np.float32(rng.uniform(0.1, 1.5))), float(np.float32(-rng.uniform(0.1, 1.5)))]
                al.alSourcefv(src, 0x1030, (C.c_float * 2)(*ang))       # AL_STEREO_ANGLES
                angles[k] = ang
            if k % 2 == 0:
                ref.set_direct_filter(src, ref.make_filter(float(rng.uniform(0.3, 1.0)), float(rng.uniform(0.1, 1.0))))
            ref.connect_send(src, slot, 0, refal.AL_FILTER_NULL if k % 3 else
                             ref.make_filter(float(rng.uniform(0.3, 1.0)), float(rng.uniform(0.1, 1.0))))
            if k % 2:
                # spatialized multi-channel source (AL_SOURCE_SPATIALIZE_SOFT): attenuated, its channels
                # drawn toward the source direction according to the radius
                al.alSourcei(src, 0x1214, 1)
                al.alSource3f(src, 0x1004, *_f3(rng, 4.0))
                al.alSourcef(src, 0x1031, float(rng.uniform(0.0, 3.0)))
        assert al.alGetError() == 0
        ref.play_all()
        ref.render(64)
        nslots, wet = ref.slot_info()
        lis = ListenerParams()
        hz.refh_listener_params(ref.ctx, C.byref(lis))
        dscale = np.zeros(32, dtype=np.float32); dindex = np.zeros(32, dtype=np.uint32)
        nd = hz.refh_dry_ambi_map(ref.dev, dscale.ctypes.data, dindex.ctypes.data)
        wscale = np.zeros(32, dtype=np.float32); windex = np.zeros(32, dtype=np.uint32)
        nw = hz.refh_slot_ambi_map(ref.ctx, 0, wscale.ctypes.data, windex.ctypes.data)
        ns = ref.desc.num_sends
        env = VoiceEnv()
        env.struct_size = C.sizeof(env)
        env.device_rate, env.num_sends = ref.desc.sample_rate, ns
        env.render_mode = hz.refh_device_render_mode(ref.dev)
        env.wet_stride = nw
        env.dry = MixMap(nd, dscale.ctypes.data, dindex.ctypes.data)
        env.wet[0] = MixMap(nw, wscale.ctypes.data, windex.ctypes.data)
        V = len(names)
        snaps = [ref.snapshot(wet_channels=wet[0], channel=c) for c in range(8)]
        ents, _ = ref.voice_filters(V)
        filt = {(v, p): (a, lp, hp) for v, p, a, lp, hp in ents}
        for k, name in enumerate(names):
            _, nch, layout = LAYOUTS[name]
            sp = SourceProps()
            brate = C.c_uint32(0)
            assert hz.refh_source_props(ref.ctx, k, C.byref(sp), C.byref(brate)) == 0
            ang = angles.get(k, [float(np.float32(np.pi / 6)), float(np.float32(-np.pi / 6))])
            setup = ChannelSetup(C.sizeof(ChannelSetup), layout, (C.c_float * 2)(*ang), 0.0, abi.NO_SLOT, k % 2)
            step = C.c_uint32(0)
            hg = np.zeros(8, dtype=np.float32)
            dirs = np.zeros((8, 4), dtype=np.float32)
            dg = np.full((8, nd), 9.0, dtype=np.float32)
            sg = np.full((8, ns, nw), 9.0, dtype=np.float32)
            fl = (abi.VoiceFilter * (1 + abi.MAX_SENDS))()
            rc = prod.b200mix_calc_voice_channels(C.byref(sp), C.byref(lis), C.byref(env), brate.value, C.byref(setup),
                                                  C.byref(step), hg.ctypes.data, dirs.ctypes.data, dg.ctypes.data,
                                                  sg.ctypes.data, fl)
            assert rc == nch, (name, rc)
            for c in range(nch):
                n, params, coeffs, dry, send, _ = snaps[c]
                assert step.value == params[k].step, (name, step.value, params[k].step)
                if env.render_mode == 2:
                    out = np.zeros((ref.desc.ir_size, 2), dtype=np.float32)
                    dl = (C.c_uint32 * 2)()
                    if hg[c] != 0.0 or np.abs(coeffs[k]).max() > 0:
                        assert prod.b200mix_hrtf_get_coeffs(hrtf, dirs[c][0], dirs[c][1], dirs[c][2], dirs[c][3],
                                                            out.ctypes.data, dl) == 0
                        assert np.array_equal(out.view(np.uint32), coeffs[k].view(np.uint32)), (name, c)
                        assert list(dl) == list(params[k].hrtf_delay), (name, c)
                    assert np.float32(hg[c]).view(np.uint32) == np.float32(params[k].hrtf_gain).view(np.uint32), (name, c)
                else:
                    assert np.array_equal(dg[c].view(np.uint32), dry[k].view(np.uint32)), (name, c, dg[c], dry[k])
                assert np.array_equal(sg[c][0].view(np.uint32), send[k][0].view(np.uint32)), (name, c, sg[c][0], send[k][0])
            for path in range(1 + ns):
                act, lp, hp = filt[(k, path)]
                f = fl[path]
                assert bool(f.active) == bool(act), (name, path)
                assert np.array_equal(np.array(list(f.lowpass), dtype=np.float32).view(np.uint32), lp.view(np.uint32))
                assert np.array_equal(np.array(list(f.highpass), dtype=np.float32).view(np.uint32), hp.view(np.uint32))
    finally:
        ref.close()
        prod.b200mix_hrtf_free(hrtf)




@pytest.mark.parametrize("devname", ["hrtf", "stereo"])
def test_calc_voice_bformat_first_order(devname):
    """b200mix_calc_voice_bformat against live first-order B-Format sources of the reference (2D and
    3D buffers, FuMa / ACN channel order, FuMa / SN3D / N3D normalisation, arbitrary source
    orientation, head-relative or not, rotated listener): every channel's dry and send gain rows,
    step and filters, bit for bit."""
    prod = mixlib.product().lib
    prod.b200mix_calc_voice_bformat.argtypes = [C.POINTER(SourceProps), C.POINTER(ListenerParams), C.POINTER(VoiceEnv),
                                                C.c_uint32, C.POINTER(BFormatSetup), C.POINTER(C.c_uint32)] + [C.c_void_p] * 3
    _, hz = refal.libs()
    hz.refh_listener_params.argtypes = [C.c_void_p, C.POINTER(ListenerParams)]
    hz.refh_listener_params.restype = None
    hz.refh_source_props.argtypes = [C.c_void_p, C.c_int, C.POINTER(SourceProps), C.POINTER(C.c_uint32)]
    hz.refh_device_render_mode.argtypes = [C.c_void_p]
    hz.refh_device_ambi.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    hz.refh_device_ambi.restype = None
    hz.refh_dry_ambi_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    hz.refh_slot_ambi_map.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(1200)
    attrs = {"hrtf": {refal.ALC_HRTF_SOFT: 1}, "stereo": {refal.ALC_HRTF_SOFT: 0}}[devname]
    a2 = dict(attrs)
    a2[refal.ALC_STEREO_SOURCES] = 16
    ref, _ = scenes.make_ref_scene(0, 1 if devname == "hrtf" else 0, abi.RS_LINEAR, attrs=a2, max_sources=1)
    try:
        al = ref.al
        al.alListenerfv.argtypes = [C.c_int, C.POINTER(C.c_float)]
        al.alListenerf.argtypes = [C.c_int, C.c_float]
        al.alSourcefv.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_float)]
        al.alBufferi.argtypes = [C.c_uint, C.c_int, C.c_int]
        at = rng.standard_normal(3)
        up = np.cross(np.cross(at, rng.standard_normal(3)), at)
        al.alListenerfv(0x100F, (C.c_float * 6)(*[float(x) for x in np.concatenate([at, up])]))
        al.alListenerf(refal.AL_GAIN, 0.9)
        slot = ref.add_reverb_slot()
        combos = [(is2d, layout, scaling) for is2d in (0, 1) for layout in (0, 1) for scaling in (0, 1, 2)]
        for k, (is2d, layout, scaling) in enumerate(combos):
            nch = 3 if is2d else 4
            pcm = np.ascontiguousarray((rng.standard_normal((1500, nch)) * 3000).astype(np.int16))
            b = C.c_uint(0); s = C.c_uint(0)
            al.alGenBuffers(1, C.byref(b))
            al.alBufferi(b, 0x1997, layout)           # AL_AMBISONIC_LAYOUT_SOFT: AL_FUMA_SOFT / AL_ACN_SOFT
            al.alBufferi(b, 0x1998, scaling)          # AL_AMBISONIC_SCALING_SOFT: FuMa / SN3D / N3D
            al.alBufferData(b, 0x20022 if is2d else 0x20032, pcm.ctypes.data, pcm.nbytes, 32000)
            al.alGenSources(1, C.byref(s))
            al.alSourcei(s, refal.AL_BUFFER, b.value)
            al.alSourcei(s, refal.AL_LOOPING, 1)
            al.alSourcef(s, refal.AL_GAIN, float(rng.uniform(0.2, 1.2)))
            al.alSourcef(s, refal.AL_PITCH, float(rng.uniform(0.5, 1.5)))
            sat = rng.standard_normal(3)
            sup = np.cross(np.cross(sat, rng.standard_normal(3)), sat)
            al.alSourcefv(s, 0x100F, (C.c_float * 6)(*[float(x) for x in np.concatenate([sat, sup])]))
            al.alSourcei(s, 0x202, k % 2)             # AL_SOURCE_RELATIVE
            ref.buffers.append(b.value); ref.sources.append(s.value); ref._keep.append(pcm)
            if k % 3 == 0:
                ref.set_direct_filter(s.value, ref.make_filter(0.8, float(rng.uniform(0.1, 1.0))))
            ref.connect_send(s.value, slot)
        assert al.alGetError() == 0
        ref.play_all()
        ref.render(64)
        nslots, wet = ref.slot_info()
        lis = ListenerParams()
        hz.refh_listener_params(ref.ctx, C.byref(lis))
        order, is2d_dev, xover = C.c_uint32(0), C.c_uint32(0), C.c_float(0.0)
        hz.refh_device_ambi(ref.dev, C.byref(order), C.byref(is2d_dev), C.byref(xover))
        assert order.value == 1
        dscale = np.zeros(32, dtype=np.float32); dindex = np.zeros(32, dtype=np.uint32)
        nd = hz.refh_dry_ambi_map(ref.dev, dscale.ctypes.data, dindex.ctypes.data)
        wscale = np.zeros(32, dtype=np.float32); windex = np.zeros(32, dtype=np.uint32)
        nw = hz.refh_slot_ambi_map(ref.ctx, 0, wscale.ctypes.data, windex.ctypes.data)
        ns = ref.desc.num_sends
        env = VoiceEnv()
        env.struct_size = C.sizeof(env)
        env.device_rate, env.num_sends = ref.desc.sample_rate, ns
        env.render_mode = hz.refh_device_render_mode(ref.dev)
        env.wet_stride = nw
        env.dry = MixMap(nd, dscale.ctypes.data, dindex.ctypes.data)
        env.wet[0] = MixMap(nw, wscale.ctypes.data, windex.ctypes.data)
        snaps = [ref.snapshot(wet_channels=wet[0], channel=c) for c in range(4)]
        ents, _ = ref.voice_filters(len(combos))
        filt = {(v, p): (a, lp, hp) for v, p, a, lp, hp in ents}
        for k, (is2d, layout, scaling) in enumerate(combos):
            sp = SourceProps()
            brate = C.c_uint32(0)
            assert hz.refh_source_props(ref.ctx, k, C.byref(sp), C.byref(brate)) == 0
            setup = BFormatSetup(C.sizeof(BFormatSetup), is2d, layout, scaling, order.value)
            step = C.c_uint32(0)
            dg = np.full((4, nd), 9.0, dtype=np.float32)
            sg = np.full((4, ns, nw), 9.0, dtype=np.float32)
            fl = (abi.VoiceFilter * (1 + abi.MAX_SENDS))()
            rc = prod.b200mix_calc_voice_bformat(C.byref(sp), C.byref(lis), C.byref(env), brate.value, C.byref(setup),
                                                 C.byref(step), dg.ctypes.data, sg.ctypes.data, fl)
            assert rc == (3 if is2d else 4), rc
            for c in range(rc):
                n, params, coeffs, dry, send, _ = snaps[c]
                assert step.value == params[k].step
                assert np.array_equal(dg[c].view(np.uint32), dry[k].view(np.uint32)), ((is2d, layout, scaling), c, dg[c], dry[k])
                assert np.array_equal(sg[c][0].view(np.uint32), send[k][0].view(np.uint32)), ((is2d, layout, scaling), c)
            assert np.abs(dg[:rc]).max() > 0
            for path in range(1 + ns):
                act, lp, hp = filt[(k, path)]
                f = fl[path]
                assert bool(f.active) == bool(act)
                assert np.array_equal(np.array(list(f.lowpass), dtype=np.float32).view(np.uint32), lp.view(np.uint32))
        setup = BFormatSetup(C.sizeof(BFormatSetup), 0, 1, 2, 3)
        assert prod.b200mix_calc_voice_bformat(C.byref(sp), C.byref(lis), C.byref(env), brate.value, C.byref(setup),
                                               C.byref(step), dg.ctypes.data, sg.ctypes.data, fl) < 0
    finally:
        ref.close()


@pytest.mark.parametrize("devname", ["stereo", "ambi2", "ambi3", "ambi4"])
def test_calc_voice_bformat_higher_orders(devname):
    """b200mix_calc_voice_bformat against live B-Format sources of order 2..4 (AL_SOFT_bformat_hoa: 3D
    and 2D buffers, FuMa up to third order / ACN, every normalisation) on a first-order device and on
    ALC_BFORMAT3D_SOFT devices of order 2, 3 and 4 — the bands above first order turned by the
    recursion of AmbiRotator: every channel's dry and send gain rows and the step, bit for bit.
    Beds the reference would up-sample (device above the source's order, 2D beds on a periphonic
    mix from second order on) are refused."""
    prod = mixlib.product().lib
    prod.b200mix_calc_voice_bformat.argtypes = [C.POINTER(SourceProps), C.POINTER(ListenerParams), C.POINTER(VoiceEnv),
                                                C.c_uint32, C.POINTER(BFormatSetup), C.POINTER(C.c_uint32)] + [C.c_void_p] * 3
    _, hz = refal.libs()
    hz.refh_listener_params.argtypes = [C.c_void_p, C.POINTER(ListenerParams)]
    hz.refh_listener_params.restype = None
    hz.refh_source_props.argtypes = [C.c_void_p, C.c_int, C.POINTER(SourceProps), C.POINTER(C.c_uint32)]
    hz.refh_device_render_mode.argtypes = [C.c_void_p]
    hz.refh_device_ambi.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    hz.refh_device_ambi.restype = None
    hz.refh_dry_ambi_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    hz.refh_slot_ambi_map.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(1300)
    dev_order = {"stereo": 1, "ambi2": 2, "ambi3": 3, "ambi4": 4}[devname]
    a2 = {refal.ALC_HRTF_SOFT: 0, refal.ALC_STEREO_SOURCES: 16}
    if dev_order > 1:
        a2.update({refal.ALC_FORMAT_CHANNELS_SOFT: refal.ALC_BFORMAT3D_SOFT, refal.ALC_AMBISONIC_LAYOUT_SOFT: 1,
                   refal.ALC_AMBISONIC_SCALING_SOFT: 2 if dev_order == 2 else 1, refal.ALC_AMBISONIC_ORDER_SOFT: dev_order})
    ref, _ = scenes.make_ref_scene(0, 0, abi.RS_LINEAR, attrs=a2, max_sources=1)
    try:
        al = ref.al
        al.alListenerfv.argtypes = [C.c_int, C.POINTER(C.c_float)]
        al.alSourcefv.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_float)]
        al.alBufferi.argtypes = [C.c_uint, C.c_int, C.c_int]
        at = rng.standard_normal(3)
        up = np.cross(np.cross(at, rng.standard_normal(3)), at)
        al.alListenerfv(0x100F, (C.c_float * 6)(*[float(x) for x in np.concatenate([at, up])]))
        slot = ref.add_reverb_slot()
        # (order, 2D, layout, scaling): FuMa layouts stop at third order; 2D beds only where the mix is first order
        combos = [(o, 0, layout, scaling) for o in range(max(dev_order, 2), 5) for layout in (0, 1) for scaling in (0, 1, 2)
                  if not (layout == 0 and o > 3) and not (scaling == 0 and o > 3)]
        if dev_order == 1:
            combos += [(o, 1, layout, 1 + (o + layout) % 2) for o in (2, 3, 4) for layout in (0, 1) if not (layout == 0 and o > 3)]
        for k, (o, is2d, layout, scaling) in enumerate(combos):
            nch = 2 * o + 1 if is2d else (o + 1) ** 2
            pcm = np.ascontiguousarray((rng.standard_normal((700, nch)) * 3000).astype(np.int16))
            b = C.c_uint(0); s = C.c_uint(0)
            al.alGenBuffers(1, C.byref(b))
            al.alBufferi(b, 0x199D, o)                # AL_UNPACK_AMBISONIC_ORDER_SOFT
            al.alBufferi(b, 0x1997, layout)           # AL_AMBISONIC_LAYOUT_SOFT
            al.alBufferi(b, 0x1998, scaling)          # AL_AMBISONIC_SCALING_SOFT
            al.alBufferData(b, 0x20022 if is2d else 0x20032, pcm.ctypes.data, pcm.nbytes, 32000)
            assert al.alGetError() == 0, (o, is2d, layout, scaling)
            al.alGenSources(1, C.byref(s))
            al.alSourcei(s, refal.AL_BUFFER, b.value)
            al.alSourcei(s, refal.AL_LOOPING, 1)
            al.alSourcef(s, refal.AL_GAIN, float(rng.uniform(0.2, 1.2)))
            al.alSourcef(s, refal.AL_PITCH, float(rng.uniform(0.5, 1.5)))
            sat = rng.standard_normal(3)
            sup = np.cross(np.cross(sat, rng.standard_normal(3)), sat)
            al.alSourcefv(s, 0x100F, (C.c_float * 6)(*[float(x) for x in np.concatenate([sat, sup])]))
            al.alSourcei(s, 0x202, k % 2)             # AL_SOURCE_RELATIVE
            ref.buffers.append(b.value); ref.sources.append(s.value); ref._keep.append(pcm)
            ref.connect_send(s.value, slot)
        assert al.alGetError() == 0
        ref.play_all()
        ref.render(64)
        nslots, wet = ref.slot_info()
        lis = ListenerParams()
        hz.refh_listener_params(ref.ctx, C.byref(lis))
        order, is2d_dev, xover = C.c_uint32(0), C.c_uint32(0), C.c_float(0.0)
        hz.refh_device_ambi(ref.dev, C.byref(order), C.byref(is2d_dev), C.byref(xover))
        assert order.value == dev_order
        dscale = np.zeros(32, dtype=np.float32); dindex = np.zeros(32, dtype=np.uint32)
        nd = hz.refh_dry_ambi_map(ref.dev, dscale.ctypes.data, dindex.ctypes.data)
        wscale = np.zeros(32, dtype=np.float32); windex = np.zeros(32, dtype=np.uint32)
        nw = hz.refh_slot_ambi_map(ref.ctx, 0, wscale.ctypes.data, windex.ctypes.data)
        ns = ref.desc.num_sends
        env = VoiceEnv()
        env.struct_size = C.sizeof(env)
        env.device_rate, env.num_sends = ref.desc.sample_rate, ns
        env.render_mode = hz.refh_device_render_mode(ref.dev)
        env.wet_stride = nw
        env.dry = MixMap(nd, dscale.ctypes.data, dindex.ctypes.data)
        env.wet[0] = MixMap(nw, wscale.ctypes.data, windex.ctypes.data)
        snaps = [ref.snapshot(wet_channels=wet[0], channel=c) for c in range(25)]
        higher = 0.0
        for k, (o, is2d, layout, scaling) in enumerate(combos):
            sp = SourceProps()
            brate = C.c_uint32(0)
            assert hz.refh_source_props(ref.ctx, k, C.byref(sp), C.byref(brate)) == 0
            setup = BFormatSetup(C.sizeof(BFormatSetup), is2d, layout, scaling, order.value, o, is2d_dev.value)
            step = C.c_uint32(0)
            dg = np.full((25, nd), 9.0, dtype=np.float32)
            sg = np.full((25, ns, nw), 9.0, dtype=np.float32)
            fl = (abi.VoiceFilter * (1 + abi.MAX_SENDS))()
            rc = prod.b200mix_calc_voice_bformat(C.byref(sp), C.byref(lis), C.byref(env), brate.value, C.byref(setup),
                                                 C.byref(step), dg.ctypes.data, sg.ctypes.data, fl)
            mo = min(o, dev_order)        # Voice::prepare mixes the leading channels up to the device's order
            assert rc == (2 * mo + 1 if is2d else (mo + 1) ** 2), (rc, o, is2d, layout, scaling)
            for c in range(rc):
                n, params, coeffs, dry, send, _ = snaps[c]
                assert step.value == params[k].step
                assert np.array_equal(dg[c].view(np.uint32), dry[k].view(np.uint32)), ((o, is2d, layout, scaling), c, dg[c], dry[k])
                assert np.array_equal(sg[c][0].view(np.uint32), send[k][0].view(np.uint32)), ((o, is2d, layout, scaling), c)
            assert np.abs(dg[:rc]).max() > 0
            if dev_order > 1 and not is2d:
                higher = max(higher, float(np.abs(dg[4:(dev_order + 1) ** 2]).max()))
        if dev_order > 1:
            assert higher > 0.05          # the rotated higher bands reach the mix
            # a first-order bed on this device is one the reference up-samples
            setup = BFormatSetup(C.sizeof(BFormatSetup), 0, 1, 2, order.value, 1, is2d_dev.value)
            assert prod.b200mix_calc_voice_bformat(C.byref(sp), C.byref(lis), C.byref(env), brate.value, C.byref(setup),
                                                   C.byref(step), dg.ctypes.data, sg.ctypes.data, fl) == -4
            setup = BFormatSetup(C.sizeof(BFormatSetup), 1, 1, 2, order.value, dev_order, 0)
            assert prod.b200mix_calc_voice_bformat(C.byref(sp), C.byref(lis), C.byref(env), brate.value, C.byref(setup),
                                                   C.byref(step), dg.ctypes.data, sg.ctypes.data, fl) == -4
        # fourth-order FuMa does not exist
        setup = BFormatSetup(C.sizeof(BFormatSetup), 0, 0, 0, order.value, 4, 0)
        assert prod.b200mix_calc_voice_bformat(C.byref(sp), C.byref(lis), C.byref(env), brate.value, C.byref(setup),
                                               C.byref(step), dg.ctypes.data, sg.ctypes.data, fl) == -1
    finally:
        ref.close()


def test_calc_voices_batch_equals_the_single_calls():
    """b200mix_calc_voices (threaded batch) returns exactly what b200mix_calc_voice returns per source."""
    prod = mixlib.product().lib
    prod.b200mix_calc_voice.argtypes = [C.POINTER(SourceProps), C.POINTER(ListenerParams), C.POINTER(VoiceEnv),
                                        C.c_uint32, C.POINTER(abi.VoiceParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
    prod.b200mix_calc_voices.argtypes = [C.c_uint32, C.c_void_p, C.POINTER(ListenerParams), C.POINTER(VoiceEnv),
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_uint32]
    rng = np.random.default_rng(77)
    n, ns, nd, nw = 1000, 2, 4, 4
    lis = ListenerParams()
    lis.struct_size = C.sizeof(lis)
    for i in range(4):
        lis.matrix[i * 4 + i] = 1.0
    lis.gain, lis.meters_per_unit, lis.air_absorption_gain_hf = 1.0, 1.0, 0.99426
    lis.doppler_factor, lis.speed_of_sound, lis.distance_model = 1.0, 343.3, 2
    scale = np.ones(4, dtype=np.float32); index = np.arange(4, dtype=np.uint32)
    props = (SourceProps * n)()
    rates = np.full(n, 44100, dtype=np.uint32)
    for k in range(n):
        p = props[k]
        p.struct_size = C.sizeof(SourceProps)
        p.pitch, p.gain, p.max_gain = float(rng.uniform(0.5, 2)), float(rng.uniform(0.1, 1)), 1.0
        p.inner_angle = p.outer_angle = 360.0
        p.ref_distance, p.max_distance, p.rolloff_factor = 1.0, 1e6, float(rng.uniform(0, 2))
        for i in range(3):
            p.position[i] = float(rng.standard_normal() * 10)
            p.velocity[i] = float(rng.standard_normal() * 5)
        p.distance_model, p.doppler_factor = 2, 1.0
        p.direct.gain, p.direct.gain_hf, p.direct.gain_lf = 1.0, float(rng.uniform(0.2, 1)), 1.0
        p.direct.hf_reference, p.direct.lf_reference = 5000.0, 250.0
        for s in range(ns):
            p.sends[s].gain, p.sends[s].gain_hf, p.sends[s].gain_lf = 1.0, 1.0, 1.0
            p.sends[s].hf_reference, p.sends[s].lf_reference, p.sends[s].active = 5000.0, 250.0, 1
    for mode in (2, 0):
        env = VoiceEnv()
        env.struct_size = C.sizeof(env)
        env.device_rate, env.num_sends, env.render_mode, env.wet_stride = 48000, ns, mode, nw
        env.dry = MixMap(nd, scale.ctypes.data, index.ctypes.data)
        for s in range(ns):
            env.wet[s] = MixMap(nw, scale.ctypes.data, index.ctypes.data)
        want_v = (abi.VoiceParams * n)(); want_d = np.zeros((n, 4), np.float32)
        want_g = np.zeros((n, nd), np.float32); want_s = np.zeros((n, ns, nw), np.float32)
        want_f = (abi.VoiceFilter * (n * (1 + ns)))()
        for k in range(n):
            fl = (abi.VoiceFilter * (1 + abi.MAX_SENDS))()
            want_v[k].voice = k
            assert prod.b200mix_calc_voice(C.byref(props[k]), C.byref(lis), C.byref(env), int(rates[k]),
                                           C.byref(want_v[k]), want_d[k].ctypes.data, want_g[k].ctypes.data,
                                           want_s[k].ctypes.data, fl) == 0
            for j in range(1 + ns):
                want_f[k * (1 + ns) + j] = fl[j]
        for threads in (1, 8):
            got_v = (abi.VoiceParams * n)(); got_d = np.zeros((n, 4), np.float32)
            got_g = np.zeros((n, nd), np.float32); got_s = np.zeros((n, ns, nw), np.float32)
            got_f = (abi.VoiceFilter * (n * (1 + ns)))()
            for k in range(n):
                got_v[k].voice = k
            assert prod.b200mix_calc_voices(n, props, C.byref(lis), C.byref(env), rates.ctypes.data, got_v,
                                            got_d.ctypes.data, got_g.ctypes.data, got_s.ctypes.data, got_f,
                                            threads) == 0
            assert bytes(got_v) == bytes(want_v)
            assert bytes(got_f) == bytes(want_f)
            assert np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32))
            assert np.array_equal(got_g.view(np.uint32), want_g.view(np.uint32))
            assert np.array_equal(got_s.view(np.uint32), want_s.view(np.uint32))
