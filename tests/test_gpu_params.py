"""The GPU parameter stage: b200mix_sources_update computes CalcVoiceParams for
point sources on the device from the sources' PROPERTIES.  Checked against the host helper
b200mix_calc_voices (itself pinned bit for bit to live voices of the compiled reference in
tests/test_source_params.py): the voice records must hold the same step / BsincPrepare state /
HRIRs / delays / gains / filter targets — bit-identical except where the host libm's float
function is not correctly rounded (<= 1 ulp, counted) — and the mixed audio must agree."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from helpers import golden, mixlib, synth
from helpers.mixlib import MixDevice
from pyb200mix import abi, scene
from pyb200mix.abi import ListenerParams, ListenerProps, MixMap, SourceProps, SourceVoice, VoiceEnv

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MHR = os.path.join(ROOT, "openal-soft_b200", "data", "Default HRTF.mhr")


def _lib():
    L = mixlib.product().lib
    L.b200mix_calc_listener_params.argtypes = [C.POINTER(ListenerProps), C.POINTER(ListenerParams)]
    L.b200mix_calc_voices.argtypes = [C.c_uint32, C.POINTER(SourceProps), C.POINTER(ListenerParams),
                                      C.POINTER(VoiceEnv), C.c_void_p, C.POINTER(abi.VoiceParams), C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(abi.VoiceFilter), C.c_uint32]
    L.b200mix_sources_update.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(SourceVoice), C.POINTER(SourceProps),
                                         C.POINTER(ListenerParams), C.POINTER(VoiceEnv)]
    L.b200mix_get_voice_targets.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 8
    L.b200mix_hrtf_load.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.b200mix_hrtf_attach.argtypes = [C.c_void_p, C.c_void_p]
    L.b200mix_voices_update_dirs.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 4
    L.b200mix_voices_filters.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.b200mix_last_error.restype = C.c_char_p
    L.b200mix_last_error.argtypes = [C.c_void_p]
    return L


def _listener(L, rng):
    lp = ListenerProps()
    lp.struct_size = C.sizeof(lp)
    at = rng.standard_normal(3)
    up = np.cross(np.cross(at, rng.standard_normal(3)), at)
    for i in range(3):
        lp.position[i] = float(rng.uniform(-2, 2))
        lp.velocity[i] = float(rng.uniform(-5, 5))
        lp.orient_at[i], lp.orient_up[i] = float(at[i]), float(up[i])
    lp.gain, lp.gain_boost = 0.8, 1.0
    lp.meters_per_unit, lp.air_absorption_gain_hf = 1.0, 0.99426
    lp.doppler_factor, lp.doppler_velocity, lp.speed_of_sound = 1.0, 1.0, 343.3
    lp.source_distance_model, lp.distance_model = 1, 2
    out = ListenerParams()
    assert L.b200mix_calc_listener_params(C.byref(lp), C.byref(out)) == 0
    return out


def _props(rng, n, num_sends, plain):
    """Random point sources: every distance model, cones, doppler, radius; unless `plain`,
    direct/send filters, air absorption and a decaying send."""
    arr = (SourceProps * n)()
    for i in range(n):
        P = arr[i]
        P.struct_size = C.sizeof(SourceProps)
        P.pitch = float(rng.uniform(0.5, 2.0)) if i % 16 != 15 else 1.0
        P.gain = float(rng.uniform(0.2, 1.0)) / math.sqrt(n) * 4.0
        P.outer_gain = float(rng.uniform(0.0, 1.0))
        P.min_gain, P.max_gain = 0.0, 1.0
        P.inner_angle = float(rng.uniform(30, 360)) if i % 3 == 0 else 360.0
        P.outer_angle = min(360.0, P.inner_angle + float(rng.uniform(0, 120)))
        P.ref_distance = float(rng.uniform(0.5, 2.0))
        P.max_distance = float(rng.uniform(5.0, 50.0))
        P.rolloff_factor = float(rng.uniform(0.2, 2.0))
        d = rng.standard_normal(3)
        d = d / np.linalg.norm(d) * rng.uniform(0.3, 12.0)
        if i == 5:
            d = np.zeros(3)                       # a source on the listener (head-relative)
        for k in range(3):
            P.position[k] = float(d[k])
            P.velocity[k] = float(rng.uniform(-20, 20))
            P.direction[k] = float(rng.standard_normal()) if i % 3 == 0 else 0.0
        P.head_relative = 1 if i == 5 else int(i % 7 == 0)
        P.distance_model = int(i % 7)
        P.dry_gain_hf_auto, P.wet_gain_auto, P.wet_gain_hf_auto = 1, 1, 1
        P.outer_gain_hf = 1.0 if plain else float(rng.uniform(0.2, 1.0))
        P.air_absorption_factor = 0.0 if plain else float(rng.uniform(0.0, 4.0)) * (i % 2)
        P.room_rolloff_factor = float(rng.uniform(0.0, 1.0))
        P.doppler_factor = float(rng.uniform(0.0, 1.5))
        P.radius = float(rng.uniform(0.0, 3.0)) if i % 4 == 0 else 0.0
        P.direct.gain = 1.0
        P.direct.gain_hf = 1.0 if plain else float(rng.uniform(0.1, 1.0)) if i % 2 else 1.0
        P.direct.gain_lf = 1.0 if plain else float(rng.uniform(0.3, 1.0)) if i % 5 == 0 else 1.0
        P.direct.hf_reference, P.direct.lf_reference = 5000.0, 250.0
        for s in range(abi.MAX_SENDS):
            S = P.sends[s]
            S.gain, S.gain_hf, S.gain_lf = 1.0, 1.0, 1.0
            S.hf_reference, S.lf_reference = 5000.0, 250.0
            S.active = 1 if s < num_sends else 0
            S.slot_room_rolloff, S.slot_decay_time, S.slot_air_absorption_gain_hf = 0.0, 1.49, 0.994
            if not plain and s < num_sends and i % 3 == 1:
                S.gain_hf = float(rng.uniform(0.2, 1.0))
        P.orient_at[2], P.orient_up[1] = -1.0, 1.0
    return arr


def _ulp_diff(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7fffffff), a)
    b = np.where(b < 0, -(b & 0x7fffffff), b)
    return np.abs(a - b)


@pytest.mark.parametrize("mode,plain", [(2, True), (2, False), (1, False), (0, True)])
def test_sources_update_equals_host_calc_voices(mode, plain):
    if mode == 2 and not os.path.exists(MHR):
        pytest.skip("HRTF data set not staged (run build())")
    L = _lib()
    rng = np.random.default_rng(900 + mode * 2 + int(plain))
    nv, num_sends, cw = 96, 1, 4
    if mode == 2:
        desc = synth.hrtf_desc(nv, 64)
    else:
        desc = synth.stereo_desc(nv, dry_channels=3)
    desc.num_sends, desc.wet_channels, desc.max_slots = num_sends, cw, 1
    cd = desc.dry_channels
    fx = golden.load("hrtf_bsinc24_reverb_v6")
    dscale = np.array([1.0, 0.8, 1.1, 0.9][:cd], dtype=np.float32)
    dindex = np.array([0, 1, 3, 2][:cd], dtype=np.uint32)
    wscale = np.ones(cw, dtype=np.float32)
    windex = np.arange(cw, dtype=np.uint32)
    env = VoiceEnv()
    env.struct_size = C.sizeof(env)
    env.device_rate, env.num_sends, env.render_mode, env.wet_stride = 48000, num_sends, mode, cw
    env.dry = MixMap(cd, dscale.ctypes.data, dindex.ctypes.data)
    env.wet[0] = MixMap(cw, wscale.ctypes.data, windex.ctypes.data)
    hrtf = C.c_void_p()
    if mode == 2:
        data = open(MHR, "rb").read()
        assert L.b200mix_hrtf_load(data, len(data), C.byref(hrtf)) == 0
    resamplers = [abi.RS_BSINC24, abi.RS_FAST_BSINC12, abi.RS_SPLINE, abi.RS_BSINC48, abi.RS_LINEAR]

    devs = []
    for _ in range(2):
        dev = MixDevice(mixlib.product(), desc)
        if mode == 2:
            dev.set_hrtf_decoder(*synth.decoder(np.random.default_rng(7)))
            assert L.b200mix_hrtf_attach(dev.h, hrtf) == 0
        else:
            dev.set_ambi_decoder((np.random.default_rng(8).standard_normal((cd, 2)) * 0.5).astype(np.float32), None, 0.0)
        for i in range(nv):
            dev.buffer_data(i, abi.FMT_I16, scene.voice_buffer_fast(i))
        dev.slot_reverb(0, abi.reverb_params_from(fx["reverb_params"].tobytes()), fx["reverb_gains"])
        devs.append(dev)
    host_dev, gpu_dev = devs

    outs = [[], []]
    worst = {}
    for upd in range(5):
        lis = _listener(L, np.random.default_rng(50 + upd // 2))
        props = _props(np.random.default_rng(1000 * upd + mode), nv, num_sends, plain)
        reset = abi.VF_RESET if upd == 0 else 0
        # ---- host path: b200mix_calc_voices -> voices_update(_dirs) + voices_filters
        vp = (abi.VoiceParams * nv)()
        sv = (SourceVoice * nv)()
        rates = np.full(nv, 48000, dtype=np.uint32)
        for i in range(nv):
            for rec in (vp[i], sv[i]):
                rec.voice, rec.buffer = i, i
                rec.flags = abi.VF_PLAYING | abi.VF_STATIC | abi.VF_LOOPING | reset
                rec.resampler = resamplers[i % len(resamplers)]
                rec.position, rec.position_frac = 100 * i, 0
                rec.loop_start, rec.loop_end = 0, scene.BUFFER_FRAMES
                for s in range(abi.MAX_SENDS):
                    rec.send_slot[s] = 0 if s < num_sends else abi.NO_SLOT
            sv[i].buffer_rate = 48000
        dirs = np.zeros((nv, 4), dtype=np.float32)
        dry = np.zeros((nv, cd), dtype=np.float32)
        send = np.zeros((nv, num_sends, cw), dtype=np.float32)
        filt = (abi.VoiceFilter * (nv * (1 + num_sends)))()
        assert L.b200mix_calc_voices(nv, props, C.byref(lis), C.byref(env), rates.ctypes.data, vp, dirs.ctypes.data,
                                     dry.ctypes.data, send.ctypes.data, filt, 1) == 0
        if mode == 2:
            assert L.b200mix_voices_update_dirs(host_dev.h, nv, vp, dirs.ctypes.data, None, send.ctypes.data) == 0
        else:
            host_dev.voices_update(vp, None, dry, send)
        if not plain:
            assert L.b200mix_voices_filters(host_dev.h, nv * (1 + num_sends), filt) == 0
        # ---- GPU path
        rc = L.b200mix_sources_update(gpu_dev.h, nv, sv, props, C.byref(lis), C.byref(env))
        assert rc == 0, L.b200mix_last_error(gpu_dev.h)

        # ---- the voice records
        for v in range(nv):
            got = []
            for dev in devs:
                step = C.c_uint32()
                bs = np.zeros(4, dtype=np.float32)
                hg = C.c_float()
                hd = (C.c_uint32 * 2)()
                hc = np.zeros((64, 2), dtype=np.float32)
                dg = np.zeros(cd, dtype=np.float32)
                sg = np.zeros((num_sends, cw), dtype=np.float32)
                fl = np.zeros((1 + num_sends, 11), dtype=np.float32)
                assert L.b200mix_get_voice_targets(dev.h, v, C.byref(step), bs.ctypes.data, C.byref(hg), hd,
                                                   hc.ctypes.data if mode == 2 else None, dg.ctypes.data,
                                                   sg.ctypes.data, fl.ctypes.data) == 0
                got.append((step.value, bs.copy(), np.float32(hg.value), (hd[0], hd[1]), hc, dg, sg, fl))
            a, b = got
            assert a[0] == b[0], (v, a[0], b[0])          # the step involves no libm call: exact
            assert a[1][1:].tobytes() == b[1][1:].tobytes(), (v, "bsinc m/l/offset")
            for name, x, y in (("bsinc_sf", a[1][:1], b[1][:1]), ("hrtf_gain", [a[2]], [b[2]]), ("hrir", a[4], b[4]),
                               ("dry", a[5], b[5]), ("send", a[6], b[6]), ("filters", a[7], b[7])):
                if mode != 2 and name in ("hrtf_gain", "hrir"):
                    continue
                if plain and name == "filters":
                    continue
                d = _ulp_diff(x, y)
                worst[name] = max(worst.get(name, 0), int(d.max()))
                # a direction that differs by an ulp moves the 4-HRIR blend weights by ~1e-7
                # a direction that differs by an ulp moves the 4-HRIR blend weights by ~1e-7: the
                # blended taps (|c| <~ 2) then differ by an ulp of 1.0, whatever their own size
                # gains are products of a few libm results (pow, cos, sqrt of the spread): a handful of ulps
                lim = 16
                absd = float(np.abs(np.asarray(x, dtype=np.float64) - y).max())
                # (the azimuth index v = (az/2pi + 1)*180 has an ulp of 3e-5: an ulp of azimuth can move
                # the blend factor by that much, times the difference of neighbouring HRIRs)
                # filter coefficients: a 1-ulp HF gain (pow of the air absorption) moves the shelf design's
                # b/a terms (sums of O(1) numbers that nearly cancel) by a few ulps of 1.0
                tol = {"hrir": 5e-6, "filters": 2e-6}.get(name, 1e-7)
                assert d.max() <= lim or absd < tol, (v, name, int(d.max()), absd)
            assert max(abs(int(a[3][0]) - int(b[3][0])), abs(int(a[3][1]) - int(b[3][1]))) <= 1, (v, "HRIR delays")
            worst["delay_mismatches"] = worst.get("delay_mismatches", 0) + int(a[3] != b[3])
        for k, dev in enumerate(devs):
            outs[k].append(dev.render(1024 if upd % 2 == 0 else 333))
    for dev in devs:
        dev.close()
    a, b = np.concatenate(outs[0], axis=1), np.concatenate(outs[1], axis=1)
    assert np.abs(a).max() > 1e-3
    err = a.astype(np.float64) - b
    assert np.sqrt((err ** 2).mean()) <= 1e-6 and np.abs(err).max() <= 1e-5, (np.abs(err).max(), worst)
    print("worst ulp differences host vs GPU parameter stage:", worst)
