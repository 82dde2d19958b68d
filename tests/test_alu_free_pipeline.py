"""The library's own parameter stage driving a mixer end to end, against the reference's audio.

A scene of moving mono sources (direct and send filters, a reverb send) is set up on the compiled
reference through the AL API.  The mixer under test never sees the reference's computed voice
parameters: every update, the sources' *properties* go through b200mix_calc_voice (the product's
host helper) and b200mix_hrtf_get_coeffs, and the resulting b200mix_voices_update /
b200mix_voices_filters / reverb structures (b200mix_reverb_params_from_efx) feed a device of the
C ABI.  The rendered audio must equal the reference's own render.  On this CPU-only box the device
is the oracle (same ABI, tests only); the CUDA device behind the same calls is checked against the
oracle by tests/test_gpu_parity.py."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import mixlib, refal, scenes
from helpers.mixlib import MixDevice
from pyb200mix import abi, scene
from test_source_params import (ListenerParams, SourceProps, VoiceEnv, MixMap, MHR, _f3)

pytestmark = pytest.mark.ref


@pytest.mark.skipif(not os.path.exists(MHR), reason="HRTF data set not staged (run build())")
@pytest.mark.parametrize("hrtf", [1, 0])
def test_calc_voice_drives_a_mixer_to_the_references_audio(hrtf):
    prod = mixlib.product().lib
    prod.b200mix_calc_voice.argtypes = [C.POINTER(SourceProps), C.POINTER(ListenerParams), C.POINTER(VoiceEnv),
                                        C.c_uint32, C.POINTER(abi.VoiceParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
    prod.b200mix_hrtf_load.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    prod.b200mix_hrtf_get_coeffs.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                             C.POINTER(C.c_uint32)]
    prod.b200mix_hrtf_free.argtypes = [C.c_void_p]
    prod.b200mix_reverb_params_from_efx.argtypes = [C.POINTER(abi.EfxReverb), C.POINTER(abi.ReverbTarget),
                                                    C.POINTER(abi.ReverbParams), C.c_void_p]
    _, hz = refal.libs()
    hz.refh_listener_params.argtypes = [C.c_void_p, C.POINTER(ListenerParams)]
    hz.refh_listener_params.restype = None
    hz.refh_source_props.argtypes = [C.c_void_p, C.c_int, C.POINTER(SourceProps), C.POINTER(C.c_uint32)]
    hz.refh_device_render_mode.argtypes = [C.c_void_p]
    hz.refh_dry_ambi_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    hz.refh_slot_ambi_map.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    hz.refh_device_ambi.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    hz.refh_device_ambi.restype = None
    hstore = C.c_void_p()
    data = open(MHR, "rb").read()
    assert prod.b200mix_hrtf_load(data, len(data), C.byref(hstore)) == 0

    V, U = 8, 6
    rng = np.random.default_rng(1500 + hrtf)
    ref, pcms = scenes.make_ref_scene(V, hrtf, abi.RS_BSINC24 if hrtf else abi.RS_SPLINE)
    dev = None
    try:
        al = ref.al
        al.alListener3f.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
        al.alListenerfv.argtypes = [C.c_int, C.POINTER(C.c_float)]
        al.alListener3f(0x1004, 0.3, -0.2, 0.5)
        at = rng.standard_normal(3)
        up = np.cross(np.cross(at, rng.standard_normal(3)), at)
        al.alListenerfv(0x100F, (C.c_float * 6)(*[float(x) for x in np.concatenate([at, up])]))
        slot_gain = 0.7
        slot = ref.add_reverb_slot(slot_gain=slot_gain)
        for k, src in enumerate(ref.sources):
            al.alSource3f(src, 0x1006, *_f3(rng, 10.0))
            al.alSourcef(src, 0x1021, float(rng.uniform(0.2, 2.0)))
            if k % 2 == 0:
                ref.set_direct_filter(src, ref.make_filter(0.9, float(rng.uniform(0.2, 1.0))))
            ref.connect_send(src, slot, 0, refal.AL_FILTER_NULL if k % 3 else ref.make_filter(0.8, 0.5))
        assert al.alGetError() == 0
        ref.play_all()

        lis = ListenerParams()
        dscale = np.zeros(32, dtype=np.float32); dindex = np.zeros(32, dtype=np.uint32)
        wscale = np.zeros(32, dtype=np.float32); windex = np.zeros(32, dtype=np.uint32)
        worst = 0.0
        for u in range(U):
            if u:
                # move half of the sources (positions only: the reference recomputes just those voices)
                for k in range(u % 2, V, 2):
                    al.alSource3f(ref.sources[k], 0x1004, *_f3(rng, 4.0))
            out_ref = ref.render(1024)
            if dev is None:
                dev = scenes.mirror_device(mixlib.oracle(), ref, V, pcms)
                if hrtf:
                    # the HRTF decoder too comes from the library's own builder, not from the reference
                    prod.b200mix_hrtf_build_decoder.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32,
                                                                C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p,
                                                                C.POINTER(C.c_float)]
                    irs, sc = C.c_uint32(0), C.c_float(0.0)
                    dco = np.zeros(4 * 128 * 2, dtype=np.float32)
                    dhf = np.zeros(4, dtype=np.float32)
                    assert prod.b200mix_hrtf_build_decoder(hstore, 1, ref.desc.ir_size, C.byref(irs), dco.ctypes.data,
                                                           dhf.ctypes.data, C.byref(sc)) == 4
                    dev.set_hrtf_decoder(dco[:4 * irs.value * 2].reshape(4, irs.value, 2), dhf,
                                         np.full(4, sc.value, dtype=np.float32))
                nd = hz.refh_dry_ambi_map(ref.dev, dscale.ctypes.data, dindex.ctypes.data)
                nw = hz.refh_slot_ambi_map(ref.ctx, 0, wscale.ctypes.data, windex.ctypes.data)
                ns = ref.desc.num_sends
                env = VoiceEnv()
                env.struct_size = C.sizeof(env)
                env.device_rate, env.num_sends = ref.desc.sample_rate, ns
                env.render_mode = hz.refh_device_render_mode(ref.dev)
                env.wet_stride = nw
                env.dry = MixMap(nd, dscale.ctypes.data, dindex.ctypes.data)
                env.wet[0] = MixMap(nw, wscale.ctypes.data, windex.ctypes.data)
                # the reverb from its EFX defaults, through the library's own parameter stage
                props = abi.EfxReverb(C.sizeof(abi.EfxReverb), 1.0, 1.0, 0.32, 0.89, 1.0, 1.49, 0.83, 1.0, 0.05, 0.007,
                                      (C.c_float * 3)(0, 0, 0), 1.26, 0.011, (C.c_float * 3)(0, 0, 0), 0.25, 0.0, 0.25,
                                      0.0, 0.994, 5000.0, 250.0, 0.0, 1)
                order, is2d, xover = C.c_uint32(0), C.c_uint32(0), C.c_float(0.0)
                hz.refh_device_ambi(ref.dev, C.byref(order), C.byref(is2d), C.byref(xover))
                tgt = abi.ReverbTarget(C.sizeof(abi.ReverbTarget), ref.desc.sample_rate, order.value, is2d.value,
                                       xover.value, slot_gain, 1.0, nd, dscale.ctypes.data, dindex.ctypes.data)
                rp = abi.ReverbParams()
                rg = np.zeros((8, nd), dtype=np.float32)
                assert prod.b200mix_reverb_params_from_efx(C.byref(props), C.byref(tgt), C.byref(rp), rg.ctypes.data) == 0
                dev.slot_reverb(0, rp, rg)
            hz.refh_listener_params(ref.ctx, C.byref(lis))
            plist, coeffs, drys, sends, fents = [], [], [], [], []
            for k in range(V):
                sp = SourceProps()
                brate = C.c_uint32(0)
                assert hz.refh_source_props(ref.ctx, k, C.byref(sp), C.byref(brate)) == 0
                vp = abi.VoiceParams()
                vp.voice, vp.buffer, vp.resampler = k, k, (abi.RS_BSINC24 if hrtf else abi.RS_SPLINE)
                vp.flags = abi.VF_PLAYING | abi.VF_STATIC | abi.VF_LOOPING | (abi.VF_RESET if u == 0 else 0)
                vp.loop_start, vp.loop_end = 0, len(pcms[k])
                for s in range(abi.MAX_SENDS):
                    vp.send_slot[s] = 0 if (s == 0 and sp.sends[0].active) else abi.NO_SLOT
                d4 = np.zeros(4, dtype=np.float32)
                dg = np.zeros(nd, dtype=np.float32)
                sg = np.zeros((ns, nw), dtype=np.float32)
                fl = (abi.VoiceFilter * (1 + abi.MAX_SENDS))()
                assert prod.b200mix_calc_voice(C.byref(sp), C.byref(lis), C.byref(env), brate.value, C.byref(vp),
                                               d4.ctypes.data, dg.ctypes.data, sg.ctypes.data, fl) == 0
                co = np.zeros((max(ref.desc.ir_size, 1), 2), dtype=np.float32)
                if hrtf:
                    dl = (C.c_uint32 * 2)()
                    assert prod.b200mix_hrtf_get_coeffs(hstore, d4[0], d4[1], d4[2], d4[3], co.ctypes.data, dl) == 0
                    vp.hrtf_delay[0], vp.hrtf_delay[1] = dl[0], dl[1]
                plist.append(vp); coeffs.append(co); drys.append(dg); sends.append(sg)
                for path in range(1 + ns):
                    f = fl[path]
                    fents.append((k, path, f.active, list(f.lowpass), list(f.highpass)))
            dev.voices_update(plist, np.stack(coeffs) if hrtf else None, np.stack(drys), np.stack(sends))
            dev.voices_filters(fents)
            out = dev.render(1024)
            err = float(np.abs(out.astype(np.float64) - out_ref).max())
            worst = max(worst, err)
            assert err <= 2e-6, (hrtf, u, err)
            assert np.abs(out_ref).max() > 1e-3
    finally:
        if dev is not None:
            dev.close()
        ref.close()
        prod.b200mix_hrtf_free(hstore)


@pytest.mark.skipif(not os.path.exists(MHR), reason="HRTF data set not staged (run build())")
@pytest.mark.parametrize("hrtf", [1, 0])
def test_multichannel_sources_as_one_voice_per_channel(hrtf):
    """Stereo music, a 5.1 source and first-order B-Format beds, driven through
    b200mix_calc_voice_channels / b200mix_calc_voice_bformat as ONE ABI VOICE PER BUFFER CHANNEL
    (B200MIX_VF_CHANNEL), must give the reference's audio: the mapping INTEGRATION.md describes for
    Voice::mChans."""
    from pyb200mix.abi import ChannelSetup, BFormatSetup
    prod = mixlib.product().lib
    prod.b200mix_calc_voice_channels.argtypes = [C.POINTER(SourceProps), C.POINTER(ListenerParams), C.POINTER(VoiceEnv),
                                                 C.c_uint32, C.POINTER(ChannelSetup), C.POINTER(C.c_uint32)] + [C.c_void_p] * 5
    prod.b200mix_calc_voice_bformat.argtypes = [C.POINTER(SourceProps), C.POINTER(ListenerParams), C.POINTER(VoiceEnv),
                                                C.c_uint32, C.POINTER(BFormatSetup), C.POINTER(C.c_uint32)] + [C.c_void_p] * 3
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
    hstore = C.c_void_p()
    data = open(MHR, "rb").read()
    assert prod.b200mix_hrtf_load(data, len(data), C.byref(hstore)) == 0

    rng = np.random.default_rng(1700 + hrtf)
    # (AL format, channels, kind, setup values)
    specs = [(0x1103, 2, "chan", 2), (0x120B, 6, "chan", 5), (0x20032, 4, "bf", (0, 1, 2)), (0x20022, 3, "bf", (1, 0, 0))]
    a2 = {refal.ALC_HRTF_SOFT: hrtf, refal.ALC_STEREO_SOURCES: 8}
    ref, _ = scenes.make_ref_scene(0, hrtf, abi.RS_SPLINE, attrs=a2, max_sources=1)
    dev = None
    try:
        al = ref.al
        al.alBufferi.argtypes = [C.c_uint, C.c_int, C.c_int]
        al.alSourcefv.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_float)]
        pcms = []
        for k, (fmt, nch, kind, val) in enumerate(specs):
            pcm = np.ascontiguousarray((rng.standard_normal((9000, nch)) * 2500).astype(np.int16))
            b = C.c_uint(0); s = C.c_uint(0)
            al.alGenBuffers(1, C.byref(b))
            if kind == "bf":
                al.alBufferi(b, 0x1997, val[1]); al.alBufferi(b, 0x1998, val[2])
            al.alBufferData(b, fmt, pcm.ctypes.data, pcm.nbytes, 44100)
            al.alGenSources(1, C.byref(s))
            al.alSourcei(s, refal.AL_BUFFER, b.value)
            al.alSourcei(s, refal.AL_LOOPING, 1)
            al.alSourcei(s, refal.AL_SOURCE_RESAMPLER_SOFT, abi.RS_SPLINE)
            al.alSourcef(s, refal.AL_GAIN, float(rng.uniform(0.2, 0.6)))
            al.alSourcef(s, refal.AL_PITCH, float(rng.uniform(0.7, 1.3)))
            if kind == "bf":
                sat = rng.standard_normal(3)
                sup = np.cross(np.cross(sat, rng.standard_normal(3)), sat)
                al.alSourcefv(s, 0x100F, (C.c_float * 6)(*[float(x) for x in np.concatenate([sat, sup])]))
            ref.buffers.append(b.value); ref.sources.append(s.value); ref._keep.append(pcm)
            pcms.append(pcm)
        assert al.alGetError() == 0
        ref.play_all()
        dscale = np.zeros(32, dtype=np.float32); dindex = np.zeros(32, dtype=np.uint32)
        lis = ListenerParams()
        # nothing moves: the reference computes its voice parameters once, on its first update
        refs = [ref.render(1024) for _ in range(3)]
        hz.refh_listener_params(ref.ctx, C.byref(lis))
        desc = abi.DeviceDesc()
        C.memmove(C.byref(desc), C.byref(ref.desc), C.sizeof(desc))
        desc.max_voices = sum(n for _, n, _, _ in specs)
        desc.max_buffers = len(specs)
        desc.max_slots = 0
        desc.wet_channels = 0
        dev = MixDevice(mixlib.oracle(), desc)
        if desc.post_process == abi.POST_HRTF:
            dev.set_hrtf_decoder(*ref.hrtf_decoder())
        else:
            dev.set_ambi_decoder(*ref.ambi_decoder())
        for k, pcm in enumerate(pcms):
            dev.buffer_data(k, abi.FMT_I16, pcm, channels=pcm.shape[1])
        nd = hz.refh_dry_ambi_map(ref.dev, dscale.ctypes.data, dindex.ctypes.data)
        env = VoiceEnv()
        env.struct_size = C.sizeof(env)
        env.device_rate, env.num_sends = ref.desc.sample_rate, 0
        env.render_mode = hz.refh_device_render_mode(ref.dev)
        env.dry = MixMap(nd, dscale.ctypes.data, dindex.ctypes.data)
        plist, coeffs, drys = [], [], []
        vidx = 0
        for k, (fmt, nch, kind, val) in enumerate(specs):
            sp = SourceProps()
            brate = C.c_uint32(0)
            assert hz.refh_source_props(ref.ctx, k, C.byref(sp), C.byref(brate)) == 0
            step = C.c_uint32(0)
            hg = np.zeros(8, dtype=np.float32); dirs = np.zeros((8, 4), dtype=np.float32)
            dg = np.zeros((8, nd), dtype=np.float32)
            fl = (abi.VoiceFilter * (1 + abi.MAX_SENDS))()
            if kind == "chan":
                setup = ChannelSetup(C.sizeof(ChannelSetup), val, (C.c_float * 2)(float(np.float32(np.pi / 6)),
                                     float(np.float32(-np.pi / 6))), 0.0, abi.NO_SLOT, 0)
                rc = prod.b200mix_calc_voice_channels(C.byref(sp), C.byref(lis), C.byref(env), brate.value,
                                                      C.byref(setup), C.byref(step), hg.ctypes.data, dirs.ctypes.data,
                                                      dg.ctypes.data, None, fl)
                per_channel_hrtf = bool(hrtf)
            else:
                setup = BFormatSetup(C.sizeof(BFormatSetup), val[0], val[1], val[2], 1)
                rc = prod.b200mix_calc_voice_bformat(C.byref(sp), C.byref(lis), C.byref(env), brate.value,
                                                     C.byref(setup), C.byref(step), dg.ctypes.data, None, fl)
                per_channel_hrtf = False      # ambisonic sources mix into the Dry bus on every device
            assert rc == nch, (kind, rc)
            for c in range(nch):
                vp = abi.VoiceParams()
                vp.voice, vp.buffer, vp.resampler, vp.step = vidx, k, abi.RS_SPLINE, step.value
                vp.flags = (abi.VF_PLAYING | abi.VF_STATIC | abi.VF_LOOPING | abi.VF_RESET | abi.vf_channel(c)
                            | (abi.VF_HRTF if per_channel_hrtf else 0))
                vp.loop_start, vp.loop_end = 0, pcms[k].shape[0]
                for s in range(abi.MAX_SENDS):
                    vp.send_slot[s] = abi.NO_SLOT
                co = np.zeros((max(ref.desc.ir_size, 1), 2), dtype=np.float32)
                if per_channel_hrtf and hg[c] != 0.0:
                    dl = (C.c_uint32 * 2)()
                    assert prod.b200mix_hrtf_get_coeffs(hstore, dirs[c][0], dirs[c][1], dirs[c][2], dirs[c][3],
                                                        co.ctypes.data, dl) == 0
                    vp.hrtf_delay[0], vp.hrtf_delay[1] = dl[0], dl[1]
                vp.hrtf_gain = float(hg[c])
                plist.append(vp); coeffs.append(co); drys.append(dg[c].copy())
                vidx += 1
        dev.voices_update(plist, np.stack(coeffs) if hrtf else None, np.stack(drys), None)
        for u in range(3):
            out = dev.render(1024)
            err = float(np.abs(out.astype(np.float64) - refs[u]).max())
            assert err <= 2e-6, (hrtf, u, err)
            assert np.abs(refs[u]).max() > 1e-3
    finally:
        if dev is not None:
            dev.close()
        ref.close()
        prod.b200mix_hrtf_free(hstore)
