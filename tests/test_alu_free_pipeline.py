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
