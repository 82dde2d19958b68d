"""Uniform ctypes front-end over the two implementations of the include/b200mix.h
surface: the CUDA product (libb200mix.so, prefix b200mix_) and the CPU oracle
(oracle/liboracle.so, prefix oracle_).  Tests drive both with the same calls."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "openal-soft_b200"))
from pyb200mix import abi  # noqa: E402

ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
PRODUCT_SO = os.path.join(ROOT, "openal-soft_b200", "libb200mix.so")


class MixLib:
    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L = self.lib
        f = lambda name: getattr(L, prefix + name)  # noqa: E731
        self.create = f("create")
        self.create.argtypes = [C.POINTER(abi.DeviceDesc), C.POINTER(C.c_void_p)]
        self.destroy = f("destroy")
        self.destroy.argtypes = [C.c_void_p]
        self.destroy.restype = None
        self.set_hrtf_decoder = f("set_hrtf_decoder")
        self.set_hrtf_decoder.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32] + [C.c_void_p] * 3
        self.set_ambi_decoder = f("set_ambi_decoder")
        self.set_ambi_decoder.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float]
        self.buffer_data = f("buffer_data")
        self.buffer_data.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_void_p, C.c_size_t]
        self.buffer_data_adpcm = f("buffer_data_adpcm")
        self.buffer_data_adpcm.argtypes = [C.c_void_p] + [C.c_uint32] * 5 + [C.c_void_p, C.c_size_t]
        self.voices_update = f("voices_update")
        self.voices_update.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        self.render = f("render")
        self.render.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.c_void_p]
        self.slot_convolution = f("slot_convolution")
        self.slot_convolution.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        self.slot_output_gains = f("slot_output_gains")
        self.slot_output_gains.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        self.slot_reverb = f("slot_reverb")
        self.slot_reverb.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(abi.ReverbParams)]
        self.slot_reverb_update = f("slot_reverb_update")
        self.slot_reverb_update.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(abi.ReverbParams), C.c_uint32]
        self.slot_target = f("slot_target")
        self.slot_target.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        self.slot_disable = f("slot_disable")
        self.slot_disable.argtypes = [C.c_void_p, C.c_uint32]
        self.voices_filters = f("voices_filters")
        self.voices_filters.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        self.biquad_coeffs = f("biquad_coeffs")
        self.biquad_coeffs.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_void_p]
        self.voice_queue = f("voice_queue")
        self.voice_queue.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
        self.render_interleaved = f("render_interleaved")
        self.render_interleaved.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                            C.c_float, C.POINTER(C.c_uint32), C.c_void_p]
        self.set_limiter = f("set_limiter")
        self.set_limiter.argtypes = [C.c_void_p, C.POINTER(abi.LimiterDesc), C.POINTER(C.c_uint32)]
        self.set_front_stabilizer = f("set_front_stabilizer")
        self.set_front_stabilizer.argtypes = [C.c_void_p, C.c_uint32, C.c_float]
        self.set_bs2b = f("set_bs2b")
        self.set_bs2b.argtypes = [C.c_void_p, C.c_uint32]
        self.set_uhj_encoder = f("set_uhj_encoder")
        self.set_uhj_encoder.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        self.set_distance_comp = f("set_distance_comp")
        self.set_distance_comp.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        self.render_begin = f("render_begin")
        self.render_begin.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        self.render_end = f("render_end")
        self.render_end.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(C.c_void_p)]
        self.get_dry = f("get_dry")
        self.get_dry.argtypes = [C.c_void_p, C.c_void_p]
        self.slot_efx = f("slot_efx")
        self.slot_efx.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(abi.EfxProps), C.POINTER(abi.EfxTarget)]
        if prefix == "b200mix_":          # sharded device sets exist on the product only
            self.shard_init = f("shard_init")
            self.shard_init.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
            self.shard_connect = f("shard_connect")
            self.shard_connect.argtypes = [C.c_void_p, C.c_void_p]
            self.shard_nccl_id = f("shard_nccl_id")
            self.shard_nccl_id.argtypes = [C.c_void_p]
            self.shard_nccl = f("shard_nccl")
            self.shard_nccl.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
            self.shard_last_us = f("shard_last_us")
            self.shard_last_us.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]


class MixDevice:
    """One device on either implementation."""

    def __init__(self, mixlib: MixLib, desc: abi.DeviceDesc):
        self.m = mixlib
        self.desc = abi.DeviceDesc()
        C.memmove(C.byref(self.desc), C.byref(desc), C.sizeof(desc))
        self.desc.struct_size = C.sizeof(abi.DeviceDesc)
        h = C.c_void_p()
        rc = self.m.create(C.byref(self.desc), C.byref(h))
        assert rc == 0, f"{mixlib.prefix}create -> {rc}"
        self.h = h

    def close(self):
        if self.h:
            self.m.destroy(self.h)
            self.h = None

    def set_hrtf_decoder(self, coeffs, hf, sc):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float32)
        hf = np.ascontiguousarray(hf, dtype=np.float32)
        sc = np.ascontiguousarray(sc, dtype=np.float32)
        rc = self.m.set_hrtf_decoder(self.h, coeffs.shape[0], coeffs.shape[1], coeffs.ctypes.data,
                                     hf.ctypes.data, sc.ctypes.data)
        assert rc == 0, rc

    def set_ambi_decoder(self, hfm, lfm, xover):
        hfm = np.ascontiguousarray(hfm, dtype=np.float32)
        lfp = None
        if lfm is not None:
            lfm = np.ascontiguousarray(lfm, dtype=np.float32)
            lfp = lfm.ctypes.data
        rc = self.m.set_ambi_decoder(self.h, hfm.shape[0], hfm.ctypes.data, lfp, xover)
        assert rc == 0, rc

    def slot_convolution(self, slot, ir, gains):
        """ir: [channels][frames] float32; gains: [channels][dry_channels]."""
        ir = np.ascontiguousarray(np.atleast_2d(ir), dtype=np.float32)
        gains = np.ascontiguousarray(np.atleast_2d(gains), dtype=np.float32)
        rc = self.m.slot_convolution(self.h, slot, ir.shape[0], ir.shape[1], ir.ctypes.data)
        assert rc == 0, rc
        rc = self.m.slot_output_gains(self.h, slot, gains.shape[0], gains.ctypes.data)
        assert rc == 0, rc

    def slot_reverb(self, slot, params, gains):
        """params: abi.ReverbParams; gains: [8][dry_channels] (early 0-3, late 0-3)."""
        params.struct_size = C.sizeof(abi.ReverbParams)
        rc = self.m.slot_reverb(self.h, slot, C.byref(params))
        assert rc == 0, rc
        gains = np.ascontiguousarray(gains, dtype=np.float32)
        rc = self.m.slot_output_gains(self.h, slot, 8, gains.ctypes.data)
        assert rc == 0, rc

    def slot_efx(self, slot, props, slot_gain, out_scale, out_index, wet_index, ambi_order=1,
                 real_center=abi.NO_SLOT, real_lfe=abi.NO_SLOT, expect=0):
        """b200mix_slot_efx: props = abi.EfxProps; the maps are the target mix's AmbiMap and the slot's
        Wet.AmbiMap indices."""
        out_scale = np.ascontiguousarray(out_scale, dtype=np.float32)
        out_index = np.ascontiguousarray(out_index, dtype=np.uint32)
        wet_index = np.ascontiguousarray(wet_index, dtype=np.uint32)
        t = abi.EfxTarget()
        t.struct_size = C.sizeof(abi.EfxTarget)
        t.sample_rate = self.desc.sample_rate
        t.slot_gain = slot_gain
        t.out_channels, t.out_scale, t.out_index = len(out_scale), out_scale.ctypes.data, out_index.ctypes.data
        t.wet_channels, t.wet_index = len(wet_index), wet_index.ctypes.data
        t.real_center, t.real_lfe, t.device_ambi_order = real_center, real_lfe, ambi_order
        props.struct_size = C.sizeof(abi.EfxProps)
        rc = self.m.slot_efx(self.h, slot, C.byref(props), C.byref(t))
        assert rc == expect, f"slot_efx -> {rc}"

    def slot_target(self, slot, target):
        rc = self.m.slot_target(self.h, slot, target)
        assert rc == 0, rc

    def slot_reverb_update(self, slot, params, full, gains):
        params.struct_size = C.sizeof(abi.ReverbParams)
        rc = self.m.slot_reverb_update(self.h, slot, C.byref(params), 1 if full else 0)
        assert rc == 0, rc
        gains = np.ascontiguousarray(gains, dtype=np.float32)
        rc = self.m.slot_output_gains(self.h, slot, 8, gains.ctypes.data)
        assert rc == 0, rc

    def buffer_data(self, buf_id, sample_type, pcm, channels=1):
        pcm = np.ascontiguousarray(pcm)
        rc = self.m.buffer_data(self.h, buf_id, sample_type, channels, pcm.shape[0], pcm.ctypes.data,
                                pcm.nbytes)
        assert rc == 0, rc

    def buffer_data_adpcm(self, buf_id, sample_type, samples_per_block, blocks, data, channels=1):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        rc = self.m.buffer_data_adpcm(self.h, buf_id, sample_type, channels, samples_per_block, blocks,
                                      data.ctypes.data, data.nbytes)
        assert rc == 0, rc

    def voices_update(self, params, coeffs=None, dry=None, send=None):
        n = len(params)
        arr = params if not isinstance(params, list) else (abi.VoiceParams * n)(*params)
        cp = dp = sp = None
        if coeffs is not None:
            coeffs = np.ascontiguousarray(coeffs, dtype=np.float32)
            cp = coeffs.ctypes.data
        if dry is not None:
            dry = np.ascontiguousarray(dry, dtype=np.float32)
            dp = dry.ctypes.data
        if send is not None:
            send = np.ascontiguousarray(send, dtype=np.float32)
            sp = send.ctypes.data
        rc = self.m.voices_update(self.h, n, arr, cp, dp, sp)
        assert rc == 0, rc

    def voices_filters(self, entries):
        """entries: iterable of (voice, path, active, lowpass[5], highpass[5])."""
        entries = list(entries)
        arr = (abi.VoiceFilter * max(len(entries), 1))()
        for i, (v, path, act, lp, hp) in enumerate(entries):
            arr[i].voice, arr[i].path, arr[i].active = int(v), int(path), int(act)
            arr[i].lowpass[:] = [float(x) for x in lp]
            arr[i].highpass[:] = [float(x) for x in hp]
        rc = self.m.voices_filters(self.h, len(entries), arr)
        assert rc == 0, rc

    def render(self, frames=1024, want_results=False):
        ch = self.desc.real_channels
        out = np.zeros((ch, frames), dtype=np.float32)
        ptrs = (C.c_void_p * ch)(*[out[c].ctypes.data for c in range(ch)])
        res = (abi.VoiceResult * max(self.desc.max_voices, 1))() if want_results else None
        rc = self.m.render(self.h, frames, ptrs, res)
        assert rc == 0, f"render -> {rc}"
        return (out, res) if want_results else out

    def voice_queue(self, voice, buffer_ids, loop_index):
        arr = (C.c_uint32 * max(len(buffer_ids), 1))(*buffer_ids)
        rc = self.m.voice_queue(self.h, voice, len(buffer_ids), arr, loop_index)
        assert rc == 0, rc

    OUT_NP = {0: np.int8, 1: np.uint8, 2: np.int16, 3: np.uint16, 4: np.int32, 5: np.uint32, 6: np.float32}

    def render_interleaved(self, frames, out_type, dither_depth, seed, frame_step=None):
        """Returns (out [frames][frame_step], results, new seed)."""
        step = frame_step or self.desc.real_channels
        out = np.zeros((frames, step), dtype=self.OUT_NP[out_type])
        sd = C.c_uint32(seed)
        res = (abi.VoiceResult * max(self.desc.max_voices, 1))()
        rc = self.m.render_interleaved(self.h, frames, out.ctypes.data, out_type, step, dither_depth,
                                       C.byref(sd), res)
        assert rc == 0, f"render_interleaved -> {rc}"
        return out, res, sd.value

    def set_limiter(self, desc):
        """Installs (or with None removes) the output limiter; returns its look-ahead."""
        la = C.c_uint32(0)
        rc = self.m.set_limiter(self.h, C.byref(desc) if desc is not None else None, C.byref(la))
        assert rc == 0, f"set_limiter -> {rc}"
        return la.value

    def set_front_stabilizer(self, center_channel, splitter_coeff):
        rc = self.m.set_front_stabilizer(self.h, center_channel, splitter_coeff)
        assert rc == 0, f"set_front_stabilizer -> {rc}"

    def set_bs2b(self, level):
        rc = self.m.set_bs2b(self.h, level)
        assert rc == 0, f"set_bs2b -> {rc}"

    def set_uhj_encoder(self, filter_length):
        """0 = IIR, 256/512 = FIR; returns the encoder's delay in samples."""
        dl = C.c_uint32(0)
        rc = self.m.set_uhj_encoder(self.h, filter_length, C.byref(dl))
        assert rc == 0, f"set_uhj_encoder -> {rc}"
        return dl.value

    def set_distance_comp(self, delays, gains):
        """Installs per-channel output delays/gains (None removes them)."""
        if delays is None:
            rc = self.m.set_distance_comp(self.h, 0, None, None)
        else:
            dl = np.ascontiguousarray(delays, dtype=np.uint32)
            g = np.ascontiguousarray(gains, dtype=np.float32)
            assert dl.shape == g.shape
            rc = self.m.set_distance_comp(self.h, len(dl), dl.ctypes.data, g.ctypes.data)
        assert rc == 0, f"set_distance_comp -> {rc}"

    def render_begin(self, frames=1024):
        """Returns (wet pointer, float count): a host pointer on the oracle, a device pointer
        on the product."""
        self._frames = frames
        ptr = C.c_void_p()
        cnt = C.c_size_t()
        rc = self.m.render_begin(self.h, frames, C.byref(ptr), C.byref(cnt))
        assert rc == 0, f"render_begin -> {rc}"
        return ptr.value, cnt.value

    def render_end(self):
        ch = self.desc.real_channels
        out = np.zeros((ch, self._frames), dtype=np.float32)
        ptrs = (C.c_void_p * ch)(*[out[c].ctypes.data for c in range(ch)])
        rc = self.m.render_end(self.h, ptrs, None, None)
        assert rc == 0, f"render_end -> {rc}"
        return out

    def shard_init(self, rank, world):
        """Returns this device's 64-byte IPC handle (peer-store transport)."""
        buf = C.create_string_buffer(64)
        rc = self.m.shard_init(self.h, rank, world, buf)
        assert rc == 0, f"shard_init -> {rc}: {self.last_error()}"
        return buf.raw

    def shard_connect(self, handles):
        blob = b"".join(handles)
        rc = self.m.shard_connect(self.h, blob)
        assert rc == 0, f"shard_connect -> {rc}: {self.last_error()}"

    def last_error(self):
        fn = getattr(self.m.lib, self.m.prefix + "last_error")
        fn.restype = C.c_char_p
        fn.argtypes = [C.c_void_p]
        return (fn(self.h) or b"").decode()

    def dry(self):
        out = np.zeros((self.desc.dry_channels, abi.LINE), dtype=np.float32)
        self.m.get_dry(self.h, out.ctypes.data)
        return out


_oracle = None


def oracle() -> MixLib:
    global _oracle
    if _oracle is None:
        _oracle = MixLib(ORACLE_SO, "oracle_")
    return _oracle


_product = None


def product() -> MixLib:
    global _product
    if _product is None:
        _product = MixLib(PRODUCT_SO, "b200mix_")
    return _product
