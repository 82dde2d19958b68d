#!/usr/bin/env python3
"""bench.py — the per-update mixing hot path on B200 (BASELINE.json metric:
"real-time HRTF voices @48kHz/1024-sample update; samples/sec mixed").

A step = ONE 1024-frame mix update of the whole voice set (the voice loop of
DeviceBase::renderSamples, alc/alu.cpp:2412) over synthetic 48 kHz mono voices.
N=1 workload = BASELINE config 2: 4096 mono voices, HRTF (64-tap HRIR pair per voice),
bsinc24, pitch in [0.5, 2.0) with 1/16 at 1.0.

  python bench.py --gpus N --steps K --warmup W            # the CUDA mixer (libb200mix.so)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU mixer

value  : voice-samples/s with everything resident in HBM, device-timed (CUDA events on the
         mixer's stream around each update, L2 flushed between updates), max over ranks.
e2e    : same metric through the C ABI with HOST buffers: per step the parameter
         snapshots of 1/8 of the voices (moving sources: new HRIR + delays + gain) go
         host->device, the planar output block and the per-voice results come back.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "openal-soft_b200"))
from pyb200mix import abi, scene  # noqa: E402

VOICES_PER_GPU = 4096
IR = 64
FRAMES = 1024
UPDATE_MS = 1000.0 * FRAMES / 48000.0
L2_FLUSH_BYTES = 256 << 20


# --------------------------------------------------------------------------- helpers
def load_product():
    path = os.path.join(ROOT, "openal-soft_b200", "libb200mix.so")
    if not os.path.exists(path):
        raise SystemExit("libb200mix.so missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    lib.b200mix_create.argtypes = [C.POINTER(abi.DeviceDesc), C.POINTER(C.c_void_p)]
    lib.b200mix_destroy.argtypes = [C.c_void_p]
    lib.b200mix_last_error.restype = C.c_char_p
    lib.b200mix_last_error.argtypes = [C.c_void_p]
    lib.b200mix_set_hrtf_decoder.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32] + [C.c_void_p] * 3
    lib.b200mix_buffer_data.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.c_void_p, C.c_size_t]
    lib.b200mix_voices_update.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 4
    lib.b200mix_render.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.c_void_p]
    lib.b200mix_render_device.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.b200mix_profile.argtypes = [C.c_void_p, C.c_int]
    lib.b200mix_last_mix_kernel_ms.restype = C.c_float
    lib.b200mix_last_mix_kernel_ms.argtypes = [C.c_void_p]
    lib.b200mix_launch_count.restype = C.c_uint64
    lib.b200mix_launch_count.argtypes = [C.c_void_p]
    lib.b200mix_stream.restype = C.c_void_p
    lib.b200mix_stream.argtypes = [C.c_void_p]
    lib.b200mix_hrtf_load.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.b200mix_hrtf_get_coeffs.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                            C.c_void_p, C.POINTER(C.c_uint32)]
    lib.b200mix_hrtf_attach.argtypes = [C.c_void_p, C.c_void_p]
    lib.b200mix_voices_update_dirs.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 4
    return lib


MHR_PATH = os.path.join(ROOT, "openal-soft_b200", "data", "Default HRTF.mhr")


def load_hrtf(lib):
    """The reference's default data set through the product's own MHR loader, or None."""
    if not os.path.exists(MHR_PATH):
        return None
    data = open(MHR_PATH, "rb").read()
    h = C.c_void_p()
    return h if lib.b200mix_hrtf_load(data, len(data), C.byref(h)) == 0 else None


def direction_of(pos):
    """{elevation, azimuth, distance, spread} as CalcHrtfPanning derives them for a source
    at `pos` (alc/alu.cpp:1210-1216)."""
    x, y, z = pos
    d = math.sqrt(x * x + y * y + z * z)
    ev = math.asin(max(-1.0, min(1.0, y / d)))
    az = math.atan2(x / d, -z / d)
    return ev, az, d, 0.0


def hrir_for(lib, hrtf, pos, out, delays):
    """HrtfStore::getCoeffs on the host for a source at `pos`."""
    ev, az, d, sp = direction_of(pos)
    lib.b200mix_hrtf_get_coeffs(hrtf, ev, az, d, sp, out.ctypes.data, delays)


def synth_voices(first, count, total, lib=None, hrtf=None, shift=0.0):
    """Post-ALU parameter snapshots for voices [first, first+count) of a `total`-voice
    scene (SURVEY §8d positions): HRIR pair + delays from Default HRTF.mhr via the product's
    HrtfStore::getCoeffs restatement (synthetic decaying filters only if the data set is
    not staged), gain 1/sqrt(total).  `shift` rotates the azimuths (moving sources)."""
    rng = np.random.default_rng(0xB200 + first)
    coeffs = (rng.standard_normal((count, IR, 2)) * np.exp(-np.arange(IR) / 10.0)[None, :, None]
              ).astype(np.float32)
    params = (abi.VoiceParams * count)()
    dl = (C.c_uint32 * 2)()
    pitches = []
    for k in range(count):
        i = first + k
        p = params[k]
        p.voice = k
        p.flags = abi.VF_PLAYING | abi.VF_STATIC | abi.VF_LOOPING | abi.VF_HRTF | abi.VF_RESET
        p.buffer = k
        p.resampler = abi.RS_BSINC24
        p.position = 0
        p.position_frac = 0
        p.loop_start = 0
        p.loop_end = scene.BUFFER_FRAMES
        pitch = scene.voice_pitch(i)
        pitches.append(pitch)
        p.step = max(1, min(int(pitch * 65536.0), 10 << 16))
        p.hrtf_delay[0] = int(rng.integers(0, 40))
        p.hrtf_delay[1] = int(rng.integers(0, 40))
        if hrtf is not None:
            x, y, z = scene.voice_position(i)
            if shift:
                cs, sn = math.cos(shift), math.sin(shift)
                x, z = x * cs - z * sn, x * sn + z * cs
            hrir_for(lib, hrtf, (x, y, z), coeffs[k], dl)
            p.hrtf_delay[0], p.hrtf_delay[1] = dl[0], dl[1]
        p.hrtf_gain = scene.voice_gain(total)
        for s in range(abi.MAX_SENDS):
            p.send_slot[s] = abi.NO_SLOT
    return params, coeffs, np.array(pitches)


def algorithmic_bytes_per_voice(mean_pitch):
    """source 1024*p*2 B + mPrevSamples R+W 2*192 + pos/frac/step 16
    + HRTF history R+W 2*256 + target coeffs Ir*8 + delays/gain 12."""
    return 1024.0 * mean_pitch * 2 + 2 * 192 + 16 + 2 * 256 + IR * 8 + 12


class ClockSampler:
    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.idx)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t.start()

    def stop(self):
        self._stop.set()
        self._t.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------- reference arm
def _ref_worker(first, count, total, steps, warmup, conn):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import refal
    dev = refal.RefDevice({refal.ALC_HRTF_SOFT: 1, refal.ALC_MONO_SOURCES: max(count, 1)})
    assert dev.hrtf_enabled()
    for k in range(count):
        i = first + k
        dev.add_voice(scene.voice_buffer_fast(i), scene.BUFFER_RATE, scene.voice_pitch(i),
                      scene.voice_position(i), scene.voice_gain(total), abi.RS_BSINC24)
    dev.play_all()
    for _ in range(warmup):
        dev.render()
    conn.send("ready")
    conn.recv()
    t0 = time.perf_counter()
    for _ in range(steps):
        dev.render()
    dt = time.perf_counter() - t0
    conn.send(dt)
    dev.close()


def run_reference(total_voices, steps, warmup, procs):
    """The reference's own SSE mixer (oracle/_ref/libopenal_ref.so through the loopback
    API), one independent loopback device per process with the voices split evenly (the
    reference mixer is single-threaded per device, core/device.h:420-421)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    per = [total_voices // procs + (1 if r < total_voices % procs else 0) for r in range(procs)]
    workers = []
    first = 0
    for r in range(procs):
        a, b = ctx.Pipe()
        pr = ctx.Process(target=_ref_worker, args=(first, per[r], total_voices, steps, warmup, b))
        pr.start()
        workers.append((pr, a))
        first += per[r]
    for _, a in workers:
        assert a.recv() == "ready"
    for _, a in workers:
        a.send("go")
    times = [a.recv() for _, a in workers]
    for pr, _ in workers:
        pr.join()
    return max(times)


def reference_available():
    return (os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libopenal_ref.so"))
            and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")))


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not reference_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built"}))
        return
    cores = os.cpu_count() or 1
    procs = max(1, cores)
    # bounded sample of the same workload: `sample_voices` of the 4096*N voices, so that
    # the whole run ends within minutes (~30 us per voice-update per core)
    total = VOICES_PER_GPU * args.gpus
    sample = min(total, 32 * procs)
    dt = run_reference(sample, args.steps, args.warmup, procs)
    value = sample * FRAMES * args.steps / dt
    line = {
        "impl": "reference", "metric": "voice-samples/s mixed (HRTF, bsinc24, 1024-frame updates)",
        "value": value, "unit": "voice-samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config2: mono 48k voices, Default HRTF, bsinc24, pitch U[0.5,2)",
                   "voices_mixed": sample, "update_frames": FRAMES},
        "cpu_baseline": {"value": value, "unit": "voice-samples/s", "cores": procs, "kind": "reference",
                         "sample": f"{sample} of {total} voices x {args.steps} updates, "
                                   f"{procs} independent loopback devices (1 per core)"},
        "e2e": {"value": value, "unit": "voice-samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "rt_voices": sample * UPDATE_MS / (1000.0 * dt / args.steps),
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- CUDA arm
def main_cuda(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the b200mix mixer has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load_product()

    total = VOICES_PER_GPU * world
    first = VOICES_PER_GPU * rank
    nv = VOICES_PER_GPU
    desc = abi.DeviceDesc()
    desc.struct_size = C.sizeof(abi.DeviceDesc)
    desc.cuda_device = local
    desc.sample_rate = 48000
    desc.dry_channels = 4
    desc.real_channels = 2
    desc.ir_size = IR
    desc.post_process = abi.POST_HRTF
    desc.real_left, desc.real_right = 0, 1
    desc.max_voices = nv
    desc.max_buffers = nv
    h = C.c_void_p()
    rc = lib.b200mix_create(C.byref(desc), C.byref(h))
    if rc != 0:
        raise SystemExit(f"b200mix_create failed: {lib.b200mix_last_error(None)}")

    def ck(rc, what):
        if rc != 0:
            raise SystemExit(f"{what} failed ({rc}): {lib.b200mix_last_error(h)}")

    rng = np.random.default_rng(7)
    dec = (rng.standard_normal((4, 91, 2)) * np.exp(-np.arange(91) / 12.0)[None, :, None] * 0.2).astype(np.float32)
    hf = np.array([2.0, 1.1547005, 1.1547005, 1.1547005], dtype=np.float32)
    sc = np.full(4, -0.9123257, dtype=np.float32)
    ck(lib.b200mix_set_hrtf_decoder(h, 4, 91, dec.ctypes.data, hf.ctypes.data, sc.ctypes.data), "set_hrtf_decoder")
    for k in range(nv):
        pcm = scene.voice_buffer_fast(first + k)
        ck(lib.b200mix_buffer_data(h, k, abi.FMT_I16, 1, pcm.shape[0], pcm.ctypes.data, pcm.nbytes), "buffer_data")
    hrtf = load_hrtf(lib)
    params, coeffs, pitches = synth_voices(first, nv, total, lib, hrtf)
    ck(lib.b200mix_voices_update(h, nv, params, coeffs.ctypes.data, None, None), "voices_update")

    stream = torch.cuda.ExternalStream(lib.b200mix_stream(h))
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    out_ptr = C.c_void_p()

    def step_device():
        ck(lib.b200mix_render_device(h, FRAMES, C.byref(out_ptr)), "render_device")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed value -------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    lib.b200mix_profile(h, 1)
    launches0 = lib.b200mix_launch_count(h)
    step_ms, mix_ms = [], []
    reduce_buf = torch.zeros(2 * FRAMES, dtype=torch.float32, device="cuda") if world > 1 else None
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                       # evict the voice state / sources from L2
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device()
        if world > 1:
            # the single per-update collective: sum of the per-GPU RealOut blocks (the
            # post-process is linear, so reducing after it equals reducing Dry/Accum)
            src = _as_tensor(out_ptr.value, 2 * FRAMES, local)
            with torch.cuda.stream(stream):
                reduce_buf.copy_(src)
                dist.reduce(reduce_buf, dst=0)
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        mix_ms.append(lib.b200mix_last_mix_kernel_ms(h))
    barrier()
    launches = lib.b200mix_launch_count(h) - launches0
    lib.b200mix_profile(h, 0)
    t_local = float(np.sum(step_ms))
    if world > 1:
        tt = torch.tensor([t_local], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_total = float(tt.item())
    else:
        t_total = t_local
    ms_per_step = t_total / args.steps
    value = total * FRAMES / (ms_per_step * 1e-3)

    # ---- end-to-end through the C ABI with host buffers ----------------------------
    out = np.zeros((2, FRAMES), dtype=np.float32)
    ptrs = (C.c_void_p * 2)(out[0].ctypes.data, out[1].ctypes.data)
    results = (abi.VoiceResult * nv)()
    nmove = nv // 8
    h2d = nmove * (C.sizeof(abi.VoiceParams) + IR * 2 * 4)
    d2h = 2 * FRAMES * 4 + nv * C.sizeof(abi.VoiceResult)
    # the application's per-update work (new positions -> new HRIRs) is prepared up
    # front: 8 rotating sets, each moving a different eighth of the voices
    move_sets = []
    for base in range(8):
        # the moved eighth gets the HRIRs of a rotated position (new coefficients AND delays)
        p2, c2, _ = synth_voices(first, nv, total, lib, hrtf, shift=0.05 * (base + 1))
        mp = (abi.VoiceParams * nmove)()
        for j in range(nmove):
            k = base + 8 * j
            C.memmove(C.byref(mp[j]), C.byref(p2[k]), C.sizeof(abi.VoiceParams))
            mp[j].flags &= ~abi.VF_RESET
            if hrtf is None:
                mp[j].hrtf_delay[0] = (params[k].hrtf_delay[0] + 3 * base + 1) % 40
        mc = np.ascontiguousarray(c2[base::8][:nmove] * np.float32(1.0 if hrtf is not None else 1.0 - 0.02 * base))
        md = np.zeros((nmove, 4), dtype=np.float32)
        cs, sn = math.cos(0.05 * (base + 1)), math.sin(0.05 * (base + 1))
        for j in range(nmove):
            x, y, z = scene.voice_position(first + base + 8 * j)
            md[j] = direction_of((x * cs - z * sn, y, x * sn + z * cs))
        move_sets.append((mp, mc, md))

    # With the data set attached the application only sends the moved sources' DIRECTIONS
    # and the 4-HRIR blend runs on the GPU (b200mix_voices_update_dirs);
    # otherwise it sends host-blended HRIRs.
    use_dirs = hrtf is not None and lib.b200mix_hrtf_attach(h, hrtf) == 0
    if use_dirs:
        h2d = nmove * (C.sizeof(abi.VoiceParams) + 16)

    def step_e2e(it):
        mp, mc, md = move_sets[it % 8]
        if use_dirs:
            ck(lib.b200mix_voices_update_dirs(h, nmove, mp, md.ctypes.data, None, None), "voices_update_dirs")
        else:
            ck(lib.b200mix_voices_update(h, nmove, mp, mc.ctypes.data, None, None), "voices_update")
        ck(lib.b200mix_render(h, FRAMES, ptrs, results), "render")

    for it in range(args.warmup):
        step_e2e(it)
    barrier()
    t0 = time.perf_counter()
    for it in range(args.steps):
        step_e2e(args.warmup + it)
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([te], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
    e2e_value = total * FRAMES * args.steps / te
    clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"
        mean_pitch = float(np.mean(pitches))
        alg_bytes = algorithmic_bytes_per_voice(mean_pitch) * nv
        mix_avg = float(np.mean([m for m in mix_ms if m > 0])) if any(m > 0 for m in mix_ms) else None
        achieved = alg_bytes / (mix_avg * 1e-3) / 1e9 if mix_avg else None
        traffic = None
        try:
            traffic = float(json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                            ["k_mix_voices"]["dram_bytes_per_launch"])
        except Exception:
            pass
        line = {
            "metric": "voice-samples/s mixed (HRTF, bsinc24, 1024-frame updates)",
            "value": value, "unit": "voice-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config2: 4096 mono 48k voices per GPU, 64-tap HRIR pair per voice, "
                                   "bsinc24, pitch U[0.5,2) (1/16 at 1.0)",
                       "voices": total, "voices_per_gpu": nv, "update_frames": FRAMES,
                       "hrir": ("Default HRTF.mhr (MinPHR03, 48 kHz, Ir=64) via b200mix_hrtf_get_coeffs"
                                if hrtf is not None else "synthetic decaying 64-tap pairs (data set not staged)"),
                       "l2": "flushed between timed updates (256 MiB memset)",
                       "parallelism": f"voices sharded over {world} GPU(s); one NCCL reduce of RealOut per update"},
            "rt_voices": total * UPDATE_MS / ms_per_step,
            "e2e": {"value": e2e_value, "unit": "voice-samples/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": 1000.0 * te / args.steps,
                    "update": ("b200mix_voices_update_dirs (directions; HRIR blend on the GPU)" if use_dirs
                               else "b200mix_voices_update (host-blended HRIRs)") + f", {nmove} moved voices/GPU/update"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "k_mix_voices", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (dram__bytes_read+write "
                                                               "of one ncu --set full capture of this workload)",
                         "peak_source": peak_src,
                         "kernel_ms": mix_avg, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "HRTF voices are FP32-FMA/shared-memory bound (~100 flop/B), not "
                                 "HBM bound; see DESIGN.md"},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    lib.b200mix_destroy(h)
    if world > 1:
        dist.destroy_process_group()


def _as_tensor(ptr, count, device_index):
    import torch

    class _Wrap:
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(w, device=torch.device("cuda", device_index))


def cpu_baseline():
    """The reference's SSE mixer on ONE host core (it is single-threaded per device by
    design), bounded sample: 512 of the 4096 voices x 24 updates."""
    if not reference_available():
        return {"value": None, "unit": "voice-samples/s", "cores": 1, "kind": "reference",
                "sample": "unavailable: oracle/_ref not built"}
    sample, steps = 512, 24
    dt = run_reference(sample, steps, 4, 1)
    return {"value": sample * FRAMES * steps / dt, "unit": "voice-samples/s", "cores": 1,
            "kind": "reference",
            "sample": f"{sample} of 4096 voices x {steps} updates, 1 loopback device, SSE4.1 kernels",
            "ms_per_voice_update": 1000.0 * dt / (sample * steps)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_cuda(args)


if __name__ == "__main__":
    main()
