import sys; sys.path.insert(0,"tests"); sys.path.insert(0,"openal-soft_b200")
import numpy as np, ctypes as C, time
from helpers import mixlib, synth
from helpers.mixlib import MixDevice
from pyb200mix import abi, scene
rng=np.random.default_rng(21); nv, ir = 16, 64
desc=synth.hrtf_desc(nv, ir)
params, coeffs, dry = synth.voice_set(rng, nv, ir)
for p in params: p.hrtf_gain*=12.0
dev=MixDevice(mixlib.product(), desc)
dev.set_hrtf_decoder(*synth.decoder(np.random.default_rng(7)))
for i in range(nv): dev.buffer_data(i, abi.FMT_I16, scene.voice_buffer_fast(i))
dev.voices_update(params, coeffs, dry, None)
def t(label):
    for r in range(3):
        t0=time.perf_counter()
        for k in range(100): dev.render(1024)
        print(label, "ms/update", (time.perf_counter()-t0)/100*1e3)
t("no limiter")
dev.set_limiter(abi.device_limiter(-0.00053)); t("device limiter")
dev.close()
