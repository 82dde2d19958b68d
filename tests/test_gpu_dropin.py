"""The drop-in, end to end: the SAME application code (tests/helpers/al_runner.py: public
AL/ALC API only — alBufferData, alSourcePlay, moving / stopped / restarted sources, EFX effect
slots with property / gain / target / type changes, alcRenderSamplesSOFT) runs on (a) the stock compiled reference and (b) libopenal_b200.so — the
reference with integration/alu_seam.patch (3 call sites in alc/alu.cpp) and the binding
integration/b200mix_seam.cpp, ALSOFT_B200MIX=1 — whose alcRenderSamplesSOFT mixes on the GPU
through libb200mix.so.  Audio must agree within north_star's budget (held 10x tighter), and the
source states / offsets the AL API reports must be identical."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
RUNNER = os.path.join(ROOT, "tests", "helpers", "al_runner.py")

pytestmark = pytest.mark.gpu


def _run(lib, tag, voices, updates, hrtf, gpu, tmp_path, fx="none"):
    out = os.path.join(str(tmp_path), f"{tag}.npz")
    env = dict(os.environ)
    env.pop("ALSOFT_B200MIX", None)
    if gpu:
        env["ALSOFT_B200MIX"] = "1"
        env["ALSOFT_B200MIX_LIB"] = os.path.join(ROOT, "openal-soft_b200", "libb200mix.so")
    p = subprocess.run([sys.executable, RUNNER, os.path.join(REF, lib), out, str(voices), str(updates), str(hrtf), "7", fx],
                       env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "b200mix:" not in p.stderr, p.stderr[-2000:]         # the seam's own error lines (device disconnected)
    return dict(np.load(out))


def _need():
    for f in ("libopenal_ref.so", "libopenal_b200.so"):
        if not os.path.exists(os.path.join(REF, f)):
            pytest.skip(f"oracle/_ref/{f} not built (python -c 'import __graft_entry__ as g; g.build()')")


@pytest.mark.parametrize("voices,updates,hrtf,fx", [(24, 12, 1, "none"), (24, 8, 0, "none"), (4096, 6, 1, "none"),
                                                    (24, 8, 1, "reverb"), (24, 8, 1, "mix"), (24, 8, 0, "mix"), (24, 8, 1, "filt"),
                                                    (24, 8, 0, "mixfilt"), (24, 8, 1, "stream"), (24, 8, 0, "stream"),
                                                    (24, 8, 1, "stereo"), (24, 8, 0, "stereo"), (24, 8, 1, "conv"), (24, 8, 0, "conv"), (24, 8, 1, "reset"), (24, 8, 0, "reset"), (12, 6, 1, "bformat"), (12, 6, 0, "bformat"),
                                                    (2048, 6, 1, "mix")])
def test_patched_reference_renders_through_libb200mix(voices, updates, hrtf, fx, tmp_path):
    _need()
    cpu = _run("libopenal_ref.so", "cpu", voices, updates, hrtf, False, tmp_path, fx)
    gpu = _run("libopenal_b200.so", "gpu", voices, updates, hrtf, True, tmp_path, fx)
    # ("reset" toggles HRTF with alcResetDeviceSOFT half-way)
    assert int(cpu["hrtf_status"]) == int(gpu["hrtf_status"]) == ((1 if hrtf else 0) ^ (1 if fx == "reset" else 0))
    ref, out = cpu["out"].astype(np.float64), gpu["out"].astype(np.float64)
    assert np.abs(ref).max() > 1e-2
    err = out - ref
    rms, mx = float(np.sqrt((err ** 2).mean())), float(np.abs(err).max())
    # effect scenes: 4x looser (still inside north_star) — the chorus LFO and the equalizer's shelves
    # amplify last-bit differences (see tests/test_gpu_parity.py EFX cases)
    k = 1.0 if fx == "none" else 4.0
    assert rms <= k * 1e-6 and mx <= k * 1e-5, f"rms {rms:.3e} max {mx:.3e}"
    # what the application sees through alGetSourcei: play states and sample offsets
    assert np.array_equal(cpu["states"], gpu["states"])
    assert np.array_equal(cpu["offsets"], gpu["offsets"])


def test_patched_library_without_the_switch_is_the_stock_mixer(tmp_path):
    """ALSOFT_B200MIX unset: libopenal_b200.so mixes on the CPU, bit for bit like the reference."""
    _need()
    a = _run("libopenal_ref.so", "a", 8, 3, 1, False, tmp_path)
    b = _run("libopenal_b200.so", "b", 8, 3, 1, False, tmp_path)
    assert np.array_equal(a["out"], b["out"])
