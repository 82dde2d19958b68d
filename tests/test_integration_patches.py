"""The reference-side change set under integration/ stays applicable: the patches apply cleanly to
the reference tree they were written against (skipped where /root/reference is absent — the GPU
box), and they are as small as INTEGRATION.md says: includes plus six call sites."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCHES = {"alu_seam.patch": "alc/alu.cpp", "convolution_seam.patch": "alc/effects/convolution.cpp",
           "device_seam.patch": "core/device.cpp", "alc_seam.patch": "alc/alc.cpp"}


@pytest.mark.parametrize("name", sorted(PATCHES))
def test_patch_applies_to_the_reference(name, tmp_path):
    src = os.path.join(REF, PATCHES[name])
    if not os.path.exists(src):
        pytest.skip("reference tree absent")
    out = os.path.join(str(tmp_path), "patched.cpp")
    p = subprocess.run(["patch", "-s", "-o", out, src, os.path.join(ROOT, "integration", name)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    text = open(out).read()
    assert '#include "b200mix_seam.h"' in text
    assert text.count("b200seam_") == {"alu_seam.patch": 4, "convolution_seam.patch": 1, "device_seam.patch": 1,
                                        "alc_seam.patch": 1}[name]


def test_patches_touch_only_a_handful_of_lines():
    added = 0
    for name in PATCHES:
        for line in open(os.path.join(ROOT, "integration", name)):
            if line.startswith("+") and not line.startswith("+++"):
                added += 1
    assert added <= 28, added
