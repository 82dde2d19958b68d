"""The product's C-ABI library loads and exports every symbol include/b200mix.h declares
(no compute calls: this runs without a GPU), and fails loudly when no CUDA device exists."""
import ctypes as C
import os
import re

import pytest

from helpers import mixlib
from pyb200mix import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200mix.h")).read()
    return sorted(set(re.findall(r"B200MIX_API\s+[\w\s\*]+?\b(b200mix_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for need in ["b200mix_create", "b200mix_destroy", "b200mix_buffer_data", "b200mix_voices_update",
                 "b200mix_render", "b200mix_render_device", "b200mix_set_hrtf_decoder",
                 "b200mix_set_ambi_decoder", "b200mix_last_error", "b200mix_launch_count"]:
        assert need in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(mixlib.PRODUCT_SO), "libb200mix.so not built (run __graft_entry__.build())"
    lib = C.CDLL(mixlib.PRODUCT_SO)
    for name in _declared():
        assert hasattr(lib, name), name


def test_struct_sizes_match_the_header():
    assert C.sizeof(abi.DeviceDesc) == 14 * 4
    assert C.sizeof(abi.VoiceParams) == 18 * 4
    assert C.sizeof(abi.VoiceResult) == 16


def test_create_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    from helpers import synth
    lib = C.CDLL(mixlib.PRODUCT_SO)
    lib.b200mix_create.argtypes = [C.POINTER(abi.DeviceDesc), C.POINTER(C.c_void_p)]
    lib.b200mix_last_error.restype = C.c_char_p
    lib.b200mix_last_error.argtypes = [C.c_void_p]
    h = C.c_void_p()
    d = synth.hrtf_desc(4)
    rc = lib.b200mix_create(C.byref(d), C.byref(h))
    assert rc == -2 and not h.value          # B200MIX_ERR_CUDA: there is no CPU fallback
    assert b"CUDA" in lib.b200mix_last_error(None)
