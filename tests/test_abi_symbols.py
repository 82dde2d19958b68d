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


def test_header_is_plain_c_and_cpp():
    """include/b200mix.h must compile on its own as C99 and as C++ without warnings (it is the
    interface a C host or a cgo/JNI/ctypes binding reads)."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "b200mix.h")
    gcc, gxx = shutil.which("gcc"), shutil.which("g++")
    if not gcc or not gxx:
        pytest.skip("no host compiler")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call([gxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", hdr])


def test_c_host_can_call_the_parameter_stage():
    """A plain C program (tools/bench_param_stage.c) builds against the header and the library and
    runs the host parameter stage."""
    import shutil
    import subprocess
    import tempfile
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no host compiler")
    libdir = os.path.dirname(mixlib.PRODUCT_SO)
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "bps")
        subprocess.check_call([gcc, "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tools", "bench_param_stage.c"), "-L", libdir, "-lb200mix", "-lm",
                               "-o", exe])
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = libdir + os.pathsep + env.get("LD_LIBRARY_PATH", "")
        out = subprocess.check_output([exe], env=env, text=True)
        assert "sources/s" in out


def test_standalone_c_example_builds_and_fails_loudly_without_a_gpu():
    """examples/standalone_host.c (a host using only the library) compiles against the header as C,
    sets up its HRTF device from the data set and — on this CPU-only box — stops at b200mix_create."""
    import shutil
    import subprocess
    import tempfile
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    gcc = shutil.which("gcc")
    mhr = os.path.join(ROOT, "openal-soft_b200", "data", "Default HRTF.mhr")
    if not gcc or not os.path.exists(mhr):
        pytest.skip("no host compiler / data set")
    libdir = os.path.dirname(mixlib.PRODUCT_SO)
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "standalone_host")
        subprocess.check_call([gcc, "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "examples", "standalone_host.c"), "-L", libdir, "-lb200mix", "-lm",
                               "-o", exe])
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = libdir + os.pathsep + env.get("LD_LIBRARY_PATH", "")
        out = subprocess.check_output([exe, mhr], env=env, text=True)
        assert "no CPU path" in out
