#!/usr/bin/env python3
"""Writes the build-configuration headers the reference's sources expect
(config.h, config_simd.h, config_backends.h, version.h, default_hrtf.hpp) into
oracle/_ref/gen/.  These are OUR statements of this container's configuration
(Linux x86-64, GCC, SSE..SSE4.1, loopback/null/wave backends only) — the
reference's own build system (cmake) is not run.  TEST INFRASTRUCTURE ONLY.
usage: gen_headers.py <reference_root> <outdir>
"""
import sys, os

ref, out = sys.argv[1], sys.argv[2]
os.makedirs(out, exist_ok=True)

def w(name, text):
    p = os.path.join(out, name)
    old = open(p).read() if os.path.exists(p) else None
    if old != text:
        open(p, "w").write(text)

w("config.h", """#pragma once
#define FORCE_ALIGN
#define ALSOFT_EMBED_HRTF_DATA
#define HAVE_DLFCN_H
#define HAVE_CPUID_H
#define HAVE_GCC_GET_CPUID
#define HAVE_PTHREAD_SETSCHEDPARAM
#define HAVE_PTHREAD_SETNAME_NP
#define HAVE_CXXMODULES 0
#define HAVE_DYNLOAD 1
#define HAVE_RTKIT 0
#define ALSOFT_UWP 0
#define ALSOFT_EAX 0
""")
w("config_simd.h", """#pragma once
#define HAVE_SSE 1
#define HAVE_SSE2 1
#define HAVE_SSE3 1
#define HAVE_SSE4_1 1
#define HAVE_SSE_INTRINSICS 1
#define HAVE_NEON 0
""")
backends = ["ALSA", "OSS", "PIPEWIRE", "SOLARIS", "SNDIO", "WASAPI", "DSOUND", "WINMM",
            "PORTAUDIO", "PULSEAUDIO", "JACK", "COREAUDIO", "OPENSL", "OBOE", "SDL3", "SDL2"]
w("config_backends.h", "#pragma once\n" + "".join(f"#define HAVE_{b} 0\n" for b in backends)
  + "#define HAVE_WAVE 1\n")
w("version.h", """#pragma once
#define ALSOFT_VERSION "1.25.2"
#define ALSOFT_VERSION_NUM 1,25,2,0
#define ALSOFT_GIT_BRANCH "oracle"
#define ALSOFT_GIT_COMMIT_HASH "2dc741b5"
""")
# Embedded default HRTF data set: a byte array of hrtf/Default HRTF.mhr.
data = open(os.path.join(ref, "hrtf", "Default HRTF.mhr"), "rb").read()
rows = []
for i in range(0, len(data), 24):
    rows.append(",".join("'\\x%02x'" % b for b in data[i:i + 24]))
w("default_hrtf.hpp", "#pragma once\nconstexpr char default_hrtf[] = {\n" + ",\n".join(rows) + "\n};\n")
print("generated headers in", out)
