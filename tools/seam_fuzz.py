#!/usr/bin/env python3
"""Runs seeded random API sequences (tests/helpers/al_runner.py, scene "fuzzN") on the stock
reference and on the patched library behind the oracle shim, and compares audio, source states and
offsets — the check of tests/test_seam_cpu.py for any number of seeds.  No GPU needed.

    python tools/seam_fuzz.py 0 40            # seeds 0..39: the basic set of calls
    python tools/seam_fuzz.py 100 140         # from 100: effect swaps, deferred updates, new source formats ...
    python tools/seam_fuzz.py 200 220         # from 200: the same in ragged update sizes
    python tools/seam_fuzz.py 7 8 --updates 30 --diagnose

--diagnose prints the peak error of every update and the first update whose source states or
offsets differ; AL_RUNNER_FUZZ_LOG=1 lists the calls, AL_RUNNER_FUZZ_SKIP=u:k,... /
AL_RUNNER_FUZZ_ONLY=u:k,... mask single calls (update u, k-th call) to bisect a failure."""
import argparse
import os
import pathlib
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "openal-soft_b200"))
import test_seam_cpu as seam  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("first", type=int)
    ap.add_argument("last", type=int, help="exclusive")
    ap.add_argument("--voices", type=int, default=24)
    ap.add_argument("--updates", type=int, default=30)
    ap.add_argument("--diagnose", action="store_true")
    a = ap.parse_args()
    failed = 0
    for seed in range(a.first, a.last):
        for hrtf in (1, 0):
            fx = f"fuzz{seed}"
            with tempfile.TemporaryDirectory() as d:
                d = pathlib.Path(d)
                if a.diagnose:
                    cpu = seam._run("libopenal_ref.so", "cpu", a.voices, a.updates, hrtf, False, d, fx)
                    via = seam._run("libopenal_b200.so", "seam", a.voices, a.updates, hrtf, True, d, fx)
                    err = np.abs(via["out"].astype(np.float64) - cpu["out"]).max(axis=(1, 2))
                    print(f"seed {seed} hrtf {hrtf} peak error per update:", " ".join(f"{e:.1e}" for e in err))
                    for key in ("states", "offsets"):
                        for u in range(a.updates):
                            if not np.array_equal(cpu[key][u], via[key][u]):
                                w = np.nonzero(cpu[key][u] != via[key][u])[0]
                                print(f"  {key} differ first at update {u}: columns {w}, reference {cpu[key][u][w]}, seam {via[key][u][w]}")
                                break
                    continue
                try:
                    seam.test_seam_drives_the_abi_like_the_stock_mixer(a.voices, a.updates, hrtf, fx, d)
                    print(f"ok   seed {seed} hrtf {hrtf}", flush=True)
                except BaseException as e:  # noqa: BLE001  (pytest.skip and assertion errors alike)
                    failed += 1
                    print(f"FAIL seed {seed} hrtf {hrtf}: {str(e)[-600:]}", flush=True)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
