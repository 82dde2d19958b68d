"""Host-side filter parameter helper of the product (b200mix_biquad_coeffs, no GPU involved)
against the oracle restatement of BiquadFilter::SetParams, which is itself pinned
bit-for-bit against the compiled reference in test_oracle_vs_ref.py."""
import ctypes as C

import numpy as np

from helpers import mixlib


def test_biquad_coeffs_helper_matches_oracle_bit_exact():
    prod = mixlib.product()
    orc = mixlib.oracle()
    rng = np.random.default_rng(11)
    for _ in range(2000):
        typ = int(rng.integers(0, 6))
        f0 = float(rng.uniform(1e-4, 0.6))
        gain = float(10.0 ** rng.uniform(-6.0, 1.0))
        slope = float(rng.uniform(0.05, 1.0))
        a = np.zeros(5, dtype=np.float32)
        b = np.zeros(5, dtype=np.float32)
        assert orc.biquad_coeffs(typ, f0, gain, slope, a.ctypes.data) == 0
        assert prod.biquad_coeffs(typ, f0, gain, slope, b.ctypes.data) == 0
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (typ, f0, gain, slope)


def test_biquad_coeffs_rejects_bad_arguments():
    prod = mixlib.product()
    out = np.zeros(5, dtype=np.float32)
    assert prod.biquad_coeffs(6, 0.1, 1.0, 1.0, out.ctypes.data) < 0
    assert prod.biquad_coeffs(0, 0.1, 1.0, 0.0, out.ctypes.data) < 0
    assert prod.biquad_coeffs(0, 0.1, 1.0, 1.0, None) < 0


def test_unit_gain_shelf_is_transparent():
    """gain 1 gives b == a (numerator equals denominator): the reference skips such filters
    (FilterActive false, alc/alu.cpp:1623-1625)."""
    prod = mixlib.product()
    c = np.zeros(5, dtype=np.float32)
    assert prod.biquad_coeffs(0, 5000.0 / 48000.0, 1.0, 1.0, c.ctypes.data) == 0
    assert abs(c[0] - 1.0) < 1e-6 and abs(c[1] - c[3]) < 1e-6 and abs(c[2] - c[4]) < 1e-6
