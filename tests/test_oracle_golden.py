"""The CPU oracle (oracle/almix_oracle.c) against the committed golden vectors that
were rendered by the compiled, unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from helpers import golden, mixlib


@pytest.mark.parametrize("name", golden.names())
def test_oracle_matches_reference_c_kernels_bit_exact(name):
    fx = golden.load(name)
    out, _ = golden.replay(mixlib.oracle(), fx)
    ref = fx["out_c"]
    assert out.shape == ref.shape
    if "conv_taps" in fx or "conv_taps_chain" in fx or "uhj_fir" in fx:
        # the convolution slot (and the FIR UHJ encoder's phase shifter) is restated by its
        # definition (direct linear convolution, f64 accumulation), not by pffft's float
        # butterflies: equal within rounding, not bitwise
        assert np.abs(out.astype(np.float64) - ref).max() <= 5e-7
        return
    if "pshifter" in name:
        # the pitch shifter's real FFT (pffft, single precision) is evaluated as a complex FFT in
        # double: rounding differences only.  The reference's own SSE and C builds differ by
        # 1.5e-6 / 1.9e-6 on these two scenes.
        assert np.abs(out.astype(np.float64) - ref).max() <= 4e-6
        assert np.sqrt(((out.astype(np.float64) - ref) ** 2).mean()) <= 6e-7
        return
    # the restatement follows the reference's C kernels operation for operation
    assert np.array_equal(out, ref), f"max diff {np.abs(out - ref).max():.3e}"


@pytest.mark.parametrize("name", golden.names())
def test_oracle_matches_reference_sse_kernels(name):
    fx = golden.load(name)
    out, _ = golden.replay(mixlib.oracle(), fx)
    ref = fx["out_sse"].astype(np.float64)
    err = out - ref
    if "out_type" in fx:
        # integer output: the SSE kernels' last-bit differences can move a sample across a
        # rounding boundary — by one step at most, and rarely
        assert np.abs(err).max() <= 1 and (err != 0).mean() <= 0.01
        assert np.ptp(ref) > 8
        return
    gap_rms, gap_max = golden.kernel_set_gap(fx)
    assert np.sqrt((err ** 2).mean()) <= max(1e-7, 1.5 * gap_rms)
    assert np.abs(err).max() <= max(1e-6, 1.5 * gap_max)
    assert np.abs(ref).max() > 1e-3  # the scene is not silent
