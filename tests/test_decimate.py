"""EXTENSION (parity unpinned by the reference): the polyphase decimation kernel against scipy in float64."""
import numpy as np
import pytest

from audio_suite_b200 import decimate as D


def _check(dev, precision, tol):
    from scipy import signal
    rng = np.random.default_rng(12)
    for n, q, taps in ((1000, 4, 33), (4097, 16, 321), (777, 3, 8), (50, 7, 101), (30000, 16, 321)):
        x = rng.standard_normal((3, n))
        h = rng.standard_normal(taps)
        want = np.stack([signal.upfirdn(h, xi, up=1, down=q) for xi in x])
        got = D.upfirdn_decimate(x, h, q, dev, precision)
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < tol, (n, q, taps)
    for n, q in ((48000, 16), (7680, 16), (1001, 5), (96, 2)):
        x = rng.standard_normal(n)
        want = signal.resample_poly(x, 1, q)
        got = D.decimate(x, q, dev, precision)
        assert got.shape == want.shape and np.max(np.abs(got - want)) < tol * 4, (n, q)


@pytest.mark.parametrize("precision,tol", [("f64", 1e-12), ("f32", 2e-5)])
def test_polyphase_decimation_against_scipy_emulated(emul, precision, tol):
    _check(emul, precision, tol)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("f64", 1e-12), ("f32", 2e-5)])
def test_polyphase_decimation_against_scipy(cuda_dev, precision, tol):
    _check(cuda_dev, precision, tol)
