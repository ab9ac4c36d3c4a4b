"""CPU checks of the oracle's spectrum0.ar restatement (coda's time-series SE; coda is not vendored by the reference, so
these are analytic pins: the zero-frequency spectrum of a known AR process, white noise, and the degenerate branch)."""
import numpy as np

from oracle import blocks as ob


def test_white_noise_spectrum_is_the_variance():
    rng = np.random.default_rng(0)
    x = 3.0 * rng.standard_normal(20000) + 7.0
    spec, order = ob.spectrum0_ar(x)
    assert 0 <= order <= 43
    assert abs(spec / 9.0 - 1.0) < 0.15


def test_ar1_spectrum_at_zero():
    rng = np.random.default_rng(1)
    n, phi = 40000, 0.7
    e = rng.standard_normal(n)
    x = np.zeros(n)
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    spec, order = ob.spectrum0_ar(x)
    assert order >= 1
    assert abs(spec * (1.0 - phi) ** 2 - 1.0) < 0.1


def test_degenerate_series_and_pooling():
    assert ob.spectrum0_ar(np.full(50, -4.0)) == (0.0, 0)
    assert ob.spectrum0_ar(np.arange(30.0))[0] == 0.0             # pure trend: residuals of lm(x ~ t) vanish
    rng = np.random.default_rng(2)
    x = rng.standard_normal((600, 2))
    one = ob.time_series_se(x, 1)
    three = ob.time_series_se(x, 3)
    naive = x.std(axis=0, ddof=1) / np.sqrt(600)
    assert np.all(np.abs(one / naive - 1.0) < 0.25) and np.all(np.abs(three / naive - 1.0) < 0.35)
