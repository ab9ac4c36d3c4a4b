"""Posterior summaries of device-resident draws: mean, SD and exact type-7 quantiles (radix select)
against NumPy on the fetched draws (np.quantile's default "linear" method is R's type 7, the one coda::summary.mcmc
uses in R/summary.bvar.R:121-140)."""
import ctypes as C

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _device_summary(x, probs):
    """x: [n_draws][rows] host array -> (mean, sd, quantiles) through bvg_summary_draws on a device copy."""
    import torch
    from bvartools_b200._lib import check, lib
    d = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    n, rows = x.shape
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    mean, sd, q = np.empty(rows), np.empty(rows), np.empty((probs.size, rows))
    dp = C.POINTER(C.c_double)
    check(lib().bvg_summary_draws(C.c_void_p(d.data_ptr()), n, rows, probs.ctypes.data_as(dp), probs.size,
                                  mean.ctypes.data_as(dp), sd.ctypes.data_as(dp), q.ctypes.data_as(dp), None))
    return mean, sd, q


@pytest.mark.parametrize("n,rows", [(1, 3), (2, 1), (7, 5), (1000, 21), (4097, 150), (300, 700)])
def test_summary_matches_numpy(gpu, n, rows):
    rng = np.random.default_rng(n + rows)
    x = rng.standard_normal((n, rows)) * rng.uniform(0.01, 100.0, rows) + rng.uniform(-50.0, 50.0, rows)
    probs = [0.0, 0.025, 0.25, 0.5, 0.75, 0.975, 1.0]
    mean, sd, q = _device_summary(x, probs)
    np.testing.assert_allclose(mean, x.mean(axis=0), rtol=1e-12, atol=1e-12)
    if n > 1:
        np.testing.assert_allclose(sd, x.std(axis=0, ddof=1), rtol=1e-11)
    np.testing.assert_allclose(q, np.quantile(x, probs, axis=0), rtol=1e-13, atol=1e-13)


def test_quantiles_with_ties_signs_and_zeros(gpu):
    rng = np.random.default_rng(5)
    x = np.round(rng.standard_normal((501, 4)) * 3.0)            # heavy ties, both signs, +-0
    x[::7, 1] = -0.0
    x[:, 2] = 1.25                                                # a constant row
    x[:, 3] = np.abs(x[:, 3]) * 1e-300                            # subnormal-adjacent magnitudes
    probs = [0.01, 0.1, 0.5, 0.9, 0.99]
    _, _, q = _device_summary(x, probs)
    np.testing.assert_allclose(q, np.quantile(x, probs, axis=0), rtol=1e-13, atol=0)


def test_sampler_summary_on_resident_draws(gpu):
    from bvartools_b200 import ChainSampler, _abi
    obj = helpers.model_c3(K=3, p=2, nobs=80, iterations=60, burnin=20)           # SSVS: coef, sigma, lambda
    s = ChainSampler(obj, n_chains=9, seed=4)
    s.run(0)
    host = _abi.OutBuffers(s.spec)
    s.fetch(host)
    for which in ("coef", "sigma", "lambda"):
        got = s.summary(which, ci=0.9)
        x = host.arrays[which].reshape(-1, host.arrays[which].shape[-1])
        assert got["n"] == x.shape[0]
        np.testing.assert_allclose(got["mean"], x.mean(axis=0), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(got["sd"], x.std(axis=0, ddof=1), rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(got["quantiles"], np.quantile(x, [0.05, 0.5, 0.95], axis=0), rtol=1e-13, atol=1e-15)
    s.close()
