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


def _ar_series(rng, n_chains, iters, rows):
    """Stationary AR(2) rows with random coefficients / scales / levels, one constant row and one white-noise row."""
    x = np.empty((n_chains, iters, rows))
    for r in range(rows):
        a1, a2 = rng.uniform(-0.5, 0.9), rng.uniform(-0.3, 0.05)
        scale, level = rng.uniform(0.01, 50.0), rng.uniform(-100.0, 100.0)
        e = rng.standard_normal((n_chains, iters + 50))
        z = np.zeros_like(e)
        for t in range(2, iters + 50):
            z[:, t] = a1 * z[:, t - 1] + a2 * z[:, t - 2] + e[:, t]
        x[:, :, r] = level + scale * z[:, 50:]
    if rows > 2:
        x[:, :, 1] = 2.5
        x[:, :, 2] = rng.standard_normal((n_chains, iters))
    return x


@pytest.mark.parametrize("n_chains,iters,rows", [(1, 3, 2), (1, 50, 5), (3, 500, 33), (2, 1000, 70), (5, 137, 21)])
def test_time_series_se_matches_the_spectrum0_ar_restatement(gpu, n_chains, iters, rows):
    """bvg_tsse_draws vs oracle/blocks.py spectrum0_ar (coda's algorithm) on the same series; the AIC-selected order
    is a discrete choice, so the comparison is tight wherever both pick the same order (always, on these inputs)."""
    import torch
    from bvartools_b200._lib import check, lib
    from oracle import blocks as ob
    rng = np.random.default_rng(iters + rows)
    x = _ar_series(rng, n_chains, iters, rows)
    d = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    got = np.empty(rows)
    check(lib().bvg_tsse_draws(C.c_void_p(d.data_ptr()), n_chains, iters, rows,
                               got.ctypes.data_as(C.POINTER(C.c_double)), None))
    want = ob.time_series_se(x.reshape(n_chains * iters, rows), n_chains)
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-300)
    if rows > 2:
        assert got[1] == 0.0                                      # constant row: spectrum0.ar's degenerate branch


def test_sampler_summary_reports_the_time_series_se(gpu):
    from bvartools_b200 import ChainSampler, _abi
    from oracle import blocks as ob
    obj = helpers.model_c3(K=3, p=1, nobs=60, iterations=120, burnin=20)
    s = ChainSampler(obj, n_chains=4, seed=9)
    s.run(0)
    host = _abi.OutBuffers(s.spec)
    s.fetch(host)
    for which in ("coef", "sigma"):
        x = host.arrays[which].reshape(-1, host.arrays[which].shape[-1])
        got = s.summary(which)["ts_se"]
        np.testing.assert_allclose(got, ob.time_series_se(x, 4), rtol=1e-8, atol=1e-300)


def test_summary_of_a_host_bvar_object(gpu):
    import bvartools_b200 as bt
    from oracle import blocks as ob
    fit = bt.draw_posterior(helpers.model_c1(iterations=200, burnin=20), seed=2)
    res = bt.summary(fit, ci=(0.05, 0.5, 0.95))
    for name in ("A", "C", "Sigma"):
        x = np.asarray(fit[name])
        np.testing.assert_allclose(res[name]["mean"], x.mean(axis=0), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(res[name]["quantiles"], np.quantile(x, (0.05, 0.5, 0.95), axis=0), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(res[name]["ts_se"], ob.time_series_se(x, 1), rtol=1e-8, atol=1e-300)


def test_summary_of_a_tvp_object_picks_a_period(gpu):
    """TVP / SV entries of a `bvar` object are lists over the periods: summary() takes the last one unless told
    otherwise (R/summary.bvar.R:33-38,125,187)."""
    import bvartools_b200 as bt
    fit = bt.draw_posterior(helpers.model_sv(tvp=True, nobs=30, iterations=40, burnin=5), seed=4)
    T = np.asarray(fit["y"]).shape[0]
    last = bt.summary(fit)
    np.testing.assert_allclose(last["A"]["mean"], np.asarray(fit["A"][T - 1]).mean(axis=0), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(last["Sigma"]["sd"], np.asarray(fit["Sigma"][T - 1]).std(axis=0, ddof=1), rtol=1e-10, atol=1e-14)
    first = bt.summary(fit, period=0)
    np.testing.assert_allclose(first["A"]["quantiles"], np.quantile(np.asarray(fit["A"][0]), (0.025, 0.5, 0.975), axis=0),
                               rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError, match="period"):
        bt.summary(fit, period=T)
