"""Model comparison: loglik_normal and the summary.bvarlist information criteria.

CPU: the oracle restatement (oracle/blocks.py: loglik_normal, summary_bvarlist_row -- src/loglik_normal.cpp:37-58,
R/summary.bvarlist.R:143-185) is pinned analytically against SciPy's multivariate normal log-density (the reference has
no fixture for it: parity unpinned otherwise).  GPU: the fused device reduction against the oracle's per-draw loop,
<= 1e-10 relative, for constant / SV / TVP draws, host-buffer and resident-draw entry points."""
import numpy as np
import pytest
from scipy.stats import multivariate_normal

import helpers
from oracle import blocks as ob

TOL = 1e-10


def _spd(rng, k):
    a = rng.standard_normal((k, k + 2))
    return a @ a.T / (k + 2) + 0.3 * np.eye(k)


@pytest.mark.parametrize("k,T", [(1, 5), (3, 73), (6, 40)])
def test_oracle_loglik_normal_is_the_gaussian_log_density(k, T):
    rng = np.random.default_rng(k * 100 + T)
    u = rng.standard_normal((k, T))
    sigma = _spd(rng, k)
    want = multivariate_normal(mean=np.zeros(k), cov=sigma).logpdf(u.T)
    np.testing.assert_allclose(ob.loglik_normal(u, sigma), np.atleast_1d(want), rtol=1e-12, atol=1e-12)
    stack = np.vstack([_spd(rng, k) for _ in range(T)])
    got = ob.loglik_normal(u, stack)
    for t in range(T):
        w = multivariate_normal(mean=np.zeros(k), cov=stack[t * k:(t + 1) * k]).logpdf(u[:, t])
        assert abs(got[t] - w) <= 1e-12 * max(1.0, abs(w))


def test_oracle_summary_row_formulae():
    rng = np.random.default_rng(3)
    T, k, M, draws = 30, 2, 3, 7
    y, x = rng.standard_normal((T, k)), rng.standard_normal((T, M))
    pars = rng.standard_normal((draws, k * M)) * 0.1
    sig = np.stack([_spd(rng, k).reshape(-1, order="F") for _ in range(draws)])
    row = ob.summary_bvarlist_row(y, x, pars, sig)
    ll = row["LL"]
    assert row["AIC"] == pytest.approx(2 * M - 2 * ll)
    assert row["BIC"] == pytest.approx(np.log(T) * M - 2 * ll)
    assert row["HQ"] == pytest.approx(2 * np.log(np.log(T)) * M - 2 * ll)
    # a single draw: LL is the Gaussian log-likelihood of the residuals
    one = ob.summary_bvarlist_row(y, x, pars[:1], sig[:1])
    u = y.T - pars[0].reshape(k, M, order="F") @ x.T
    want = multivariate_normal(mean=np.zeros(k), cov=sig[0].reshape(k, k, order="F")).logpdf(u.T).sum()
    assert one["LL"] == pytest.approx(want, rel=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("k,T,tv", [(1, 9, False), (3, 73, False), (3, 73, True), (6, 300, False), (20, 64, True)])
def test_loglik_normal_block_matches_oracle(gpu, k, T, tv):
    from bvartools_b200 import loglik_normal
    rng = np.random.default_rng(7 * k + T)
    u = rng.standard_normal((k, T))
    sigma = np.vstack([_spd(rng, k) for _ in range(T)]) if tv else _spd(rng, k)
    got, want = loglik_normal(u, sigma), ob.loglik_normal(u, sigma)
    assert helpers.relerr(got, want) < TOL


@pytest.mark.gpu
def test_loglik_normal_error_behaviour(gpu):
    from bvartools_b200 import loglik_normal
    from bvartools_b200._lib import BvgError
    u = np.ones((2, 4))
    u[1, 2] = np.nan
    with pytest.raises(BvgError):
        loglik_normal(u, np.eye(2))
    with pytest.raises(ValueError):
        loglik_normal(np.ones((2, 4)), np.eye(3))


def _bvar_rows(obj, res):
    """temp_pars / Sigma draws of one chain in the layout summary.bvarlist sees (draws x rows; lists over periods)."""
    post = res["posteriors"]
    T, k = obj["data"]["Y"].shape
    parts = [np.asarray(post[name]["coeffs"]) for name in ("a", "b", "c") if post.get(name) is not None]
    tvp = bool(obj["model"].get("tvp"))
    sv = bool(obj["model"].get("sv"))
    if tvp:      # each block is time-major on its own (rows_block x T): cbind per period, R/summary.bvarlist.R:68-71
        pars = [np.concatenate([p[t * (p.shape[0] // T):(t + 1) * (p.shape[0] // T)] for p in parts], axis=0).T
                for t in range(T)]
    else:
        pars = np.concatenate(parts, axis=0).T                            # draws x n
    sg = np.asarray(post["sigma"]["coeffs"])
    sig = [sg[t * k * k:(t + 1) * k * k].T for t in range(T)] if sv else sg.T
    return pars, sig, tvp, sv


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c1", "sv", "tvp_sv"])
def test_summary_bvarlist_matches_oracle_loop(gpu, case):
    """Draws of a short GPU run -> bvar objects -> summary_bvarlist (fused device reduction) vs the reference's per-draw
    loop restated in the oracle."""
    import bvartools_b200 as bt
    if case == "c1":
        objs = [helpers.model_c1(iterations=40, burnin=10), helpers.model_c1(iterations=25, burnin=5)]
    elif case == "sv":
        objs = [helpers.model_sv(tvp=False, nobs=40, iterations=12, burnin=4)]
    else:
        objs = [helpers.model_sv(tvp=True, nobs=30, iterations=10, burnin=4)]
    fitted = bt.draw_posterior(objs, seed=11)
    rows = bt.summary_bvarlist(fitted)
    for idx, (obj, fit, row) in enumerate(zip(objs, fitted, rows)):
        assert not fit.get("error"), fit
        res = (bt.bvartvpalg if obj["model"].get("tvp") else bt.bvaralg)(obj, seed=11 + idx)
        pars, sig, tvp, sv = _bvar_rows(obj, res)
        want = ob.summary_bvarlist_row(obj["data"]["Y"], obj["data"]["Z"], pars, sig, tvp=tvp, sv=sv)
        for key in ("LL", "AIC", "BIC", "HQ"):
            assert abs(row[key] - want[key]) <= TOL * abs(want[key]), (case, key, row[key], want[key])
        if row["ll_t"] is not None:
            assert helpers.relerr(row["ll_t"], want["ll_t"]) < TOL


@pytest.mark.gpu
def test_resident_draws_and_host_draws_agree_and_shard(gpu):
    """ChainSampler.loglik() (draws stay in HBM) == the host-buffer path on the fetched draws; per-period sums of two
    shards add up to those of the full job (what a multi-GPU run all-reduces)."""
    import bvartools_b200 as bt
    from bvartools_b200 import ChainSampler, _abi
    obj = helpers.model_c1(iterations=30, burnin=5)
    full = ChainSampler(obj, n_chains=12, seed=5)
    full.run(0)
    ll_full, n_full = full.loglik()
    host = _abi.OutBuffers(full.spec)
    full.fetch(host)
    T, k = obj["data"]["Y"].shape
    coef = host.arrays["coef"].reshape(-1, host.arrays["coef"].shape[-1])
    sig = host.arrays["sigma"].reshape(-1, k * k)
    want = np.zeros(T)
    for j in range(coef.shape[0]):
        u = obj["data"]["Y"].T - coef[j].reshape(k, -1, order="F") @ obj["data"]["Z"].T
        want += ob.loglik_normal(u, sig[j].reshape(k, k, order="F"))
    assert n_full == 12 * 30
    assert helpers.relerr(ll_full, want) < TOL
    tot, n_tot = full.loglik_total()                       # cross-moment kernel: no pass over the periods
    assert n_tot == n_full and abs(tot - want.sum()) <= TOL * abs(want.sum())
    parts = np.zeros(T)
    for off, cnt in ((0, 5), (5, 7)):
        s = ChainSampler(obj, n_chains=cnt, seed=5, chain_offset=off)
        s.run(0)
        ll, n = s.loglik()
        parts += ll
        s.close()
    np.testing.assert_allclose(parts, ll_full, rtol=1e-12)
    full.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k,p", [(1, 1), (3, 4), (6, 2)])
def test_total_kernel_matches_per_period_kernel(gpu, k, p):
    """The HBM-bound total (cross-moment form) against the per-period kernel and the oracle on random draws."""
    import ctypes as C
    from bvartools_b200._lib import check, lib
    rng = np.random.default_rng(10 * k + p)
    y = helpers.synth_var(k, p, 90, 4)
    m = helpers.gen_var(y, p=p, deterministic="const", iterations=5, burnin=1)
    Y, Z = np.ascontiguousarray(m["data"]["Y"]), np.ascontiguousarray(m["data"]["Z"])
    T, M = Z.shape
    draws = 37
    ols = np.linalg.lstsq(Z, Y, rcond=None)[0].T                                   # k x M
    coef = np.stack([(ols + 0.05 * rng.standard_normal(ols.shape)).reshape(-1, order="F") for _ in range(draws)])
    sig = np.stack([_spd(rng, k).reshape(-1, order="F") for _ in range(draws)])
    dp = C.POINTER(C.c_double)
    tot = C.c_double(0.0)
    check(lib().bvg_loglik_total_host(coef.ctypes.data_as(dp), sig.ctypes.data_as(dp), draws, k, M, T,
                                      Y.ctypes.data_as(dp), Z.ctypes.data_as(dp), C.byref(tot)))
    ll_t = np.empty(T)
    check(lib().bvg_loglik_draws_host(coef.ctypes.data_as(dp), sig.ctypes.data_as(dp), draws, k, M, T, 0, 0,
                                      Y.ctypes.data_as(dp), Z.ctypes.data_as(dp), ll_t.ctypes.data_as(dp)))
    want = ob.summary_bvarlist_row(Y, Z, coef, sig)
    assert abs(ll_t.sum() / draws - want["LL"]) <= TOL * abs(want["LL"])
    assert abs(tot.value / draws - want["LL"]) <= TOL * abs(want["LL"])


@pytest.mark.gpu
@pytest.mark.parametrize("tvp", [False, True])
def test_resident_draws_time_varying_models(gpu, tvp):
    """ChainSampler.loglik / loglik_total on SV and TVP-SV draws resident in HBM against the oracle loop."""
    from bvartools_b200 import ChainSampler, _abi
    obj = helpers.model_sv(tvp=tvp, nobs=30, iterations=6, burnin=2)
    s = ChainSampler(obj, n_chains=3, seed=9)
    s.run(0)
    ll_t, n = s.loglik()
    tot, n2 = s.loglik_total()
    host = _abi.OutBuffers(s.spec)
    s.fetch(host)
    s.close()
    Y, Z = obj["data"]["Y"], obj["data"]["Z"]
    T, k = Y.shape
    want = np.zeros(T)
    coef, sig = host.arrays["coef"], host.arrays["sigma"]
    for c in range(coef.shape[0]):
        for j in range(coef.shape[1]):
            u = np.empty((k, T))
            for t in range(T):
                a = coef[c, j].reshape(T, -1)[t] if tvp else coef[c, j]
                u[:, t] = Y[t] - a.reshape(k, -1, order="F") @ Z[t]
            stack = np.vstack([sig[c, j].reshape(T, k * k)[t].reshape(k, k, order="F") for t in range(T)])
            want += ob.loglik_normal(u, stack)
    assert n == n2 == 3 * 6
    assert helpers.relerr(ll_t, want) < TOL
    assert abs(tot - want.sum()) <= TOL * abs(want.sum())

