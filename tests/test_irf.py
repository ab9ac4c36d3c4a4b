"""Impulse responses per draw: the device kernel + radix-select bands against the oracle restatement
of .ir (src/ir.cpp:4-72) and irf.bvar (R/irf.bvar.R:140-216).  CPU: the oracle against the textbook recursion."""
import numpy as np
import pytest

import helpers
from oracle import blocks as ob

TOL = 1e-10


def test_oracle_ir_matches_companion_powers():
    """FEIR Phi_i = J C^i J' with the companion matrix C (Luetkepohl 2006, 2.1.22)."""
    rng = np.random.default_rng(1)
    k, p, h = 3, 2, 7
    A = rng.standard_normal((k, k * p)) * 0.3
    C = np.zeros((k * p, k * p))
    C[:k] = A
    C[k:, :-k] = np.eye(k * (p - 1))
    S = np.eye(k)
    for imp in range(1, k + 1):
        for resp in range(1, k + 1):
            th = ob.ir(A, S, h, "feir", imp, resp)
            want = [np.linalg.matrix_power(C, i)[:k, :k][resp - 1, imp - 1] for i in range(h + 1)]
            np.testing.assert_allclose(th, want, rtol=1e-12, atol=1e-14)
    # h < p: only the first h lags enter (ir.cpp:17-21)
    th = ob.ir(A, S, 1, "feir", 1, 2)
    np.testing.assert_allclose(th, [0.0, A[1, 0]], atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("type_,shock,cumulative", [("feir", 1, False), ("feir", 2.5, True), ("oir", 1, False),
                                                    ("oir", "sd", False), ("gir", "nsd", True), ("gir", 0.5, False)])
def test_irf_on_resident_draws_matches_oracle(gpu, type_, shock, cumulative):
    from bvartools_b200 import ChainSampler, _abi
    obj = helpers.model_c1(iterations=40, burnin=10)
    s = ChainSampler(obj, n_chains=4, seed=6)
    s.run(0)
    host = _abi.OutBuffers(s.spec)
    s.fetch(host)
    k, p = 3, 2
    coef = host.arrays["coef"].reshape(-1, host.arrays["coef"].shape[-1])
    sig = host.arrays["sigma"].reshape(-1, k * k)
    A_draws = coef[:, :k * k * p]
    for imp, resp, h in ((0, 1, 6), (2, 2, 1), (1, 0, 12)):
        kw = dict(n_ahead=h, ci=0.9, shock=shock, type=type_, cumulative=cumulative)
        want_d = ob.irf_bvar(A_draws, sig, k, imp + 1, resp + 1, keep_draws=True, **kw)
        got_d = s.irf(imp, resp, keep_draws=True, **kw)
        assert helpers.relerr(got_d, want_d) < TOL, (type_, imp, resp)
        want_b = ob.irf_bvar(A_draws, sig, k, imp + 1, resp + 1, **kw)
        got_b = s.irf(imp, resp, **kw)
        np.testing.assert_allclose(got_b, want_b, rtol=1e-10, atol=1e-14)
    with pytest.raises(ValueError):
        s.irf(0, 1, type="sir")
    s.close()


@pytest.mark.gpu
def test_irf_of_tvp_sv_draws_at_a_period(gpu):
    from bvartools_b200 import ChainSampler, _abi
    obj = helpers.model_sv(tvp=True, nobs=30, iterations=8, burnin=2)
    s = ChainSampler(obj, n_chains=3, seed=2)
    s.run(0)
    host = _abi.OutBuffers(s.spec)
    s.fetch(host)
    Y = obj["data"]["Y"]
    T, k = Y.shape
    p = s.lags
    n = host.arrays["coef"].shape[-1] // T
    coef = host.arrays["coef"].reshape(-1, T, n)
    sig = host.arrays["sigma"].reshape(-1, T, k * k)
    for period in (None, 0, 11):
        t = T - 1 if period is None else period
        want = ob.irf_bvar(coef[:, t, :k * k * p], sig[:, t], k, 1, 2, n_ahead=5, type="oir", shock="sd", keep_draws=True)
        got = s.irf(0, 1, n_ahead=5, type="oir", shock="sd", keep_draws=True, period=period)
        assert helpers.relerr(got, want) < TOL
    s.close()


def test_oracle_vardecomp_oir_rows_sum_to_one():
    """Orthogonalised FEVD shares of a response add up to one at every horizon (Luetkepohl 2006, 2.3.37)."""
    rng = np.random.default_rng(2)
    k, p, h = 3, 2, 6
    A = rng.standard_normal((k, k * p)) * 0.3
    B = rng.standard_normal((k, k))
    S = B @ B.T + np.eye(k)
    for resp in range(1, k + 1):
        np.testing.assert_allclose(ob.vardecomp(A, S, h, "oir", resp).sum(axis=1), 1.0, rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("type_,norm", [("oir", False), ("gir", False), ("gir", True)])
def test_fevd_on_resident_draws_matches_oracle(gpu, type_, norm):
    from bvartools_b200 import ChainSampler, _abi
    obj = helpers.model_c1(iterations=50, burnin=10)
    s = ChainSampler(obj, n_chains=5, seed=8)
    s.run(0)
    host = _abi.OutBuffers(s.spec)
    s.fetch(host)
    k, p = 3, 2
    coef = host.arrays["coef"].reshape(-1, host.arrays["coef"].shape[-1])
    sig = host.arrays["sigma"].reshape(-1, k * k)
    for resp, h in ((0, 5), (2, 1), (1, 10)):
        want = ob.fevd_bvar(coef[:, :k * k * p], sig, k, resp + 1, n_ahead=h, type=type_, normalise_gir=norm)
        got = s.fevd(resp, n_ahead=h, type=type_, normalise_gir=norm)
        assert helpers.relerr(got, want) < TOL, (type_, resp, h)
    with pytest.raises(ValueError):
        s.fevd(0, type="sgir")
    s.close()


@pytest.mark.gpu
def test_irf_and_fevd_of_a_host_bvar_object(gpu):
    """Package-level irf() / fevd() on a `bvar` object (draws on the host, as the reference's methods receive them)."""
    import bvartools_b200 as bt
    obj = helpers.model_c1(iterations=60, burnin=10)
    fit = bt.draw_posterior(obj, seed=4)
    k = 3
    A, S = np.asarray(fit["A"]), np.asarray(fit["Sigma"])
    want = ob.irf_bvar(A, S, k, 1, 2, n_ahead=8, type="oir", shock="sd", ci=0.9)
    np.testing.assert_allclose(bt.irf(fit, 0, 1, n_ahead=8, type="oir", shock="sd", ci=0.9), want, rtol=1e-10, atol=1e-14)
    want_v = ob.fevd_bvar(A, S, k, 3, n_ahead=6, type="gir", normalise_gir=True)
    assert helpers.relerr(bt.fevd(fit, 2, n_ahead=6, type="gir", normalise_gir=True), want_v) < TOL



@pytest.mark.gpu
@pytest.mark.parametrize("type", ["sir", "sgir"])
def test_structural_impulse_responses(gpu, type):
    """sir / sgir (src/ir.cpp:35-37,42-46) from the A0 draws of a structural model: resident draws (compact a0 rows) and
    the host `bvar` object (full A0 matrices) against the literal restatement."""
    import bvartools_b200 as bt
    from bvartools_b200 import ChainSampler, _abi
    obj = helpers.model_structural(K=3, p=2, nobs=60, iterations=40, burnin=10)
    k = 3
    s = ChainSampler(obj, n_chains=2, seed=5)
    s.run(0)
    host = _abi.OutBuffers(s.spec)
    s.fetch(host)
    coef = host.arrays["coef"].reshape(-1, host.arrays["coef"].shape[-1])
    sig = host.arrays["sigma"].reshape(coef.shape[0], -1)
    n_a = k * k * 2
    a0c = coef[:, -3:]
    A0 = np.tile(np.eye(k).reshape(-1, order="F"), (coef.shape[0], 1))
    lower = [i + k * j for j in range(k) for i in range(k) if i > j]
    A0[:, lower] = a0c
    want = ob.irf_bvar(coef[:, :n_a], sig, k, 1, 3, n_ahead=6, type=type, shock="sd", keep_draws=True, A0_draws=A0)
    got = s.irf(0, 2, n_ahead=6, type=type, shock="sd", keep_draws=True)
    assert helpers.relerr(got, want) < TOL
    s.close()
    fit = bt.draw_posterior(obj, seed=5)
    wantf = ob.irf_bvar(np.asarray(fit["A"]), np.asarray(fit["Sigma"]), k, 2, 3, n_ahead=5, type=type, shock=1, ci=0.9,
                        A0_draws=np.asarray(fit["A0"]))
    np.testing.assert_allclose(bt.irf(fit, 1, 2, n_ahead=5, type=type, ci=0.9), wantf, rtol=1e-10, atol=1e-14)
    with pytest.raises(ValueError, match="A0"):
        bt.irf({k_: v for k_, v in fit.items() if k_ != "A0"}, 1, 2, type=type)


@pytest.mark.gpu
@pytest.mark.parametrize("type", ["sir", "sgir"])
def test_structural_variance_decomposition(gpu, type):
    """sir / sgir of .vardecomp (src/vardecomp.cpp:32-36,41-45): resident draws (compact a0 rows) and the host object."""
    import bvartools_b200 as bt
    from bvartools_b200 import ChainSampler, _abi
    obj = helpers.model_structural(K=3, p=2, nobs=60, iterations=30, burnin=10)
    k = 3
    s = ChainSampler(obj, n_chains=2, seed=8)
    s.run(0)
    host = _abi.OutBuffers(s.spec)
    s.fetch(host)
    coef = host.arrays["coef"].reshape(-1, host.arrays["coef"].shape[-1])
    sig = host.arrays["sigma"].reshape(coef.shape[0], -1)
    A0 = np.tile(np.eye(k).reshape(-1, order="F"), (coef.shape[0], 1))
    A0[:, [i + k * j for j in range(k) for i in range(k) if i > j]] = coef[:, -3:]
    want = ob.fevd_bvar(coef[:, :k * k * 2], sig, k, 2, n_ahead=6, type=type, normalise_gir=True, A0_draws=A0)
    assert helpers.relerr(s.fevd(1, n_ahead=6, type=type, normalise_gir=True), want) < TOL
    s.close()
    fit = bt.draw_posterior(obj, seed=8)
    wantf = ob.fevd_bvar(np.asarray(fit["A"]), np.asarray(fit["Sigma"]), k, 3, n_ahead=4, type=type,
                         A0_draws=np.asarray(fit["A0"]))
    assert helpers.relerr(bt.fevd(fit, 2, n_ahead=4, type=type), wantf) < TOL
