"""CPU tests of the oracle itself: pinned against the only numeric known-answers the reference tree holds
(README summary table, analytic flat-prior posterior), literal-vs-structured self-consistency, and agreement of
the two independent restatements (NumPy and plain C)."""
import json
import pathlib

import numpy as np
import pytest

import helpers
from oracle import blocks, samplers
from oracle.dense_ref import bvar_dense
from oracle.variates import VariateSource

GOLD = pathlib.Path(__file__).resolve().parent / "golden"


def _batch_mcse(x, nb=20):
    n = (x.shape[0] // nb) * nb
    bm = x[:n].reshape(nb, -1, *x.shape[1:]).mean(axis=1)
    return bm.std(axis=0, ddof=1) / np.sqrt(nb)


def test_e1_fixture_matches_readme_sample():
    y, names = helpers.e1_readme_sample()
    assert y.shape == (75, 3) and names == ["invest", "income", "cons"]
    m = helpers.model_c1()
    assert m["data"]["Y"].shape == (73, 3) and m["data"]["Z"].shape == (73, 7) and m["data"]["SUR"].shape == (219, 21)
    # OLS on the fixture reproduces the numbers quoted in BASELINE.md / SURVEY.md section 4
    ols = m["initial"]["coefficients"]["draw"].reshape(3, 7, order="F")
    np.testing.assert_allclose(ols[0], [-0.3196, 0.1460, 0.9612, -0.1606, 0.1146, 0.9344, -1.6722], atol=5e-4)


@pytest.fixture(scope="module")
def c1_long_run():
    obj = helpers.model_c1(iterations=10000, burnin=1000)
    res = samplers.bvaralg(obj, VariateSource(rng=np.random.default_rng(20240113)), literal=False)
    return obj, res["posteriors"]


def test_oracle_reproduces_readme_table(c1_long_run):
    """Statistical pin: README.md:240-299 (R-level post_normal + rWishart loop, 10000 draws) vs the oracle's
    bvaralg restatement.  Means within 4 MCSE (both runs are Monte Carlo -> sqrt(2) * MCSE), SDs within 6 %."""
    obj, post = c1_long_run
    table = json.loads((GOLD / "readme_table_c1.json").read_text())
    coef = np.vstack([post["a"]["coeffs"], post["c"]["coeffs"]]).T        # iter x 21 (coefficient id j*k+i)
    names = ["invest.01", "income.01", "cons.01", "invest.02", "income.02", "cons.02", "const"]
    mcse = _batch_mcse(coef)
    for i, var in enumerate(["invest", "income", "cons"]):
        for j, reg in enumerate(names):
            row = table["coef"][var][reg]
            col = j * 3 + i
            tol = 4.0 * np.sqrt(2.0) * max(mcse[col], row["TSSD"])
            assert abs(coef[:, col].mean() - row["Mean"]) < tol, (var, reg, coef[:, col].mean(), row["Mean"], tol)
            assert abs(coef[:, col].std(ddof=1) / row["SD"] - 1.0) < 0.06
            assert abs(np.quantile(coef[:, col], 0.5) - row["q50"]) < 6.0 * row["TSSD"] * 1.3
    sig = post["sigma"]["coeffs"].T
    smc = _batch_mcse(sig)
    vn = ["invest", "income", "cons"]
    for a in range(3):
        for b in range(3):
            row = table["sigma"][f"{vn[a]}_{vn[b]}"]
            col = a + 3 * b
            tol = 4.0 * np.sqrt(2.0) * max(smc[col], row["TSSD"])
            assert abs(sig[:, col].mean() - row["Mean"]) < tol, (a, b, sig[:, col].mean(), row["Mean"])


def test_oracle_analytic_flat_prior(c1_long_run):
    """Flat coefficient prior: E[a] = OLS, E[Sigma] = (S0 + SSE) / (T - M + df - k - 1) = SSE / 63."""
    obj, post = c1_long_run
    coef = np.vstack([post["a"]["coeffs"], post["c"]["coeffs"]]).T
    ols = obj["initial"]["coefficients"]["draw"].reshape(-1)
    assert np.all(np.abs(coef.mean(axis=0) - ols) < 4.0 * _batch_mcse(coef) + 1e-12)
    Y, Z = obj["data"]["Y"], obj["data"]["Z"]
    U = Y - Z @ np.linalg.lstsq(Z, Y, rcond=None)[0]
    expect = (U.T @ U + 1e-4 * np.eye(3)) / 63.0
    sig = post["sigma"]["coeffs"].T
    assert np.all(np.abs(sig.mean(axis=0) - expect.reshape(-1, order="F")) < 4.0 * _batch_mcse(sig))


@pytest.mark.parametrize("builder", [
    lambda: helpers.model_c1(iterations=12, burnin=4),
    lambda: helpers.model_c3(K=3, p=2, nobs=60, iterations=10, burnin=3),
    lambda: helpers.model_c3(K=3, p=2, nobs=60, iterations=10, burnin=3, bvs=True),
    lambda: helpers.model_c1(iterations=10, burnin=3, sigma=dict(shape=3, rate=1e-4), coef=dict(v_i=1, v_i_det=0.1)),
    lambda: helpers.model_sv(nobs=40, iterations=8, burnin=3),
    lambda: helpers.model_sv(nobs=24, iterations=6, burnin=2, tvp=True),
])
def test_literal_equals_structured(builder):
    """SURVEY.md appendix D 'oracle self-check': dense-kron literal path vs Kronecker/banded path, same variates."""
    obj = builder()
    per, _ = helpers.make_variates(obj, 1, seed=3)
    lit = helpers.oracle_stack(obj, per, literal=True)
    stru = helpers.oracle_stack(obj, per, literal=False)
    for name in lit:
        assert helpers.relerr(stru[name], lit[name]) < 1e-9, name
    if "lambda" in lit:
        assert np.array_equal(lit["lambda"], stru["lambda"])


def test_literal_equals_structured_c4_full_size():
    """Config c4 at BASELINE size (T = 200, state dimension nT = 4200): the dense (nT)^2 LU + Cholesky path of
    bvartvpalg.cpp:290-297 against the block-bidiagonal restatement the device tests use as their checker."""
    obj = helpers.model_sv(tvp=True, nobs=202, iterations=1, burnin=1)
    per, _ = helpers.make_variates(obj, 1, seed=3)
    lit = helpers.oracle_stack(obj, per, literal=True)
    stru = helpers.oracle_stack(obj, per, literal=False)
    assert lit["coef"].shape == (1, 1, 4200)
    for name in lit:
        assert helpers.relerr(stru[name], lit[name]) < 1e-9, name


@pytest.mark.parametrize("builder", [
    lambda: helpers.model_c1(iterations=25, burnin=5),
    lambda: helpers.model_c3(K=3, p=2, nobs=80, iterations=20, burnin=5),
])
def test_c_port_equals_numpy_oracle(builder):
    obj = builder()
    per, stacked = helpers.make_variates(obj, 2, seed=9)
    got = bvar_dense(obj, n_chains=2, inject=stacked)
    want = helpers.oracle_stack(obj, per)
    for name in got:
        assert helpers.relerr(got[name], want[name]) < 1e-10, name
    if "lambda" in got:
        assert np.array_equal(got["lambda"], want["lambda"])


def test_stochvol_tables_and_shapes():
    assert abs(blocks.KSC_P.sum() - 1.0) < 1e-12 and abs(blocks.OCSN_P.sum() - 1.0) < 1e-4
    rng = np.random.default_rng(1)
    T, k = 50, 2
    y = rng.standard_normal((T, k)) * 0.5
    h0 = np.log(np.tile(y.var(axis=0), (T, 1)))
    for table in ("ksc1998", "ocsn2007"):
        h, s = blocks.stochvol_mixture(y, h0, np.full(k, 0.05), h0[0], np.full(k, 1e-4), rng.random((k, T)),
                                       rng.standard_normal((k, T)), table=table, return_s=True)
        assert h.shape == (T, k) and s.min() >= 0 and s.max() < len(blocks.MIXTURES[table][0])
    with pytest.raises(ValueError):
        blocks.stochvol_mixture(y * np.nan, h0, np.full(k, 0.05), h0[0], np.full(k, 1e-4), rng.random((k, T)),
                                rng.standard_normal((k, T)))


def test_post_normal_matches_sur_form_and_cholesky_moments():
    """post_normal (Kronecker form) and post_normal_sur (SUR form) are the same map eps -> a."""
    obj = helpers.model_c1()
    y, x, z = obj["data"]["Y"].T, obj["data"]["Z"].T, obj["data"]["SUR"]
    rng = np.random.default_rng(2)
    S = np.cov(rng.standard_normal((3, 30)))
    si = np.linalg.inv(S)
    mu = rng.standard_normal(21)
    vi = np.diag(rng.uniform(0.1, 2.0, 21))
    eps = rng.standard_normal(21)
    a1 = blocks.post_normal(y, x, si, mu, vi, eps)
    a2 = blocks.post_normal_sur(y, z, si, mu, vi, eps)
    a3 = blocks.post_normal_sur(y, z, np.tile(si, (y.shape[1], 1)), mu, vi, eps, svd=True)
    np.testing.assert_allclose(a1, a2, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(a1, a3, rtol=1e-9, atol=1e-10)


def test_kalman_dk_constant_equals_stacked_inputs():
    rng = np.random.default_rng(4)
    k, T, n = 2, 12, 4
    y = rng.standard_normal((k, T))
    z = rng.standard_normal((k * T, n))
    su = np.cov(rng.standard_normal((k, 20)))
    svv = np.diag(rng.uniform(0.01, 0.1, n))
    B = np.eye(n)
    e0, ey, ea = rng.standard_normal(n), rng.standard_normal((T, k)), rng.standard_normal((T, n))
    a1 = blocks.kalman_dk(y, z, su, svv, B, np.zeros(n), np.eye(n), e0, ey, ea)
    a2 = blocks.kalman_dk(y, z, np.tile(su, (T, 1)), np.tile(svv, (T, 1)), np.tile(B, (T, 1)), np.zeros(n),
                          np.eye(n), e0, ey, ea)
    assert a1.shape == (n, T + 1)
    np.testing.assert_allclose(a1, a2, rtol=1e-12, atol=1e-12)


def test_kalman_dk_posterior_mean_matches_precision_form():
    """With zero simulation noise removed (eps = 0) DK returns the smoothed mean; for a random-walk state it must
    equal the mean of the Chan-Jeliazkov precision form used by bvartvpalg (bvartvpalg.cpp:293-296)."""
    rng = np.random.default_rng(5)
    k, T, n = 2, 10, 3
    y = rng.standard_normal((k, T))
    z = rng.standard_normal((k * T, n))
    su = np.diag(rng.uniform(0.5, 1.5, k))
    q = rng.uniform(5.0, 20.0, n)                       # state precisions
    a0 = rng.standard_normal(n)
    # bvartvpalg's a_0 = a_init + v_0, i.e. DK's initial state a_0 ~ N(a_init, Sigma_v)
    sm = blocks.kalman_dk(y, z, su, np.diag(1.0 / q), np.eye(n), a0, np.diag(1.0 / q),
                          np.zeros(n), np.zeros((T, k)), np.zeros((T, n)))
    Zt = [z[t * k:(t + 1) * k] for t in range(T)]
    sig_t = np.tile(np.linalg.inv(su)[None], (T, 1, 1))
    mean = samplers._tvp_state_banded(Zt, sig_t, q, a0, y, np.ones(n), np.zeros(n * T)).reshape(n, T, order="F")
    np.testing.assert_allclose(sm[:, :T], mean, rtol=1e-8, atol=1e-10)


# ---- second, independent restatement (plain C, dense) of the blocks only one restatement covered ---------------------

@pytest.mark.parametrize("table", ["ksc1998", "ocsn2007"])
def test_c_stochvol_equals_numpy_oracle(table):
    """src/stochvol_ksc1998.cpp:60-108 / stochvol_ocsn2007.cpp:60-114 read twice: NumPy (oracle/blocks.py) and plain C
    (oracle/blocks_ref.c) on the same injected uniforms / normals; indicators exact, h to 1e-12."""
    from oracle import blocks_ref
    rng = np.random.default_rng(5)
    T, k = 73, 3
    y = rng.standard_normal((T, k)) * np.array([0.3, 1.0, 2.5])
    h0 = np.log(np.tile(y.var(axis=0), (T, 1))) + 0.1 * rng.standard_normal((T, k))
    sigma, const = np.array([0.02, 0.05, 0.4]), np.array([1e-4, 1e-3, 1e-5])
    u, eps = rng.random((k, T)), rng.standard_normal((k, T))
    want_h, want_s = blocks.stochvol_mixture(y, h0, sigma, h0[0], const, u, eps, table=table, return_s=True)
    got_h, got_s = blocks_ref.stochvol(y, h0, sigma, h0[0], const, u, eps, table=table)
    assert np.array_equal(got_s, want_s)
    assert helpers.relerr(got_h, want_h) < 1e-12


@pytest.mark.parametrize("tvp,tv_sigma", [(False, False), (True, False), (False, True), (True, True)])
def test_c_bvs_equals_numpy_oracle(tvp, tv_sigma):
    """src/bvs.cpp:73-159 read twice; constant / time-varying a, constant / per-period Sigma^-1, include subset.  The
    decisions must agree exactly over many calls (each call visits up to n sites)."""
    from oracle import blocks_ref
    rng = np.random.default_rng(11 + 2 * tvp + tv_sigma)
    k, T, n = 3, 40, 12
    mismatches = sites = 0
    for rep in range(40):
        z = rng.standard_normal((k * T, n))
        a = rng.standard_normal((n, T)) * 0.3 if tvp else rng.standard_normal(n) * 0.3
        lam = np.diag((rng.random(n) < 0.6).astype(np.float64))
        th = (np.diag(lam)[:, None] * a) if tvp else np.diag(lam) * a
        mean = np.stack([z[t * k:(t + 1) * k] @ (th[:, t] if tvp else th) for t in range(T)], axis=1)
        y = mean + 0.5 * rng.standard_normal((k, T))
        if tv_sigma:
            blocks_t = []
            for t in range(T):
                w = rng.standard_normal((k, k + 2))
                blocks_t.append(np.linalg.inv(w @ w.T / (k + 2)))
            sig = np.vstack(blocks_t)
        else:
            w = rng.standard_normal((k, k + 2))
            sig = np.linalg.inv(w @ w.T / (k + 2))
        prior = rng.uniform(0.2, 0.8, n)
        u1, u2 = rng.random(n), rng.random(n)
        include = None if rep % 2 == 0 else np.sort(rng.choice(n, size=7, replace=False)) + 1
        want = np.diag(blocks.bvs(y, z, a, lam, sig, prior, u1, u2, include=include))
        got = np.diag(blocks_ref.bvs(y, z, a, lam, sig, prior, u1, u2, include=include))
        mismatches += int((want != got).sum())
        sites += n if include is None else len(include)
    assert sites > 300 and mismatches == 0


@pytest.mark.parametrize("masked", [False, True])
def test_c_tvp_state_equals_numpy_oracle(masked):
    """src/bvartvpalg.cpp:290-297 read three ways: the literal dense NumPy form (H, kron, block_diag as written), the
    banded structured form of oracle/samplers.py and the plain-C dense form of oracle/blocks_ref.c."""
    import scipy.linalg as sla
    from oracle import blocks_ref
    rng = np.random.default_rng(21 + masked)
    k, T, M = 3, 25, 4
    n = k * M
    x = rng.standard_normal((M, T))
    Zt = [np.kron(x[:, t][None, :], np.eye(k)) for t in range(T)]                  # k x n
    lam = (rng.random(n) < 0.7).astype(np.float64) if masked else np.ones(n)
    sig_t = []
    for t in range(T):
        w = rng.standard_normal((k, k + 3))
        sig_t.append(np.linalg.inv(w @ w.T / (k + 3)))
    sig_t = np.array(sig_t)
    y = rng.standard_normal((k, T))
    sv_i = rng.uniform(5.0, 200.0, n)
    a_init = rng.standard_normal(n) * 0.2
    eps = rng.standard_normal(n * T)
    # literal dense form, as the reference writes it
    H = np.eye(n * T)
    idx = np.arange(n * (T - 1))
    H[idx + n, idx] = -1.0
    zz = sla.block_diag(*Zt) * np.tile(lam, T)[None, :]
    zzss = zz.T @ sla.block_diag(*sig_t)
    hsh = H.T @ np.kron(np.eye(T), np.diag(sv_i)) @ H
    post_v = hsh + zzss @ zz
    post_mu = sla.solve(post_v, hsh @ np.tile(a_init, T) + zzss @ y.reshape(-1, order="F"))
    lit = post_mu + sla.solve_triangular(sla.cholesky(post_v, lower=False), eps, lower=False)
    band = samplers._tvp_state_banded(Zt, sig_t, sv_i, a_init, y, lam, eps)
    got = blocks_ref.tvp_state(y, np.vstack(Zt) * lam[None, :], np.vstack(list(sig_t)), sv_i, a_init, eps)
    assert helpers.relerr(got, lit) < 1e-11
    assert helpers.relerr(got, band) < 1e-11
