"""Tier-1 parity of the batched stand-alone building blocks (post_normal, post_normal_sur, ssvs, bvs, stochvol_*,
kalman_dk) against the oracle restatement of the reference functions (oracle/blocks.py), on injected variates.

Every test runs twice: `gpu` = the CUDA kernels through the C ABI of libbvargpu.so (the parity test proper), and
`emu` = the same group-generic *_problem templates compiled with g++ (tests/emu, test-only) so that the arithmetic
is also checked on machines without a GPU.  Bar: continuous outputs <= 1e-10 relative, indicators bit-exact."""
import ctypes as C

import numpy as np
import pytest

import helpers
from bvartools_b200 import blocks
from oracle import blocks as ob

TOL = 1e-10


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def blk(request, monkeypatch):
    if request.param == "gpu":
        request.getfixturevalue("gpu")
        return blocks
    emu = request.getfixturevalue("emu")
    from bvartools_b200._lib import lib
    L = lib()

    def fn(name):
        f = getattr(emu, "emu_" + name)
        f.argtypes = getattr(L, "bvg_" + name).argtypes
        f.restype = C.c_int
        return f
    monkeypatch.setattr(blocks, "_fn", fn)
    monkeypatch.setattr(blocks, "require_device", lambda: None)
    return blocks


def _var_data(k=3, p=2, nobs=50, seed=5):
    y = helpers.synth_var(k, p, nobs, seed)
    m = helpers.gen_var(y, p=p, deterministic="const", iterations=5, burnin=1)
    Y, Z, SUR = m["data"]["Y"], m["data"]["Z"], m["data"]["SUR"]
    return Y.T.copy(), Z.T.copy(), SUR                     # y k x T, x M x T, z kT x n


def _spd(rng, k, B=None):
    def one():
        a = rng.standard_normal((k, k + 2))
        return a @ a.T / (k + 2) + 0.3 * np.eye(k)
    return one() if B is None else np.stack([one() for _ in range(B)])


@pytest.mark.parametrize("method", ["eig", "chol"])
def test_post_normal(blk, method):
    rng = np.random.default_rng(1)
    y, x, _ = _var_data()
    k, M = y.shape[0], x.shape[0]
    n = k * M
    B = 4
    sig = _spd(rng, k, B)
    a_prior = rng.standard_normal(n) * 0.1
    v = rng.standard_normal((n, n)) * 0.05
    v_i = v @ v.T + np.diag(rng.uniform(0.5, 2.0, n))         # dense prior precision
    eps = rng.standard_normal((B, n))
    got = blk.post_normal(y, x, sig, a_prior, v_i, eps=eps, method=method)
    for b in range(B):
        if method == "eig":
            want = ob.post_normal(y, x, sig[b], a_prior, v_i, eps[b])
        else:                                               # bvaralg.cpp:306-307 mapping
            import scipy.linalg as sla
            P = v_i + np.kron(x @ x.T, sig[b])
            rhs = v_i @ a_prior + (sig[b] @ y @ x.T).reshape(-1, order="F")
            want = sla.solve(P, rhs) + sla.solve_triangular(sla.cholesky(P, lower=False), eps[b], lower=False)
        assert helpers.relerr(got[b], want) < TOL
    one = blk.post_normal(y, x, sig[0], a_prior, v_i, eps=eps[0], method=method)
    assert one.shape == (n, 1) and np.array_equal(one[:, 0], got[0])


@pytest.mark.parametrize("tv", [False, True])
def test_post_normal_sur(blk, tv):
    rng = np.random.default_rng(2)
    y, _, z = _var_data(nobs=40)
    k, T = y.shape
    n = z.shape[1]
    B = 3
    if tv:
        sig = np.stack([np.concatenate([_spd(rng, k) for _ in range(T)], axis=0) for _ in range(B)])    # kT x k
    else:
        sig = _spd(rng, k, B)
    a_prior = rng.standard_normal(n) * 0.1
    v_i = np.diag(rng.uniform(0.5, 2.0, n))
    eps = rng.standard_normal((B, n))
    got = blk.post_normal_sur(y, z, sig, a_prior, v_i, eps=eps)
    for b in range(B):
        assert helpers.relerr(got[b], ob.post_normal_sur(y, z, sig[b], a_prior, v_i, eps[b])) < TOL
        assert helpers.relerr(got[b], ob.post_normal_sur(y, z, sig[b], a_prior, v_i, eps[b], svd=True)) < 1e-9


def test_post_normal_matches_sur_form(blk):
    """post_normal (Kronecker form) == post_normal_sur with z = kron(x', I_k) on the same variates."""
    rng = np.random.default_rng(3)
    y, x, z = _var_data()
    k, n = y.shape[0], z.shape[1]
    sig = _spd(rng, k, 2)
    eps = rng.standard_normal((2, n))
    v_i = np.diag(rng.uniform(0.1, 1.0, n))
    a = blk.post_normal(y, x, sig, np.zeros(n), v_i, eps=eps)
    b = blk.post_normal_sur(y, z, sig, np.zeros(n), v_i, eps=eps)
    assert helpers.relerr(a, b) < TOL


@pytest.mark.parametrize("include", [None, "some"])
def test_ssvs(blk, include):
    rng = np.random.default_rng(4)
    n, B = 37, 50
    a = rng.standard_normal((B, n)) * rng.choice([0.01, 1.0], size=(B, n))
    tau0 = rng.uniform(0.01, 0.1, n)
    tau1 = tau0 * rng.uniform(5, 100, n)
    pp = rng.uniform(0.2, 0.8, n)
    u = rng.random((B, n))
    inc = None if include is None else np.sort(rng.choice(np.arange(1, n + 1), 20, replace=False))
    got = blk.ssvs(a, tau0, tau1, pp, include=inc, u=u)
    flips = 0
    for b in range(B):
        want = ob.ssvs(a[b], tau0, tau1, pp, u[b], include=inc)
        assert np.array_equal(got["lambda"][b], want["lambda"])                       # indicators bit-exact
        assert np.array_equal(got["v_i"][b], np.diag(want["v_i"]))
        flips += int((want["lambda"] == 0).sum())
    assert 0 < flips < B * n
    one = blk.ssvs(a[0], tau0, tau1, pp, include=inc, u=u[0])
    assert one["v_i"].shape == (n, n) and one["lambda"].shape == (n, 1)
    assert np.array_equal(np.diag(one["v_i"]), got["v_i"][0])


@pytest.mark.parametrize("a_tv,sigma_tv", [(False, False), (False, True), (True, False), (True, True)])
def test_bvs(blk, a_tv, sigma_tv):
    rng = np.random.default_rng(6)
    y, _, z = _var_data(nobs=36)
    k, T = y.shape
    n = z.shape[1]
    B = 24
    ols = np.linalg.lstsq(z, y.reshape(-1, order="F"), rcond=None)[0]
    if a_tv:
        a = ols[None, :, None] + 0.05 * rng.standard_normal((B, n, T))
    else:
        a = ols[None, :] + 0.05 * rng.standard_normal((B, n))
    a[:, rng.choice(n, 6, replace=False)] *= 0.02            # some irrelevant coefficients
    lam = (rng.random((B, n)) < 0.7).astype(np.float64)
    if sigma_tv:
        sig = np.stack([np.concatenate([_spd(rng, k) for _ in range(T)], axis=0) for _ in range(B)])
    else:
        sig = _spd(rng, k, B)
    pp = rng.uniform(0.3, 0.7, n)
    inc = np.arange(1, n - k + 1)                            # exclude deterministic terms
    u1, u2 = rng.random((B, n)), rng.random((B, n))
    got = blk.bvs(y, z, a if a_tv else a[:, :, None], lam, sig, pp, include=inc, u1=u1, u2=u2)
    changed = 0
    for b in range(B):
        want = ob.bvs(y, z, a[b] if a_tv else a[b][:, None], np.diag(lam[b]), sig[b], pp, u1[b], u2[b], include=inc)
        assert np.array_equal(got[b], np.diag(want)), b                                # indicators bit-exact
        changed += int((np.diag(want) != lam[b]).sum())
    assert changed > 0
    one = blk.bvs(y, z, a[0] if a_tv else a[0][:, None], np.diag(lam[0]), sig[0], pp, include=inc, u1=u1[0], u2=u2[0])
    assert one.shape == (n, n) and np.array_equal(np.diag(one), got[0])


@pytest.mark.parametrize("table,T", [("ksc1998", 73), ("ocsn2007", 73), ("ksc1998", 200)])
def test_stochvol(blk, table, T):
    rng = np.random.default_rng(7)
    k, B = 3, 4
    h_true = np.cumsum(0.1 * rng.standard_normal((B, T, k)), axis=1) - 1.0
    y = np.exp(0.5 * h_true) * rng.standard_normal((B, T, k))
    h = h_true + 0.2 * rng.standard_normal((B, T, k))
    sigma = rng.uniform(0.01, 0.1, (B, k))
    h_init = h[:, 0, :] + 0.1
    const = np.full((B, k), 1e-4)
    u, eps = rng.random((B, k, T)), rng.standard_normal((B, k, T))
    fn = blk.stochvol_ksc1998 if table == "ksc1998" else blk.stochvol_ocsn2007
    got, s = fn(y, h, sigma, h_init, const, u=u, eps=eps, return_s=True)
    for b in range(B):
        want, ws = ob.stochvol_mixture(y[b], h[b], sigma[b], h_init[b], const[b], u[b], eps[b], table=table,
                                       return_s=True)
        assert np.array_equal(s[b], ws)                                               # mixture components bit-exact
        assert helpers.relerr(got[b], want) < TOL
    one = blk.stoch_vol(y[0], h[0], sigma[0], h_init[0], const[0], u=u[0], eps=eps[0])
    if table == "ksc1998":
        assert np.array_equal(one, got[0])


@pytest.mark.parametrize("tv", [0, 7])
def test_kalman_dk(blk, tv):
    rng = np.random.default_rng(8)
    y, _, z = _var_data(k=2, p=1, nobs=31)
    k, T = y.shape
    n = z.shape[1]
    B = 3

    def stack(fn, rows):
        return np.stack([np.concatenate([fn() for _ in range(T)], axis=0) if tv else fn() for _ in range(B)])
    su = stack(lambda: _spd(rng, k), k)
    sv = stack(lambda: 0.01 * _spd(rng, n), n)
    Bm = stack(lambda: np.eye(n) * 0.95 + 0.01 * rng.standard_normal((n, n)), n)
    a_init = rng.standard_normal((B, n)) * 0.1
    P_init = np.stack([_spd(rng, n) for _ in range(B)])
    neps = n + T * (k + n)
    eps = rng.standard_normal((B, neps))
    got = blk.kalman_dk(y, z, su, sv, Bm, a_init, P_init, eps=eps)
    assert got.shape == (B, n, T + 1)
    for b in range(B):
        e0 = eps[b, :n]
        rest = eps[b, n:].reshape(T, k + n)
        want = ob.kalman_dk(y, z, su[b], sv[b], Bm[b], a_init[b], P_init[b], e0, rest[:, :k], rest[:, k:])
        assert helpers.relerr(got[b], want) < TOL
    one = blk.kalman_dk(y, z, su[0], sv[0], Bm[0], a_init[0], P_init[0], eps=eps[0])
    assert one.shape == (n, T + 1) and np.array_equal(one, got[0])


def test_error_behaviour_matches_reference(blk):
    rng = np.random.default_rng(9)
    y, x, z = _var_data()
    k, n = y.shape[0], z.shape[1]
    sig = _spd(rng, k)
    bad = sig.copy()
    bad[0, 1] = np.nan
    with pytest.raises(Exception, match="sigma_i.*NAs"):
        blk.post_normal(y, x, bad, np.zeros(n), np.eye(n))
    with pytest.raises(ValueError, match="a_prior"):
        blk.post_normal(y, x, sig, np.zeros(n - 1), np.eye(n))
    with pytest.raises(ValueError, match="v_i_prior"):
        blk.post_normal_sur(y, z, sig, np.zeros(n), np.eye(n + 1))
    yy = np.ones((10, 2))
    yy[3, 0] = np.nan
    with pytest.raises(Exception, match="'y' contains NAs"):
        blk.stochvol_ksc1998(yy, np.zeros((10, 2)), [0.05, 0.05], [0, 0], [1e-4, 1e-4])
    with pytest.raises(ValueError, match="same dimensions"):
        blk.stochvol_ksc1998(np.ones((10, 2)), np.zeros((9, 2)), [0.05, 0.05], [0, 0], [1e-4, 1e-4])


@pytest.mark.gpu
def test_blocks_philox_statistics(gpu):
    """Philox path of the blocks: post_normal draws have the analytic posterior mean / covariance; the eig and
    chol mappings give the same distribution; ssvs inclusion frequency equals the posterior probability."""
    rng = np.random.default_rng(10)
    y, x, _ = _var_data()
    k, M = y.shape[0], x.shape[0]
    n = k * M
    sig = _spd(rng, k)
    v_i = np.eye(n) * 0.5
    B = 20000
    V = np.linalg.inv(v_i + np.kron(x @ x.T, sig))
    mean = V @ (sig @ y @ x.T).reshape(-1, order="F")
    for method in ("eig", "chol"):
        d = blocks.post_normal(y, x, np.broadcast_to(sig, (B, k, k)), np.zeros(n), v_i, seed=5, method=method)
        se = np.sqrt(np.diag(V) / B)
        assert np.all(np.abs(d.mean(axis=0) - mean) < 4.5 * se)
        assert np.allclose(np.cov(d.T), V, rtol=0.15, atol=0.05 * np.diag(V).max())
    a = np.full((B, 4), 0.05)
    r = blocks.ssvs(a, np.full(4, 0.02), np.full(4, 0.5), np.full(4, 0.5), seed=3)
    p = ob.ssvs(a[0], np.full(4, 0.02), np.full(4, 0.5), np.full(4, 0.5), np.full(4, 0.5))["p_post"]
    assert np.all(np.abs(r["lambda"].mean(axis=0) - p) < 4.5 * np.sqrt(p * (1 - p) / B))
