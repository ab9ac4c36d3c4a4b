"""Model builders of the BASELINE configs and of the smaller variants the tests use.

Inputs only: data generation (synthetic, seeded) + the gen_var / add_priors mirror.  Used by bench.py and the tests."""
from __future__ import annotations

import numpy as np

from .data import e1_readme_sample
from .model import add_priors, gen_var


def model_c1(iterations=50, burnin=10, coef=None, sigma=None, **kw):
    x, names = e1_readme_sample()
    m = gen_var(x, p=2, deterministic="const", iterations=iterations, burnin=burnin, names=names, **{
        k: v for k, v in kw.items() if k in ("tvp", "sv")})
    return add_priors(m, coef=coef or dict(v_i=0, v_i_det=0), sigma=sigma or dict(df=1, scale=1e-4),
                      ssvs=kw.get("ssvs"), bvs=kw.get("bvs"))


def model_grid_c2b(iterations=5000, burnin=1000, p=(0, 1, 2, 3, 4)):
    """Config c2-B: the model-comparison grid of vignettes/model-comparison.Rmd:44-60 -- e1, p = 0:4 + const on the
    common sample (T = 71), priors v_i = 0, v_i_det = 0, df = "k", scale = 1e-4.  Returns the list of models."""
    x, names = e1_readme_sample()
    ms = gen_var(x, p=list(p), deterministic="const", iterations=iterations, burnin=burnin, names=names)
    return [add_priors(m, coef=dict(v_i=0, v_i_det=0), sigma=dict(df="k", scale=1e-4)) for m in ms]


def synth_var(K, p, nobs, seed, intercept=0.1):
    """Stationary Gaussian VAR(p) data in the style of SURVEY.md 8d (c3)."""
    rng = np.random.default_rng(seed)
    A = [0.5 * np.eye(K) + 0.05 * (np.ones((K, K)) - np.eye(K)), 0.2 * np.eye(K)] + [np.zeros((K, K))] * max(0, p - 2)
    A = A[:max(p, 1)]
    idx = np.arange(K)
    Sig = 0.5 * np.eye(K) + 0.5 * 0.3 ** np.abs(idx[:, None] - idx[None, :])
    Lc = np.linalg.cholesky(Sig)
    burn = 200
    y = np.zeros((nobs + burn, K))
    for t in range(len(A), nobs + burn):
        v = intercept + Lc @ rng.standard_normal(K)
        for l, Al in enumerate(A):
            v = v + Al @ y[t - 1 - l]
        y[t] = v
    return y[burn:]


def model_c3(K=6, p=4, nobs=195, iterations=30, burnin=10, seed=195006, bvs=False):
    y = synth_var(K, p, nobs, seed)
    m = gen_var(y, p=p, deterministic="const", iterations=iterations, burnin=burnin)
    if bvs:
        return add_priors(m, coef=dict(v_i=1, v_i_det=0.1), sigma=dict(df="k", scale=1e-2),
                          bvs=dict(inprior=0.5, exclude_det=True))
    return add_priors(m, coef=dict(v_i=1, v_i_det=0.1), sigma=dict(df=0, scale=1e-5),
                      ssvs=dict(inprior=0.5, semiautomatic=(0.01, 10.0), exclude_det=True))


def model_sv(K=3, p=2, nobs=60, iterations=20, burnin=5, seed=320200, tvp=False, bvs=False, covar=False, psi_bvs=False):
    y = synth_var(K, p, nobs, seed)
    m = gen_var(y, p=p, deterministic="const", iterations=iterations, burnin=burnin, tvp=tvp, sv=True)
    return add_priors(m, coef=dict(v_i=1, v_i_det=0.1, shape=3, rate=1e-4, rate_det=0.01),
                      sigma=dict(shape=3, rate=1e-4, mu=0, v_i=0.01, sigma_h=0.05, constant=1e-4, covar=covar),
                      bvs=dict(inprior=0.5, exclude_det=True, covar=psi_bvs) if (bvs or psi_bvs) else None)


def model_covar_gamma(K=3, p=2, nobs=70, iterations=20, burnin=5, seed=77, bvs=False, psi_sel=None):
    """Covariance (psi) block with inverse-gamma variances (sigma = list(shape, rate, covar = TRUE)); psi_sel =
    "ssvs" / "bvs": variable selection on the covariances as well (ssvs$covar / bvs$covar)."""
    y = synth_var(K, p, nobs, seed)
    m = gen_var(y, p=p, deterministic="const", iterations=iterations, burnin=burnin)
    ssvs = dict(inprior=0.5, tau=(0.05, 10.0), exclude_det=True, covar=True) if psi_sel == "ssvs" else None
    bv = dict(inprior=0.5, exclude_det=True, covar=(psi_sel == "bvs")) if (bvs or psi_sel == "bvs") else None
    return add_priors(m, coef=dict(v_i=1, v_i_det=0.1), sigma=dict(shape=3, rate=1e-4, covar=True), ssvs=ssvs, bvs=bv)


def model_tvp_covar(K=3, p=1, nobs=24, iterations=8, burnin=3, seed=41, sv=True, bvs=False, psi_bvs=False):
    """TVP-VAR with time-varying covariances (Psi_t random walks, bvartvpalg.cpp:339-407) and SV or inverse-gamma
    measurement variances."""
    y = synth_var(K, p, nobs, seed)
    m = gen_var(y, p=p, deterministic="const", iterations=iterations, burnin=burnin, tvp=True, sv=sv)
    sigma = dict(shape=3, rate=1e-4, covar=True)
    if sv:
        sigma.update(mu=0, v_i=0.01, sigma_h=0.05, constant=1e-4)
    return add_priors(m, coef=dict(v_i=1, v_i_det=0.1, shape=3, rate=1e-4, rate_det=0.01), sigma=sigma,
                      bvs=dict(inprior=0.5, exclude_det=True, covar=psi_bvs) if (bvs or psi_bvs) else None)


def model_structural(K=3, p=1, nobs=50, iterations=12, burnin=4, seed=9, sv=False, bvs=False, tvp=False, ssvs=False):
    """Structural VAR (A0 coefficients: equation i contains -y_j, j < i) with inverse-gamma or SV error variances."""
    y = synth_var(K, p, nobs, seed)
    m = gen_var(y, p=p, deterministic="const", iterations=iterations, burnin=burnin, structural=True, sv=sv, tvp=tvp)
    sigma = dict(shape=3, rate=1e-3)
    if sv:
        sigma.update(mu=0, v_i=0.01, sigma_h=0.05, constant=1e-4)
    return add_priors(m, coef=dict(v_i=1, v_i_det=0.1, shape=3, rate=1e-4, rate_det=0.01), sigma=sigma,
                      bvs=dict(inprior=0.5, exclude_det=True) if bvs else None,
                      ssvs=dict(inprior=0.5, semiautomatic=(0.1, 10), exclude_det=True) if ssvs else None)


def model_c5(K=20, p=4, nobs=204, iterations=200, burnin=50, seed=2042004):
    """Config c5: large BVAR, K = 20, p = 4, const (n = 1620), Minnesota prior, sigma df = k, scale 1.
    DGP: diagonally dominant stable VAR(1), AR(1)-correlated errors."""
    rng = np.random.default_rng(seed)
    A1 = 0.4 * np.eye(K) + rng.uniform(-0.02, 0.02, (K, K)) * (1 - np.eye(K))
    idx = np.arange(K)
    Lc = np.linalg.cholesky(0.3 ** np.abs(idx[:, None] - idx[None, :]))
    y = np.zeros((nobs + 100, K))
    for t in range(1, nobs + 100):
        y[t] = 0.1 + A1 @ y[t - 1] + Lc @ rng.standard_normal(K)
    m = gen_var(y[100:], p=p, deterministic="const", iterations=iterations, burnin=burnin)
    return add_priors(m, coef=dict(minnesota=dict(kappa0=2, kappa1=0.5, kappa3=5)), sigma=dict(df="k", scale=1))


def baseline_workload(name, iterations=None, burnin=None):
    """(model, label) of a BASELINE.json config at its full size unless iterations / burnin are overridden."""
    if name == "c2a":
        it, bi = iterations or 5000, 1000 if burnin is None else burnin
        return model_c1(iterations=it, burnin=bi), \
            "c2-A: e1 VAR(2)+const, independent normal-Wishart (v_i=0, df=1, scale=1e-4), %d iter / %d burn-in" % (it, bi)
    if name == "c3":
        it, bi = iterations or 5000, 1000 if burnin is None else burnin
        return model_c3(iterations=it, burnin=bi), \
            "c3: synthetic VAR(4) K=6 T=191 SSVS (semiautomatic 0.01/10), %d iter / %d burn-in" % (it, bi)
    if name == "c4":
        it, bi = iterations or 500, 500 if burnin is None else burnin
        return model_sv(K=3, p=2, nobs=202, iterations=it, burnin=bi, tvp=True), \
            "c4: synthetic TVP-VAR(2) K=3 T=200 + KSC-1998 SV, %d iter / %d burn-in" % (it, bi)
    if name == "c5":
        it, bi = iterations or 200, 50 if burnin is None else burnin
        return model_c5(iterations=it, burnin=bi), \
            "c5: synthetic large BVAR K=20 p=4 T=200 (n=1620), Minnesota prior, %d iter / %d burn-in" % (it, bi)
    raise ValueError(f"unknown workload {name}")
