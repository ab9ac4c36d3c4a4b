"""Injected-variate conventions shared by the oracle and the CUDA path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference consumes R's / Armadillo's sequential RNG streams, which cannot be reproduced without R.  Parity
is therefore defined on *sites*: each random variate a sweep needs is addressed by (sweep, block, index) and the
oracle and the GPU kernels consume the same arrays.  Shapes (S = iterations + burnin sweeps):

  coef_normal     (S, n)          N(0,1)   coefficient draw, index = coefficient id j*k+i  (bvaralg.cpp:307)
  ssvs_u          (S, n)          U(0,1)   SSVS Bernoulli, indexed by coefficient id        (bvaralg.cpp:322)
  bvs_u1, bvs_u2  (S, n)          U(0,1)   BVS prior gate / acceptance, by coefficient id   (bvaralg.cpp:338,351)
  sv_u            (S, k, T)       U(0,1)   mixture indicator, per series                    (stochvol_ksc1998.cpp:96)
  sv_normal       (S, k, T)       N(0,1)   log-volatility path                              (stochvol_ksc1998.cpp:104)
  sigma_h_gamma   (S, k)          Gamma(shape_post_i, 1), multiplied by the scale, inverted (bvaralg.cpp:470)
  h_init_normal   (S, k)          N(0,1)                                                    (bvaralg.cpp:477)
  omega_gamma     (S, k)          Gamma(shape_post_i, 1) for the inverse-gamma variances    (bvaralg.cpp:484)
  wishart_chi2    (S, k)          chi^2(df - i), i = 0..k-1                                 (bvaralg.cpp:511)
  wishart_normal  (S, k(k-1)/2)   N(0,1)   strict upper triangle of the Bartlett factor, column by column
  state_normal    (S, n*T)        N(0,1)   TVP state path, time-major                       (bvartvpalg.cpp:297)
  sigma_v_gamma   (S, n)          Gamma(shape_post_i, 1) (precision, not inverted)          (bvartvpalg.cpp:476)
  a_init_normal   (S, n)          N(0,1)                                                    (bvartvpalg.cpp:482)

Conventions of the un-vendored back ends (from memory of their sources; PARITY UNPINNED):
  * Rcpp::rbinom(1, 1, p) (R nmath inversion): lambda = 1{u < p} if p > 0.5 else 1{u >= 1 - p}.
  * arma::wishrnd(S, df): D = chol(S) (upper), A upper triangular with A_ii = sqrt(chi2(df - i)) and N(0,1)
    above the diagonal filled column by column, W = (A D)' (A D).
  * arma::randg(distr_param(a, b)) = b * Gamma(a, 1)  (shape a, scale b).
  * arma::shuffle of the inclusion index has no effect on the result, so the
    uniforms are indexed by coefficient id rather than by loop position.
"""
from __future__ import annotations

import numpy as np


def bernoulli_rbinom(p, u):
    """R's rbinom(1, 1, p) driven by one uniform (see module docstring)."""
    p = np.asarray(p, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    return np.where(p > 0.5, u < p, u >= 1.0 - p).astype(np.float64)


class VariateSource:
    """Hands out per-sweep variates: from injected arrays when present, else from a NumPy Generator."""

    def __init__(self, injected: dict | None = None, rng: np.random.Generator | None = None):
        self.injected = injected or {}
        self.rng = rng
        if not self.injected and rng is None:
            self.rng = np.random.default_rng(0)

    def _get(self, name, sweep, shape, sampler):
        if name in self.injected:
            v = np.asarray(self.injected[name][sweep], dtype=np.float64)
            assert v.shape == tuple(shape), (name, v.shape, shape)
            return v
        if self.rng is None:
            raise KeyError(f"variate block '{name}' not injected and no rng given")
        return sampler(self.rng)

    def normal(self, name, sweep, shape):
        return self._get(name, sweep, shape, lambda r: r.standard_normal(shape))

    def uniform(self, name, sweep, shape):
        return self._get(name, sweep, shape, lambda r: r.random(shape))

    def gamma(self, name, sweep, shape_param):
        shape_param = np.asarray(shape_param, dtype=np.float64)
        return self._get(name, sweep, shape_param.shape, lambda r: r.standard_gamma(shape_param))

    def chi2(self, name, sweep, df):
        df = np.asarray(df, dtype=np.float64)
        return self._get(name, sweep, df.shape, lambda r: r.chisquare(df))


def draw_variates(dims: dict, sweeps: int, rng: np.random.Generator, kinds) -> dict:
    """Pre-draw injected arrays for `sweeps` sweeps.  dims: n, k, T, shape_h, shape_omega, df, shape_v."""
    n, k, T = dims["n"], dims["k"], dims["T"]
    out = {}
    for kind in kinds:
        if kind == "coef_normal":
            out[kind] = rng.standard_normal((sweeps, n))
        elif kind in ("ssvs_u", "bvs_u1", "bvs_u2"):
            out[kind] = rng.random((sweeps, n))
        elif kind == "sv_u":
            out[kind] = rng.random((sweeps, k, T))
        elif kind == "sv_normal":
            out[kind] = rng.standard_normal((sweeps, k, T))
        elif kind == "sigma_h_gamma":
            out[kind] = rng.standard_gamma(np.broadcast_to(dims["shape_h"], (sweeps, k)))
        elif kind == "h_init_normal":
            out[kind] = rng.standard_normal((sweeps, k))
        elif kind == "omega_gamma":
            out[kind] = rng.standard_gamma(np.broadcast_to(dims["shape_omega"], (sweeps, k)))
        elif kind == "wishart_chi2":
            out[kind] = rng.chisquare(np.broadcast_to(dims["df"] - np.arange(k), (sweeps, k)))
        elif kind == "wishart_normal":
            out[kind] = rng.standard_normal((sweeps, k * (k - 1) // 2))
        elif kind == "psi_normal":
            out[kind] = rng.standard_normal((sweeps, k * (k - 1) // 2 * (T if dims.get("tvp_psi") else 1)))
        elif kind == "psi_sigma_v_gamma":
            out[kind] = rng.standard_gamma(np.broadcast_to(dims["shape_psi"], (sweeps, k * (k - 1) // 2)))
        elif kind == "psi_init_normal":
            out[kind] = rng.standard_normal((sweeps, k * (k - 1) // 2))
        elif kind in ("psi_u1", "psi_u2"):
            out[kind] = rng.random((sweeps, k * (k - 1) // 2))
        elif kind == "state_normal":
            out[kind] = rng.standard_normal((sweeps, n * T))
        elif kind == "sigma_v_gamma":
            out[kind] = rng.standard_gamma(np.broadcast_to(dims["shape_v"], (sweeps, n)))
        elif kind == "a_init_normal":
            out[kind] = rng.standard_normal((sweeps, n))
        else:
            raise KeyError(kind)
    return out
