"""Host-side mirror of the reference's stand-alone sampler building blocks, calling the CUDA library.

Same names, argument meaning and error behaviour as the exported functions of the reference (the ones users call
from their own R-level Gibbs loops, README.md:201-216, vignettes/ssvs.Rmd:98-121):

  post_normal(y, x, sigma_i, a_prior, v_i_prior)                  src/post_normal.cpp:61-93
  post_normal_sur(y, z, sigma_i, a_prior, v_i_prior, svd=False)   src/post_normal_sur.cpp:60-113
  ssvs(a, tau0, tau1, prob_prior, include=None)                   src/svss.cpp:76-111
  bvs(y, z, a, lambda_, sigma_i, prob_prior, include=None)        src/bvs.cpp:73-159
  stochvol_ksc1998 / stochvol_ocsn2007 / stoch_vol(y, h, sigma, h_init, constant)
                                                                  src/stochvol_ksc1998.cpp:60-108, _ocsn2007.cpp:60-114
  kalman_dk(y, z, sigma_u, sigma_v, B, a_init, P_init)            src/kalman_dk.cpp:91-184

On top of the reference's one-problem-per-call interface every function is *batched*: pass the per-draw arguments
(sigma_i, a, lambda, h, ...) with one extra leading axis of length B and B independent problems are solved in one
kernel launch (one CTA per problem).  `seed` keys the Philox stream; `eps` / `u` inject the random variates
(tier-1 parity tests).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._lib import BvgError, check, lib, require_device

__all__ = ["post_normal", "post_normal_sur", "ssvs", "bvs", "stochvol_ksc1998", "stochvol_ocsn2007", "stoch_vol",
           "kalman_dk"]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def _fn(name):
    return getattr(lib(), "bvg_" + name)


def _seed(seed):
    return int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & (2 ** 64 - 1)


def _cm(a):
    """column-major flat copy of a matrix (Armadillo layout)"""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, order="F"))


def _cmb(a, ndim):
    """[B] stack of column-major matrices; a single problem (a.ndim == ndim) becomes B = 1"""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == ndim:
        a = a[None]
    if a.ndim != ndim + 1:
        raise ValueError("wrong number of dimensions")
    flat = np.ascontiguousarray(np.stack([x.reshape(-1, order="F") for x in a]))
    return flat, a.shape[0], a.shape[1:]


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _include(include, n):
    if include is None:
        return None, 0
    inc = np.ascontiguousarray(np.asarray(include, dtype=np.int64).reshape(-1) - 1, dtype=np.int32)   # R is 1-based
    if inc.size and (inc.min() < 0 or inc.max() >= n):
        raise BvgError(-8, "include index out of range")
    return inc, inc.shape[0]


def _bcast(flat, B):
    return flat if flat.shape[0] == B else np.ascontiguousarray(np.broadcast_to(flat, (B,) + flat.shape[1:]))


def post_normal(y, x, sigma_i, a_prior, v_i_prior, *, seed=None, eps=None, method="eig"):
    """Posterior draw of the VAR coefficients (src/post_normal.cpp:61-93).  y k x T, x M x T, sigma_i k x k (or
    [B] x k x k), a_prior n, v_i_prior n x n.  Returns n x 1 (or B x n).  method "eig" = the reference's symmetric
    eigen square root, "chol" = the Cholesky mapping of the built-in samplers (bvaralg.cpp:306-307)."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    if y.shape[1] != x.shape[1]:
        raise ValueError("Arguments 'x' and 'y' contain different amounts of observations.")       # :66-68
    k, T = y.shape
    M = x.shape[0]
    n = k * M
    a_prior = np.asarray(a_prior, dtype=np.float64).reshape(-1)
    v_i_prior = np.asarray(v_i_prior, dtype=np.float64)
    if a_prior.shape[0] != n:
        raise ValueError("Argument 'a_prior' does not contain the required amount of elements.")   # :75-77
    if v_i_prior.shape[0] != n:
        raise ValueError("Argument 'v_i_prior' does not contain the required amount of elements.")  # :78-80
    sig, B, _ = _cmb(sigma_i, 2)
    single = np.asarray(sigma_i).ndim == 2
    if np.isnan(sig).any():
        raise BvgError(-6, "Argument 'sigma_i' contains NAs.")                                     # :63-65
    require_device()
    e = None if eps is None else np.ascontiguousarray(np.asarray(eps, dtype=np.float64).reshape(B, n))
    out = np.empty((B, n))
    check(_fn("post_normal")(B, k, T, M, _p(_cm(y)), _p(_cm(x)), _p(sig), _p(np.ascontiguousarray(a_prior)),
                                _p(_cm(v_i_prior)), _p(e), _seed(seed), 1 if method == "chol" else 0, _p(out)))
    return out[0].reshape(n, 1) if single else out


def post_normal_sur(y, z, sigma_i, a_prior, v_i_prior, svd=False, *, seed=None, eps=None, method="eig"):
    """SUR form (src/post_normal_sur.cpp:60-113).  y k x T, z kT x n, sigma_i k x k or kT x k (time varying), or
    [B] stacks of either.  `svd=True` (U sqrt(s) V' of the SPD posterior covariance) equals the eigen square root."""
    y = np.asarray(y, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    k, T = y.shape
    n = z.shape[1]
    a_prior = np.asarray(a_prior, dtype=np.float64).reshape(-1)
    v_i_prior = np.asarray(v_i_prior, dtype=np.float64)
    if a_prior.shape[0] != n:
        raise ValueError("Argument 'a_prior' does not contain the required amount of elements.")
    if v_i_prior.shape[0] != n:
        raise ValueError("Argument 'v_i_prior' does not contain the required amount of elements.")
    sig, B, shp = _cmb(sigma_i, 2)
    single = np.asarray(sigma_i).ndim == 2
    tv = 1 if shp[0] > k else 0                                                                    # :79-82
    if np.isnan(sig).any():
        raise BvgError(-6, "Argument 'sigma_i' contains NAs.")
    require_device()
    e = None if eps is None else np.ascontiguousarray(np.asarray(eps, dtype=np.float64).reshape(B, n))
    out = np.empty((B, n))
    check(_fn("post_normal_sur")(B, k, T, n, _p(_cm(y)), _p(_cm(z)), _p(sig), tv,
                                    _p(np.ascontiguousarray(a_prior)), _p(_cm(v_i_prior)), _p(e), _seed(seed),
                                    1 if method == "chol" else 0, _p(out)))
    return out[0].reshape(n, 1) if single else out


def ssvs(a, tau0, tau1, prob_prior, include=None, *, seed=None, u=None):
    """SSVS indicator draw (src/svss.cpp:76-111).  a n (or B x n).  Returns {"v_i": n x n diagonal matrix,
    "lambda": n} like the reference (batched: "v_i" is the B x n array of diagonals)."""
    a_in = np.asarray(a, dtype=np.float64)
    single = a_in.ndim == 1 or (a_in.ndim == 2 and a_in.shape[1] == 1)
    av = np.ascontiguousarray(a_in.reshape(1, -1) if single else a_in)
    B, n = av.shape
    t0 = np.ascontiguousarray(np.asarray(tau0, dtype=np.float64).reshape(-1))
    t1 = np.ascontiguousarray(np.asarray(tau1, dtype=np.float64).reshape(-1))
    pp = np.ascontiguousarray(np.asarray(prob_prior, dtype=np.float64).reshape(-1))
    if not (t0.shape[0] == t1.shape[0] == pp.shape[0] == n):
        raise ValueError("tau0, tau1 and prob_prior must have one element per coefficient")
    inc, ninc = _include(include, n)
    require_device()
    uu = None if u is None else np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(B, n))
    lam, vd = np.empty((B, n)), np.empty((B, n))
    check(_fn("ssvs")(B, n, _p(av), _p(t0), _p(t1), _p(pp), None if inc is None else inc.ctypes.data_as(_ip),
                         ninc, _p(uu), _seed(seed), _p(lam), _p(vd)))
    if single:
        return {"v_i": np.diag(vd[0]), "lambda": lam[0].reshape(n, 1)}
    return {"v_i": vd, "lambda": lam}


def bvs(y, z, a, lambda_, sigma_i, prob_prior, include=None, *, seed=None, u1=None, u2=None):
    """BVS indicator draw (src/bvs.cpp:73-159).  y k x T, z kT x n, a n x 1 or n x T (time varying), lambda_ n x n
    diagonal indicator matrix, sigma_i k x k or kT x k; [B] stacks of a / lambda_ / sigma_i for batches.
    Returns the updated n x n indicator matrix (batched: B x n diagonals)."""
    y = np.asarray(y, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    k, T = y.shape
    n = z.shape[1]
    a_in = np.asarray(a, dtype=np.float64)
    lam_in = np.asarray(lambda_, dtype=np.float64)
    single = lam_in.ndim == 2 and lam_in.shape == (n, n) and a_in.ndim <= 2
    if single:
        a_in = a_in.reshape(n, -1)[None]
        lam_d = np.ascontiguousarray(np.diag(lam_in)[None])
        sig, B, shp = _cmb(sigma_i, 2)
    else:
        if a_in.ndim == 2:
            a_in = a_in[:, :, None]
        lam_d = np.ascontiguousarray(lam_in if lam_in.ndim == 2 else np.stack([np.diag(m) for m in lam_in]))
        sig, B, shp = _cmb(sigma_i, 2)
    B = max(B, a_in.shape[0], lam_d.shape[0])
    a_tv = 1 if a_in.shape[2] > 1 else 0                                                           # :83-86
    if a_tv and a_in.shape[2] != T:
        raise ValueError("time-varying 'a' must have one column per observation")
    af = _bcast(np.ascontiguousarray(np.stack([m.reshape(-1, order="F") for m in a_in])), B)
    lam_d = _bcast(lam_d, B)
    sig = _bcast(sig, B)
    sigma_tv = 1 if shp[0] > k else 0                                                              # :88-91
    pp = np.ascontiguousarray(np.asarray(prob_prior, dtype=np.float64).reshape(-1))
    inc, ninc = _include(include, n)
    require_device()
    v1 = None if u1 is None else np.ascontiguousarray(np.asarray(u1, dtype=np.float64).reshape(B, n))
    v2 = None if u2 is None else np.ascontiguousarray(np.asarray(u2, dtype=np.float64).reshape(B, n))
    out = np.empty((B, n))
    check(_fn("bvs")(B, k, T, n, _p(_cm(y)), _p(_cm(z)), _p(af), a_tv, _p(lam_d), _p(sig), sigma_tv, _p(pp),
                        None if inc is None else inc.ctypes.data_as(_ip), ninc, _p(v1), _p(v2), _seed(seed), _p(out)))
    return np.diag(out[0]) if single else out


def _stochvol(y, h, sigma, h_init, constant, mixture, seed, u, eps, return_s):
    y_in = np.asarray(y, dtype=np.float64)
    h_in = np.asarray(h, dtype=np.float64)
    if np.isnan(y_in).any():
        raise BvgError(-6, "Argument 'y' contains NAs.")                                           # ksc1998:63-65
    if y_in.shape != h_in.shape:
        raise ValueError("Arguments 'y' and 'h' do not have the same dimensions.")                 # :66-71
    single = y_in.ndim == 2
    yf, B, (T, k) = _cmb(y_in, 2)
    hf, _, _ = _cmb(h_in, 2)
    sg = _bcast(np.ascontiguousarray(np.asarray(sigma, dtype=np.float64).reshape(-1, k)), B)
    hi = _bcast(np.ascontiguousarray(np.asarray(h_init, dtype=np.float64).reshape(-1, k)), B)
    cc = _bcast(np.ascontiguousarray(np.asarray(constant, dtype=np.float64).reshape(-1, k)), B)
    require_device()
    uu = None if u is None else np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(B, k * T))
    ee = None if eps is None else np.ascontiguousarray(np.asarray(eps, dtype=np.float64).reshape(B, k * T))
    out = np.empty((B, k * T))
    s = np.empty((B, k * T), dtype=np.int32) if return_s else None
    check(_fn("stochvol")(B, T, k, _p(yf), _p(hf), _p(sg), _p(hi), _p(cc), mixture, _p(uu), _p(ee), _seed(seed),
                             _p(out), None if s is None else s.ctypes.data_as(_ip)))
    hh = out.reshape(B, k, T).transpose(0, 2, 1)                     # col-major T x k per problem
    res = hh[0] if single else hh
    if return_s:
        ss = s.reshape(B, k, T).transpose(0, 2, 1)
        return res, (ss[0] if single else ss)
    return res


def stochvol_ksc1998(y, h, sigma, h_init, constant, *, seed=None, u=None, eps=None, return_s=False):
    """Kim-Shephard-Chib (1998) seven-component mixture step (src/stochvol_ksc1998.cpp:60-108).  y, h T x k (or
    B x T x k); sigma, h_init, constant k.  u / eps: k x T injected uniforms / normals per series."""
    return _stochvol(y, h, sigma, h_init, constant, 7, seed, u, eps, return_s)


def stochvol_ocsn2007(y, h, sigma, h_init, constant, *, seed=None, u=None, eps=None, return_s=False):
    """Omori-Chib-Shephard-Nakajima (2007) ten-component mixture step (src/stochvol_ocsn2007.cpp:60-114)."""
    return _stochvol(y, h, sigma, h_init, constant, 10, seed, u, eps, return_s)


def stoch_vol(y, h, sigma, h_init, constant, **kw):
    """Wrapper of stochvol_ksc1998 (src/stoch_vol.cpp:40-42)."""
    return stochvol_ksc1998(y, h, sigma, h_init, constant, **kw)


def kalman_dk(y, z, sigma_u, sigma_v, B, a_init, P_init, *, seed=None, eps=None):
    """Durbin-Koopman (2002) simulation smoother (src/kalman_dk.cpp:91-184).  y k x T, z kT x n, sigma_u k x k or
    kT x k, sigma_v n x n or nT x n, B n x n or nT x n, a_init n, P_init n x n ([batch] stacks of the last five for
    batches).  eps: n + T (k + n) normals in the reference's consumption order (eps0, then per t eps_y, eps_a).
    Returns n x (T+1) (or batch x n x (T+1))."""
    y = np.asarray(y, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    k, T = y.shape
    n = z.shape[1]
    single = np.asarray(P_init).ndim == 2
    su, b1, s1 = _cmb(sigma_u, 2)
    sv, b2, s2 = _cmb(sigma_v, 2)
    bm, b3, s3 = _cmb(B, 2)
    pi, b4, _ = _cmb(P_init, 2)
    ai = np.ascontiguousarray(np.asarray(a_init, dtype=np.float64).reshape(-1, n))
    nb = max(b1, b2, b3, b4, ai.shape[0])
    tv = (1 if s1[0] > k else 0) | (2 if s2[0] > n else 0) | (4 if s3[0] > n else 0)               # :99-121
    su, sv, bm, pi, ai = (_bcast(v, nb) for v in (su, sv, bm, pi, ai))
    require_device()
    neps = n + T * (k + n)
    ee = None if eps is None else np.ascontiguousarray(np.asarray(eps, dtype=np.float64).reshape(nb, neps))
    out = np.empty((nb, n * (T + 1)))
    check(_fn("kalman_dk")(nb, k, T, n, _p(_cm(y)), _p(_cm(z)), _p(su), _p(sv), _p(bm), tv, _p(ai), _p(pi),
                              _p(ee), _seed(seed), _p(out)))
    res = out.reshape(nb, T + 1, n).transpose(0, 2, 1)
    return res[0] if single else res


def loglik_normal(u, sigma):
    """Log-likelihood of a multivariate normal per period (src/loglik_normal.cpp:37-58, R/RcppExports.R:313):
    u is K x T, sigma K x K or the (K T) x K stack of per-period matrices; returns a length-T vector."""
    require_device()
    u = np.ascontiguousarray(np.asarray(u, dtype=np.float64))
    sigma = np.asarray(sigma, dtype=np.float64)
    if u.ndim != 2:
        raise ValueError("u must be a K x T matrix")
    k, T = u.shape
    if sigma.ndim != 2 or sigma.shape[1] != k or sigma.shape[0] not in (k, k * T):
        raise ValueError("sigma must be K x K or (K T) x K")
    out = np.empty(T, dtype=np.float64)
    uf, sf = _cm(u), _cm(sigma)
    check(lib().bvg_loglik_normal(_p(uf), _p(sf), k, T, sigma.shape[0], 1, _p(out)))
    return out

