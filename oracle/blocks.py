"""Oracle restatement of the stand-alone sampler building blocks -- TEST INFRASTRUCTURE ONLY.

Each function follows the cited reference lines literally (same dense matrices, same order of operations) and
takes its random variates as arguments.  See oracle/__init__.py for the pinning status.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .variates import bernoulli_rbinom

# --- mixture tables -------------------------------------------------------------------------------------------
# KSC-1998 seven-component table, means already shifted by -1.2704 (src/stochvol_ksc1998.cpp:74-78)
KSC_P = np.array([0.00730, 0.10556, 0.00002, 0.04395, 0.34001, 0.24566, 0.25750])
KSC_MU = np.array([-10.12999, -3.97281, -8.56686, 2.77786, 0.61942, 1.79518, -1.08819]) - 1.2704
KSC_S2 = np.array([5.79596, 2.61369, 5.17950, 0.16735, 0.64009, 0.34023, 1.26261])
# OCSN-2007 ten-component table (src/stochvol_ocsn2007.cpp:74-84), no shift
OCSN_P = np.array([0.00609, 0.04775, 0.13057, 0.20674, 0.22715, 0.18842, 0.12047, 0.05591, 0.01575, 0.00115])
OCSN_MU = np.array([1.92677, 1.34744, 0.73504, 0.02266, -0.85173, -1.97278, -3.46788, -5.55246, -8.68384,
                    -14.65000])
OCSN_S2 = np.array([0.11265, 0.17788, 0.26768, 0.40611, 0.62699, 0.98583, 1.57469, 2.54498, 4.16591, 7.33342])

MIXTURES = {"ksc1998": (KSC_P, KSC_MU, KSC_S2), "ocsn2007": (OCSN_P, OCSN_MU, OCSN_S2)}


def _normpdf(x, mu, sd):
    """arma::normpdf(x, mu, sd) element-wise: exp(-0.5 ((x-mu)/sd)^2) / (sd sqrt(2 pi))."""
    zz = (x - mu) / sd
    return np.exp(-0.5 * zz * zz) / (sd * np.sqrt(2.0 * np.pi))


def stochvol_mixture(y, h, sigma, h_init, constant, u, eps, table="ksc1998", return_s=False):
    """src/stochvol_ksc1998.cpp:60-108 (table='ksc1998') / src/stochvol_ocsn2007.cpp:60-114 (table='ocsn2007').

    y, h: T x k; sigma, h_init, constant: k; u, eps: k x T injected U(0,1) / N(0,1) per series.
    Returns the new h (T x k) [and the mixture indicators s (T x k) if return_s].
    """
    p_i, mu, sigma2 = MIXTURES[table]
    J = p_i.shape[0]
    y = np.array(y, dtype=np.float64)
    h = np.array(h, dtype=np.float64)
    if np.isnan(y).any():
        raise ValueError("Argument 'y' contains NAs.")
    if y.shape != h.shape:
        raise ValueError("Arguments 'y' and 'h' do not have the same dimensions.")
    tt, k = y.shape
    hh = np.eye(tt)
    hh[np.arange(1, tt), np.arange(tt - 1)] = -1.0
    hh = hh.T @ hh                                           # :85-87
    s_all = np.zeros((tt, k), dtype=np.int64)
    for i in range(k):
        ys = np.log(y[:, i] ** 2 + constant[i])              # :91
        q = p_i[None, :] * _normpdf(ys[:, None], h[:, i][:, None] + mu[None, :], np.sqrt(sigma2)[None, :])
        q = q / q.sum(axis=1, keepdims=True)                 # :95
        s = J - (u[i][:, None] < np.cumsum(q, axis=1)).sum(axis=1)   # :96
        s = np.minimum(s, J - 1)                             # out-of-bounds in the reference (UB) -> clamp
        s_all[:, i] = s
        sigh_hh = hh / sigma[i]                              # :99
        sigs = np.diag(1.0 / sigma2[s])                      # :100-101
        post_h_v = sigh_hh + sigs
        rhs = sigh_hh @ np.ones(tt) * h_init[i] + sigs @ (ys - mu[s])
        post_h_mu = sla.solve(post_h_v, rhs)                 # :103
        h[:, i] = post_h_mu + sla.solve_triangular(sla.cholesky(post_h_v, lower=False), eps[i], lower=False)
    return (h, s_all) if return_s else h


def post_normal(y, x, sigma_i, a_prior, v_i_prior, eps):
    """src/post_normal.cpp:61-93.  y k x T, x M x T; eps: n N(0,1).  Symmetric eig square root."""
    if np.isnan(sigma_i).any():
        raise ValueError("Argument 'sigma_i' contains NAs.")
    if y.shape[1] != x.shape[1]:
        raise ValueError("Arguments 'x' and 'y' contain different amounts of observations.")
    m = y.shape[0] * x.shape[0]
    a_prior = np.asarray(a_prior, dtype=np.float64).reshape(-1)
    if a_prior.shape[0] != m:
        raise ValueError("Argument 'a_prior' does not contain the required amount of elements.")
    if v_i_prior.shape[0] != m:
        raise ValueError("Argument 'v_i_prior' does not contain the required amount of elements.")
    v_post = sla.inv(v_i_prior + np.kron(x @ x.T, sigma_i))                       # :86
    a_post = v_post @ (v_i_prior @ a_prior + (sigma_i @ y @ x.T).reshape(-1, order="F"))   # :87
    s, U = sla.eigh(v_post)                                                        # :91
    return a_post + (U @ np.diag(np.sqrt(s)) @ U.T) @ eps                          # :93


def post_normal_sur(y, z, sigma_i, a_prior, v_i_prior, eps, svd=False):
    """src/post_normal_sur.cpp:60-113.  y k x T, z kT x n, sigma_i k x k or kT x k (time varying)."""
    if np.isnan(sigma_i).any():
        raise ValueError("Argument 'sigma_i' contains NAs.")
    k, t = y.shape
    nvars = z.shape[1]
    a_prior = np.asarray(a_prior, dtype=np.float64).reshape(-1)
    if a_prior.shape[0] != nvars:
        raise ValueError("Argument 'a_prior' does not contain the required amount of elements.")
    if v_i_prior.shape[0] != nvars:
        raise ValueError("Argument 'v_i_prior' does not contain the required amount of elements.")
    if sigma_i.shape[0] > k:
        S_i = np.zeros((k * t, k * t))
        for i in range(t):
            S_i[i * k:(i + 1) * k, i * k:(i + 1) * k] = sigma_i[i * k:(i + 1) * k]
        ZHi = z.T @ S_i                                                            # :92-96
    else:
        ZHi = z.T @ np.kron(np.eye(t), sigma_i)                                    # :90
    ZHZ = ZHi @ z
    ZHy = ZHi @ y.reshape(-1, order="F")
    V_post = sla.inv(v_i_prior + ZHZ)                                              # :101
    mu_post = V_post @ (v_i_prior @ a_prior + ZHy)
    if svd:
        U, s, Vt = sla.svd(V_post)
        root = U @ np.diag(np.sqrt(s)) @ Vt
    else:
        s, U = sla.eigh(V_post)
        root = U @ np.diag(np.sqrt(s)) @ U.T
    return mu_post + root @ eps


def ssvs(a, tau0, tau1, prob_prior, u, include=None):
    """src/svss.cpp:76-111.  u: n uniforms indexed by coefficient id; include: 1-based ids or None (= all)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    tau0 = np.asarray(tau0, dtype=np.float64).reshape(-1)
    tau1 = np.asarray(tau1, dtype=np.float64).reshape(-1)
    prob_prior = np.asarray(prob_prior, dtype=np.float64).reshape(-1)
    m = a.shape[0]
    tau0_sq, tau1_sq = tau0 ** 2, tau1 ** 2
    u0 = 1.0 / tau0 * np.exp(-(a ** 2 / (2.0 * tau0_sq))) * (1.0 - prob_prior)     # :81
    u1 = 1.0 / tau1 * np.exp(-(a ** 2 / (2.0 * tau1_sq))) * prob_prior             # :82
    p_post = u1 / (u0 + u1)
    lam = np.ones(m)
    V = np.zeros((m, m))
    V[np.arange(m), np.arange(m)] = 1.0 / tau1_sq
    pos = np.arange(m) if include is None else np.asarray(include, dtype=np.int64).reshape(-1) - 1
    for kk in pos:
        temp = float(bernoulli_rbinom(p_post[kk], u[kk]))
        lam[kk] = temp
        if temp == 0:
            V[kk, kk] = 1.0 / tau0_sq[kk]
    return {"v_i": V, "lambda": lam, "p_post": p_post}


def bvs(y, z, a, lam, sigma_i, prob_prior, u1, u2, include=None):
    """src/bvs.cpp:73-159.  y k x T, z kT x n, a n x 1 (or n x T for TVP), lam n x n diagonal indicator matrix,
    sigma_i k x k or kT x k.  u1/u2: n uniforms by coefficient id (gate / acceptance).  Returns lambda (n x n)."""
    prob_prior = np.asarray(prob_prior, dtype=np.float64).reshape(-1)
    lpr0 = np.log(1.0 - prob_prior)
    lpr1 = np.log(prob_prior)
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    lam = np.array(lam, dtype=np.float64)
    k, t = y.shape
    m = a.shape[0]
    const_par = a.shape[1] == 1
    if sigma_i.shape[0] > k:
        S_i = np.zeros((k * t, k * t))
        for i in range(t):
            S_i[i * k:(i + 1) * k, i * k:(i + 1) * k] = sigma_i[i * k:(i + 1) * k]
    else:
        S_i = np.kron(np.eye(t), sigma_i)
    yvec = y.reshape(-1, order="F")
    AG = lam @ a
    pos = np.arange(m) if include is None else np.asarray(include, dtype=np.int64).reshape(-1) - 1
    for var in pos:
        g = np.log(u1[var])
        if lam[var, var] == 1 and g >= lpr1[var]:
            continue
        if lam[var, var] == 0 and g >= lpr0[var]:
            continue
        theta0 = AG.copy()
        theta1 = AG.copy()
        theta1[var] = a[var]
        theta0[var] = 0.0
        if const_par:
            l0_res = yvec - z @ theta0[:, 0]
            l1_res = yvec - z @ theta1[:, 0]
        else:
            l0_res = np.zeros(k * t)
            l1_res = np.zeros(k * t)
            for i in range(t):
                l0_res[i * k:(i + 1) * k] = yvec[i * k:(i + 1) * k] - z[i * k:(i + 1) * k] @ theta0[:, i]
                l1_res[i * k:(i + 1) * k] = yvec[i * k:(i + 1) * k] - z[i * k:(i + 1) * k] @ theta1[:, i]
        l0 = -0.5 * float(l0_res @ S_i @ l0_res) + lpr0[var]
        l1 = -0.5 * float(l1_res @ S_i @ l1_res) + lpr1[var]
        lam[var, var] = 1.0 if (l1 - l0) >= np.log(u2[var]) else 0.0
    return lam


def _sym_sqrt(m):
    s, U = sla.eigh(np.asarray(m, dtype=np.float64))
    return U @ np.diag(np.sqrt(s)) @ U.T


def kalman_dk(y, z, sigma_u, sigma_v, B, a_init, P_init, eps0, eps_y, eps_a):
    """src/kalman_dk.cpp:91-184 (Durbin-Koopman 2002 Alg. 2 with the Jarocinski 2015 correction).

    y k x T, z kT x n, sigma_u k x k or kT x k, sigma_v n x n or nT x n, B n x n or nT x n, a_init n, P_init n x n.
    Injected normals in the reference's consumption order: eps0 (n), then per t eps_y[t] (k), eps_a[t] (n).
    Returns n x (T+1).
    """
    y = np.asarray(y, dtype=np.float64)
    k, t = y.shape
    nvars = z.shape[1]
    if sigma_u.shape[0] == k:
        sigma_u = np.tile(sigma_u, (t, 1))
    if sigma_v.shape[0] == nvars:
        sigma_v = np.tile(sigma_v, (t, 1))
    if B.shape[0] == nvars:
        B = np.tile(B, (t, 1))
    a_init = np.asarray(a_init, dtype=np.float64).reshape(-1)
    yplus = np.zeros_like(y)
    aplus = np.zeros((nvars, t + 1))
    aplus[:, 0] = _sym_sqrt(P_init) @ eps0                                          # :130-132
    for i in range(t):
        ku, kv = slice(i * k, (i + 1) * k), slice(i * nvars, (i + 1) * nvars)
        yplus[:, i] = z[ku] @ aplus[:, i] + _sym_sqrt(sigma_u[ku]) @ eps_y[i]       # :141-143
        aplus[:, i + 1] = B[kv] @ aplus[:, i] + _sym_sqrt(sigma_v[kv]) @ eps_a[i]   # :145-147
    ystar = y - yplus
    a = np.zeros((nvars, t + 1))
    a[:, 0] = a_init
    P = np.array(P_init, dtype=np.float64)
    v = np.zeros_like(y)
    Fi = np.zeros((k * t, k))
    K = np.zeros((nvars * t, k))
    L = np.zeros((nvars * t, nvars))
    for i in range(t):
        ku, kv = slice(i * k, (i + 1) * k), slice(i * nvars, (i + 1) * nvars)
        v[:, i] = ystar[:, i] - z[ku] @ a[:, i]
        Fi[ku] = sla.inv(z[ku] @ P @ z[ku].T + sigma_u[ku])
        K[kv] = B[kv] @ P @ z[ku].T @ Fi[ku]
        L[kv] = B[kv] - K[kv] @ z[ku]
        a[:, i + 1] = B[kv] @ a[:, i] + K[kv] @ v[:, i]
        P = B[kv] @ P @ L[kv].T + sigma_v[kv]
    r = np.zeros((nvars, t))
    for i in range(t - 1, 0, -1):
        ku, kv = slice(i * k, (i + 1) * k), slice(i * nvars, (i + 1) * nvars)
        r[:, i - 1] = z[ku].T @ Fi[ku] @ v[:, i] + L[kv].T @ r[:, i]
    r0 = z[0:k].T @ Fi[0:k] @ v[:, 0] + L[0:nvars].T @ r[:, 0]
    a[:, 0] = a_init + P_init @ r0
    for i in range(t):
        kv = slice(i * nvars, (i + 1) * nvars)
        a[:, i + 1] = B[kv] @ a[:, i] + sigma_v[kv] @ r[:, i]
    return a + aplus


def loglik_normal(u, sigma):
    """src/loglik_normal.cpp:37-58, literally: part_a = -k log(2 pi) / 2, part_b = -log(det(Sigma)) / 2,
    part_c = -u_t' solve(Sigma, I) u_t / 2; sigma with more rows than u is the (k T) x k stack of per-period blocks."""
    u = np.asarray(u, dtype=float)
    sigma = np.asarray(sigma, dtype=float)
    k, t = u.shape
    dmat = np.eye(k)
    result = np.zeros(t)
    part_a = -k * np.log(2 * np.pi) / 2
    if sigma.shape[0] > k:
        for i in range(t):
            blk = sigma[i * k:(i + 1) * k, :]
            part_b = -np.log(np.linalg.det(blk)) / 2
            part_c = -float(u[:, i] @ np.linalg.solve(blk, dmat) @ u[:, i]) / 2
            result[i] = part_a + part_b + part_c
    else:
        part_b = -np.log(np.linalg.det(sigma)) / 2
        si = np.linalg.solve(sigma, dmat)
        for i in range(t):
            part_c = -float(u[:, i] @ si @ u[:, i]) / 2
            result[i] = part_a + part_b + part_c
    return result


def summary_bvarlist_row(y, x, temp_pars, sigma_draws, tvp=False, sv=False):
    """The per-model body of R/summary.bvarlist.R:143-185 for a `bvar` object: y (T x K), x (T x M), temp_pars =
    cbind(A, B, C) draws (draws x n; TVP: list over periods), sigma_draws (draws x K^2; SV: list over periods).
    Returns dict(LL, AIC, BIC, HQ, ll_t = rowMeans(LL))."""
    y = np.asarray(y, dtype=float)
    tt, k = y.shape
    xm = np.asarray(x, dtype=float).T                      # :60 x <- t(object$x)
    tot_pars = xm.shape[0]
    draws = temp_pars[0].shape[0] if tvp else temp_pars.shape[0]
    LL = np.full((tt, draws), np.nan)
    yt = y.T
    for j in range(draws):
        if tvp:
            u = np.empty_like(yt)
            for period in range(tt):
                u[:, period] = yt[:, period] - temp_pars[period][j].reshape(k, -1, order="F") @ xm[:, period]
        else:
            u = yt - temp_pars[j].reshape(k, -1, order="F") @ xm
        if sv:
            sigma = np.vstack([sigma_draws[period][j].reshape(k, k, order="F") for period in range(tt)])
        else:
            sigma = sigma_draws[j].reshape(k, k, order="F")
        LL[:, j] = loglik_normal(u, sigma)
    ll = float(np.sum(LL.mean(axis=1)))
    return {"LL": ll, "AIC": 2 * tot_pars - 2 * ll, "BIC": np.log(tt) * tot_pars - 2 * ll,
            "HQ": 2 * np.log(np.log(tt)) * tot_pars - 2 * ll, "ll_t": LL.mean(axis=1)}


def spectrum0_ar(x):
    """coda::spectrum0.ar for ONE series (the time-series SE of coda::summary.mcmc, which summary.bvar reports:
    R/summary.bvar.R:133,141,195,203).  PARITY UNPINNED: coda (CRAN, un-vendored; DESCRIPTION Imports: coda) and
    stats::ar.yw are absent here; this restates their published algorithm --
      * if the residuals of lm(x ~ 1:n) have sd < 1.5e-8 (all.equal to 0) the spectrum is 0;
      * otherwise ar(x, aic = TRUE) by Yule-Walker: order.max = min(n - 1, floor(10 log10 n)), autocovariances with
        divisor n around the sample mean, Levinson-Durbin recursion, AIC_m = n log(v_m) + 2 m + 2, first minimiser,
        var.pred = v_m n / (n - (m + 1)), and spec = var.pred / (1 - sum(ar))^2.
    Returns (spec, order)."""
    x = np.asarray(x, dtype=float).reshape(-1)
    n = x.size
    z = np.arange(1, n + 1, dtype=float)
    zc = z - z.mean()
    xc = x - x.mean()
    resid = xc - zc * (zc @ xc) / (zc @ zc)
    if n < 2 or np.sqrt(resid @ resid / (n - 1)) < 1.5e-8:
        return 0.0, 0
    L = min(n - 1, int(np.floor(10.0 * np.log10(n))))
    r = np.array([xc[:n - l] @ xc[l:] / n for l in range(L + 1)])
    v = np.empty(L + 1)
    v[0] = r[0]
    coefs = [np.zeros(0)]
    phi = np.zeros(0)
    for m in range(1, L + 1):
        kappa = (r[m] - phi @ r[m - 1:0:-1]) / v[m - 1] if m > 1 else r[1] / v[0]
        phi = np.concatenate([phi - kappa * phi[::-1], [kappa]])
        v[m] = v[m - 1] * (1.0 - kappa * kappa)
        coefs.append(phi.copy())
    with np.errstate(divide="ignore", invalid="ignore"):
        aic = n * np.log(v) + 2.0 * np.arange(L + 1) + 2.0
    order = int(np.nanargmin(aic))
    var_pred = v[order] * n / (n - (order + 1))
    return float(var_pred / (1.0 - coefs[order].sum()) ** 2), order


def time_series_se(draws, n_chains):
    """coda::summary.mcmc.list: draws [n_chains * iters][rows] (chain-major) -> sqrt(mean_c spec0_c / (iters n_chains));
    one chain gives summary.mcmc's sqrt(spec0 / iters)."""
    d = np.asarray(draws, dtype=float)
    rows = d.shape[1]
    it = d.shape[0] // n_chains
    d = d.reshape(n_chains, it, rows)
    spec = np.array([[spectrum0_ar(d[c, :, r])[0] for r in range(rows)] for c in range(n_chains)])
    return np.sqrt(spec.mean(axis=0) / (it * n_chains))


def ir(A, Sigma, h, type, impulse, response, shock=1.0, A0=None):
    """src/ir.cpp:4-72, literally (full Phi matrices): A is k x (k p), impulse / response 1-based as in R."""
    coef = np.asarray(A, dtype=float)
    k = coef.shape[0]
    p = coef.shape[1] // k
    p_cols = h * k if h < p else coef.shape[1]
    P = np.eye(k)
    if type == "feir":
        P = P * shock
    elif type == "oir":
        P = np.linalg.cholesky(np.asarray(Sigma, dtype=float))          # trans(chol(Sigma)) = lower factor
        P = P @ np.diag(1.0 / np.diag(P)) * shock
    elif type == "gir":
        S = np.asarray(Sigma, dtype=float)
        P = S / S[impulse - 1, impulse - 1] * shock
    elif type == "sir":                                                  # :35-37
        P = np.linalg.solve(np.asarray(A0, dtype=float), np.eye(k)) * shock
    elif type == "sgir":                                                 # :42-46
        S = np.asarray(Sigma, dtype=float)
        P = np.linalg.solve(np.asarray(A0, dtype=float), np.eye(k))
        P = P @ S / S[impulse - 1, impulse - 1] * shock
    else:
        raise ValueError("type")
    phi = np.zeros(((h + 1) * k, k))
    phi[:k] = np.eye(k)
    theta = np.zeros(h + 1)
    theta[0] = (np.eye(k) @ P)[response - 1, impulse - 1]
    A_temp = np.zeros((k, h * k))
    A_temp[:, :p_cols] = coef[:, :p_cols]
    for i in range(1, h + 1):
        phi_temp = np.zeros((k, k))
        for j in range(1, i + 1):
            phi_temp = phi_temp + phi[(i - j) * k:(i - j + 1) * k] @ A_temp[:, (j - 1) * k:j * k]
        phi[i * k:(i + 1) * k] = phi_temp
        theta[i] = (phi_temp @ P)[response - 1, impulse - 1]
    return theta


def irf_bvar(A_draws, Sigma_draws, k, impulse, response, n_ahead=5, ci=0.95, shock=1, type="feir", cumulative=False,
             keep_draws=False, A0_draws=None):
    """R/irf.bvar.R:140-216 for a constant-parameter `bvar`: A_draws (draws x k^2 p), Sigma_draws (draws x k^2),
    impulse / response 1-based.  Returns the (n_ahead + 1) x 3 band matrix or the draws."""
    res = []
    for i in range(A_draws.shape[0]):
        A = A_draws[i].reshape(k, -1, order="F") if A_draws.shape[1] else np.zeros((k, k))
        S = Sigma_draws[i].reshape(k, k, order="F")
        if isinstance(shock, str):
            sh = np.linalg.cholesky(S).T[impulse - 1, impulse - 1] if type == "oir" else np.sqrt(S[impulse - 1, impulse - 1])
            if shock == "nsd":
                sh = -sh
        else:
            sh = shock
        A0 = None if A0_draws is None else A0_draws[i].reshape(k, k, order="F")
        res.append(ir(A, S, n_ahead, type, impulse, response, sh, A0))
    result = np.array(res)
    if cumulative:
        result = np.cumsum(result, axis=1)
    if keep_draws:
        return result
    lo = (1 - ci) / 2
    return np.quantile(result, [lo, 0.5, 1 - lo], axis=0).T


def vardecomp(A, Sigma, h, type, response, A0=None):
    """src/vardecomp.cpp:4-87, literally; response 1-based.  Returns the (h + 1) x k table of one draw."""
    a = np.asarray(A, dtype=float)
    sigma = np.asarray(Sigma, dtype=float)
    k = a.shape[0]
    p = a.shape[1] // k
    p_cols = h * k if h < p else a.shape[1]
    if type == "oir":
        P = np.linalg.cholesky(sigma)
        sigma_mse = sigma
    elif type == "gir":
        P = sigma
        sigma_mse = sigma
    elif type == "sir":                                                  # :32-36
        a0i = np.linalg.solve(np.asarray(A0, dtype=float), np.eye(k))
        P = a0i
        sigma_mse = a0i @ a0i.T
    elif type == "sgir":                                                 # :41-45
        a0i = np.linalg.solve(np.asarray(A0, dtype=float), np.eye(k))
        P = a0i @ sigma
        sigma_mse = a0i @ sigma @ a0i.T
    else:
        raise ValueError("type")
    sigmajj = np.sqrt(sigma[response - 1, response - 1]) if type in ("gir", "sgir") else 1.0
    A_temp = np.zeros((k, h * k))
    A_temp[:, :p_cols] = a[:, :p_cols]
    result = np.zeros((h + 1, k))
    numerator = np.zeros((h + 1, k))
    mse = np.zeros(h + 1)
    ejt = np.zeros((1, k))
    ejt[0, response - 1] = 1
    phi = np.zeros(((h + 1) * k, k))
    phi_temp = np.eye(k)
    phi[:k] = phi_temp
    numerator[0] = np.square(ejt @ phi_temp @ P)
    mse[0] = (ejt @ phi_temp @ sigma_mse @ phi_temp.T @ ejt.T).item()
    result[0] = numerator[0] / sigmajj / mse[0]
    for i in range(1, h + 1):
        phi_temp = np.zeros((k, k))
        for j in range(1, i + 1):
            phi_temp = phi_temp + phi[(i - j) * k:(i - j + 1) * k] @ A_temp[:, (j - 1) * k:j * k]
        phi[i * k:(i + 1) * k] = phi_temp
        numerator[i] = numerator[i - 1] + np.square(ejt @ phi_temp @ P)
        mse[i] = mse[i - 1] + (ejt @ phi_temp @ sigma_mse @ phi_temp.T @ ejt.T).item()
        result[i] = numerator[i] / sigmajj / mse[i]
    return result


def fevd_bvar(A_draws, Sigma_draws, k, response, n_ahead=5, type="oir", normalise_gir=False, A0_draws=None):
    """R/fevd.bvar.R:128-175 for a constant-parameter `bvar`: the mean over the draws of .vardecomp."""
    acc = np.zeros((n_ahead + 1, k))
    for i in range(A_draws.shape[0]):
        A = A_draws[i].reshape(k, -1, order="F") if A_draws.shape[1] else np.zeros((k, k))
        A0 = None if A0_draws is None else A0_draws[i].reshape(k, k, order="F")
        acc += vardecomp(A, Sigma_draws[i].reshape(k, k, order="F"), n_ahead, type, response, A0)
    result = acc / A_draws.shape[0]
    if type in ("gir", "sgir") and normalise_gir:
        result = result / result.sum(axis=1, keepdims=True)
    return result


def draw_forecast(a, sigma, pred, k, p, eps, a0_i=None):
    """src/draw_forecast.cpp:6-51 for one draw, literally: a = k x n_tot coefficient matrix or None, sigma k x k,
    pred tot x (n_ahead + 1), eps (n_ahead x k) the injected arma::randn(k) of each period, a0_i = solve(A0) for
    structural models (R/predict.bvar.R:200-209), identity otherwise."""
    pred = np.array(pred, dtype=float)
    a0_i = np.eye(k) if a0_i is None else np.asarray(a0_i, dtype=float)
    n_ahead = pred.shape[1] - 1
    use_a = a is not None
    n_tot = a.shape[1] if use_a else 0
    for j in range(n_ahead):
        eigval, eigvec = np.linalg.eigh(np.asarray(sigma, dtype=float))
        u = eigvec @ np.diag(np.sqrt(eigval)) @ eigvec.T @ eps[j]
        if use_a:
            if p == 0:
                pred[:k, j + 1] = a0_i @ a @ pred[k:k + n_tot, j] + a0_i @ u
            else:
                pred[:k, j + 1] = a0_i @ a @ pred[:, j] + a0_i @ u
        else:
            pred[:k, j + 1] = pred[:, j] + a0_i @ u
        if p > 1:
            for l in range(p - 1):
                pred[(l + 1) * k:(l + 2) * k, j + 1] = pred[l * k:(l + 1) * k, j]
    return pred

