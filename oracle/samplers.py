"""Oracle restatement of the built-in Gibbs samplers -- TEST INFRASTRUCTURE ONLY.

`bvaralg`    follows src/bvaralg.cpp:9-627   (constant-parameter BVAR; Wishart / gamma / SV errors, SSVS, BVS)
`bvartvpalg` follows src/bvartvpalg.cpp:9-626 (TVP BVAR, Chan-Jeliazkov precision sampler; SV / Wishart / gamma, BVS)

`literal=True` materialises the same dense matrices the reference builds (kron(I_T, Sigma^-1), the T x T SV system,
the (nT) x (nT) TVP system) and runs the same LAPACK-level operations; `literal=False` evaluates the identical
formulas through their Kronecker / banded structure (used for long tier-2 runs; proven equal to the literal
path to <= 1e-10 in tests/test_oracle.py).  The covariance (psi) block of the constant-parameter sampler
(bvaralg.cpp:373-455, gamma / SV errors) with its SSVS / BVS twins, the TVP psi block (bvartvpalg.cpp:339-407) and
structural models (general z from data$SUR) are restated literally.  Random variates come from a VariateSource (injected arrays
or a NumPy Generator) in the site order of oracle/variates.py.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .blocks import stochvol_mixture
from .variates import VariateSource, bernoulli_rbinom


def wishrnd(S, chi2, normals):
    """arma::wishrnd(S, df) with injected variates (convention documented in oracle/variates.py)."""
    k = S.shape[0]
    D = sla.cholesky(S, lower=False)
    A = np.zeros((k, k))
    A[np.arange(k), np.arange(k)] = np.sqrt(chi2)
    c = 0
    for i in range(1, k):
        for j in range(i):
            A[j, i] = normals[c]
            c += 1
    AD = A @ D
    return AD.T @ AD


def _unpack_common(obj):
    data, model, priors, initial = obj["data"], obj["model"], obj["priors"], obj["initial"]
    y = np.asarray(data["Y"], dtype=np.float64).T.copy()          # k x T
    k, tt = y.shape
    z = None if data.get("SUR") is None else np.asarray(data["SUR"], dtype=np.float64)
    n_tot = 0 if z is None else z.shape[1]
    p = int(model["endogen"]["lags"])
    n_a = k * k * p
    n_b = 0
    if "exogen" in model:
        n_b = k * len(model["exogen"]["variables"]) * (int(model["exogen"]["lags"]) + 1)
    n_c = k * len(model.get("deterministic", []))
    return data, model, priors, initial, y, k, tt, z, n_tot, n_a, n_b, n_c


def _split_rows(posteriors, draws, lam, n_a, n_b, n_c, tvp_T=None, sig=None):
    """Row slicing of the stored coefficient draws into a / b / c (bvaralg.cpp:535-552, bvartvpalg.cpp:521-543)."""
    off = 0
    for name, cnt in (("a", n_a), ("b", n_b), ("c", n_c)):
        if cnt > 0:
            d = {}
            if tvp_T is None:
                d["coeffs"] = draws[off:off + cnt]
            else:
                ntot = draws.shape[0] // tvp_T
                cube = draws.reshape(tvp_T, ntot, -1)
                d["coeffs"] = cube[:, off:off + cnt, :].reshape(tvp_T * cnt, -1)
                d["sigma"] = sig[off:off + cnt]
            if lam is not None:
                d["lambda"] = lam[off:off + cnt]
            posteriors[name] = d
        off += cnt


def bvaralg(obj, variates: VariateSource | None = None, literal=True, table="ksc1998"):
    data, model, priors, initial, y, k, tt, z, n_tot, n_a, n_b, n_c = _unpack_common(obj)
    vs = variates or VariateSource()
    yvec = y.reshape(-1, order="F")
    use_a = n_tot > 0
    sv = bool(model["sv"])
    diag_k = np.eye(k)
    x = None if data.get("Z") is None else np.asarray(data["Z"], dtype=np.float64).T   # M x T (structured path)

    ssvs = bvs = False
    if use_a:
        pc = priors["coefficients"]
        prior_mu = np.asarray(pc["mu"], dtype=np.float64).reshape(-1)
        prior_vi = np.array(pc["v_i"], dtype=np.float64)
        if "ssvs" in pc:
            ssvs = True
            sel = pc["ssvs"]
            a_prior_incl = np.asarray(sel["inprior"], dtype=np.float64).reshape(-1)
            tau0 = np.asarray(sel["tau0"], dtype=np.float64).reshape(-1)
            tau1 = np.asarray(sel["tau1"], dtype=np.float64).reshape(-1)
            tau0sq, tau1sq = tau0 ** 2, tau1 ** 2
            if sv:
                raise ValueError("Not allowed to use SSVS with stochastic volatility.")      # :116-118
        if "bvs" in pc:
            bvs = True
            sel = pc["bvs"]
            lprior0 = np.log(1.0 - np.asarray(sel["inprior"], dtype=np.float64).reshape(-1))
            lprior1 = np.log(np.asarray(sel["inprior"], dtype=np.float64).reshape(-1))
        varsel = ssvs or bvs
        a = np.asarray(initial["coefficients"]["draw"], dtype=np.float64).reshape(-1).copy()
        if varsel:
            include = np.asarray(sel["include"], dtype=np.int64).reshape(-1) - 1             # :132
            lam = np.ones(n_tot)
            z_bvs = z
    else:
        varsel = False

    sp = priors["sigma"]
    use_gamma = False
    if sv:
        s_prior_mu = np.asarray(sp["mu"], dtype=np.float64).reshape(-1)
        s_prior_vi = np.asarray(sp["v_i"], dtype=np.float64)
        s_post_shape = np.asarray(sp["shape"], dtype=np.float64).reshape(-1) + 0.5 * tt
        s_prior_rate = np.asarray(sp["rate"], dtype=np.float64).reshape(-1)
        h = np.array(initial["sigma"]["h"], dtype=np.float64)          # T x k
        sigma_h = np.asarray(initial["sigma"]["sigma_h"], dtype=np.float64).reshape(-1).copy()
        h_const = np.asarray(initial["sigma"]["constant"], dtype=np.float64).reshape(-1)
        h_init = h[0].copy()
        sigma_i = np.diag(1.0 / np.exp(h_init))
    else:
        if "df" in sp:
            post_df = float(sp["df"]) + tt
            prior_scale = np.asarray(sp["scale"], dtype=np.float64)
        if "shape" in sp:
            use_gamma = True
            s_post_shape = np.asarray(sp["shape"], dtype=np.float64).reshape(-1) + 0.5 * tt
            s_prior_rate = np.asarray(sp["rate"], dtype=np.float64).reshape(-1)
        omega_i = np.array(initial["sigma"]["sigma_i"], dtype=np.float64)
        sigma_i = omega_i.copy()
    # per-period precision blocks: sig_t[t] is k x k.  Initially only the diagonal of sigma_i is used (:246).
    sig_t = np.tile(np.diag(np.diag(sigma_i))[None], (tt, 1, 1))
    # covariance (psi) block, bvaralg.cpp:157-204,247-249,373-455: priors$psi = {mu, v_i}
    covar = "psi" in priors
    if covar:
        if not (sv or use_gamma):
            raise ValueError("The covariance block needs a gamma or stochastic-volatility error prior.")
        psi_prior_mu = np.asarray(priors["psi"]["mu"], dtype=np.float64).reshape(-1)
        psi_prior_vi = np.asarray(priors["psi"]["v_i"], dtype=np.float64)
        n_psi = k * (k - 1) // 2
        omega_diag = np.tile(np.diag(sigma_i), tt)                    # diag_omega_i = diag_sigma_i at the start, :247-249
        Psi = np.eye(k)
        psi_ssvs = "ssvs" in priors["psi"]                            # :164-172
        psi_bvs = "bvs" in priors["psi"]                              # :177-183
        if sv and psi_ssvs:
            raise ValueError("Not allowed to use SSVS with stochastic volatility.")          # :173-175
        if psi_ssvs:
            pv = priors["psi"]["ssvs"]
            psi_incl_prior = np.asarray(pv["inprior"], dtype=np.float64).reshape(-1)
            psi_tau0 = np.asarray(pv["tau0"], dtype=np.float64).reshape(-1)
            psi_tau1 = np.asarray(pv["tau1"], dtype=np.float64).reshape(-1)
        if psi_bvs:
            pv = priors["psi"]["bvs"]
            psi_lp0 = np.log(1.0 - np.asarray(pv["inprior"], dtype=np.float64).reshape(-1))
            psi_lp1 = np.log(np.asarray(pv["inprior"], dtype=np.float64).reshape(-1))
        psi_varsel = psi_ssvs or psi_bvs
        if psi_varsel:
            psi_include = np.asarray(pv["include"], dtype=np.int64).reshape(-1) - 1          # :190
            psi_lam = np.ones(n_psi)
            psi_prior_vi = psi_prior_vi.copy()
    draws_psi_lambda = None

    iters, burnin = int(model["iterations"]), int(model["burnin"])
    draws_coef = np.zeros((n_tot, iters))
    draws_sigma = np.zeros((k * k * tt if sv else k * k, iters))
    draws_sigma_sigma = np.zeros((k * k, iters)) if sv else None
    draws_lambda = np.zeros((n_tot, iters)) if varsel else None

    def dense_D():
        return sla.block_diag(*sig_t)

    for draw in range(iters + burnin):
        if use_a:
            lam_use = lam if bvs else None
            if literal:
                zz = z_bvs * lam[None, :] if bvs else z                                     # :303 (Lambda diagonal)
                D = dense_D()
                zD = zz.T @ D
                post_v = prior_vi + zD @ zz                                                  # :305
                rhs = prior_vi @ prior_mu + zD @ yvec
            else:
                # z = kron(x_t', I_k):  z'Dz = sum_t (x_t x_t') (x) Sigma_t^-1,  z'Dy = vec(sum_t Sigma_t^-1 y_t x_t')
                G = np.einsum("jt,lt,tab->jalb", x, x, sig_t).reshape(n_tot, n_tot)
                # index (j,a) -> j*k + a  (coefficient id), matches kron(x, I_k) column order
                gy = np.einsum("tab,bt,jt->ja", sig_t, y, x).reshape(n_tot)
                if bvs:
                    G = G * lam[:, None] * lam[None, :]
                    gy = gy * lam
                post_v = prior_vi + G
                rhs = prior_vi @ prior_mu + gy
            post_mu = sla.solve(post_v, rhs)                                                 # :306
            eps = vs.normal("coef_normal", draw, (n_tot,))
            a = post_mu + sla.solve_triangular(sla.cholesky(post_v, lower=False), eps, lower=False)   # :307

            if ssvs:
                u0 = 1.0 / tau0 * np.exp(-(a ** 2 / (2.0 * tau0sq))) * (1.0 - a_prior_incl)  # :316
                u1 = 1.0 / tau1 * np.exp(-(a ** 2 / (2.0 * tau1sq))) * a_prior_incl          # :317
                post_incl = u1 / (u0 + u1)
                uu = vs.uniform("ssvs_u", draw, (n_tot,))
                for i in include:                                                            # order-free
                    ld = float(bernoulli_rbinom(post_incl[i], uu[i]))
                    lam[i] = ld
                    prior_vi[i, i] = 1.0 / tau0sq[i] if ld == 0 else 1.0 / tau1sq[i]
            if bvs:
                a_AG = lam * a                                                               # :335 (stale)
                g1 = np.log(vs.uniform("bvs_u1", draw, (n_tot,)))
                g2 = np.log(vs.uniform("bvs_u2", draw, (n_tot,)))
                D = dense_D() if literal else None
                new_lam = lam.copy()
                for j in include:
                    if lam[j] == 1 and g1[j] >= lprior1[j]:
                        continue
                    if lam[j] == 0 and g1[j] >= lprior0[j]:
                        continue
                    th0 = a_AG.copy()
                    th1 = a_AG.copy()
                    th0[j] = 0.0
                    th1[j] = a[j]
                    r0 = yvec - z_bvs @ th0
                    r1 = yvec - z_bvs @ th1
                    if literal:
                        q0, q1 = float(r0 @ D @ r0), float(r1 @ D @ r1)
                    else:
                        R0, R1 = r0.reshape(k, tt, order="F"), r1.reshape(k, tt, order="F")
                        q0 = float(np.einsum("at,tab,bt->", R0, sig_t, R0))
                        q1 = float(np.einsum("at,tab,bt->", R1, sig_t, R1))
                    l0 = -q0 / 2.0 + lprior0[j]
                    l1 = -q1 / 2.0 + lprior1[j]
                    new_lam[j] = 1.0 if (l1 - l0) >= g2[j] else 0.0
                lam = new_lam
                a = lam * a                                                                  # :359
            u = (yvec - z @ a).reshape(k, tt, order="F")                                     # :364-365
        else:
            u = y

        if covar:                                                                            # :373-455
            psi_y = u[1:, :].reshape(-1, order="F")                                          # :376
            psi_z = np.zeros(((k - 1) * tt, n_psi))
            cov_omega = np.zeros((k - 1) * tt)
            for i in range(1, k):
                for j in range(tt):
                    psi_z[j * (k - 1) + i - 1, i * (i - 1) // 2:(i + 1) * i // 2] = -u[0:i, j]          # :379-382
                    cov_omega[j * (k - 1) + i - 1] = omega_diag[j * k + i]                                # :384
            psi_z_bvs = psi_z
            if psi_bvs:
                psi_z = psi_z_bvs * psi_lam[None, :]                                         # :388-391
            zo = psi_z.T * cov_omega[None, :]
            psi_post_v = psi_prior_vi + zo @ psi_z                                           # :392
            psi_post_mu = sla.solve(psi_post_v, psi_prior_vi @ psi_prior_mu + zo @ psi_y)    # :393
            psi = psi_post_mu + sla.solve_triangular(sla.cholesky(psi_post_v, lower=False),
                                                     vs.normal("psi_normal", draw, (n_psi,)), lower=False)   # :394
            if psi_ssvs:                                                                     # :401-418
                u0 = 1.0 / psi_tau0 * np.exp(-(psi ** 2 / (2.0 * psi_tau0 ** 2))) * (1.0 - psi_incl_prior)
                u1 = 1.0 / psi_tau1 * np.exp(-(psi ** 2 / (2.0 * psi_tau1 ** 2))) * psi_incl_prior
                post_incl = u1 / (u0 + u1)
                uu = vs.uniform("psi_u1", draw, (n_psi,))
                for i in psi_include:
                    ld = float(bernoulli_rbinom(post_incl[i], uu[i]))
                    psi_lam[i] = ld
                    psi_prior_vi[i, i] = 1.0 / psi_tau0[i] ** 2 if ld == 0 else 1.0 / psi_tau1[i] ** 2
            if psi_bvs:                                                                      # :420-448
                psi_AG = psi_lam * psi
                g1 = np.log(vs.uniform("psi_u1", draw, (n_psi,)))
                g2 = np.log(vs.uniform("psi_u2", draw, (n_psi,)))
                new_lam = psi_lam.copy()
                for j in psi_include:
                    if psi_lam[j] == 1 and g1[j] >= psi_lp1[j]:
                        continue
                    if psi_lam[j] == 0 and g1[j] >= psi_lp0[j]:
                        continue
                    th0 = psi_AG.copy()
                    th1 = psi_AG.copy()
                    th0[j] = 0.0
                    th1[j] = psi[j]
                    r0 = psi_y - psi_z_bvs @ th0
                    r1 = psi_y - psi_z_bvs @ th1
                    l0 = -float(r0 @ (cov_omega * r0)) / 2.0 + psi_lp0[j]
                    l1 = -float(r1 @ (cov_omega * r1)) / 2.0 + psi_lp1[j]
                    new_lam[j] = 1.0 if (l1 - l0) >= g2[j] else 0.0
                psi_lam = new_lam
                psi = psi_lam * psi                                                          # :446
            Psi = np.eye(k)
            for i in range(1, k):
                Psi[i, 0:i] = psi[i * (i - 1) // 2:(i + 1) * i // 2]                        # :451-453
            u = Psi @ u                                                                      # :454

        if sv:
            h = stochvol_mixture(u.T, h, sigma_h, h_init, h_const,
                                 vs.uniform("sv_u", draw, (k, tt)), vs.normal("sv_normal", draw, (k, tt)),
                                 table=table) if literal else \
                _stochvol_banded(u.T, h, sigma_h, h_init, h_const,
                                 vs.uniform("sv_u", draw, (k, tt)), vs.normal("sv_normal", draw, (k, tt)), table)
            h_lag = np.vstack([h_init[None, :], h[:-1]])
            dh = h - h_lag
            post_scale = 1.0 / (s_prior_rate + (dh ** 2).sum(axis=0) * 0.5)                  # :468
            g = vs.gamma("sigma_h_gamma", draw, s_post_shape)
            sigma_h = 1.0 / (post_scale * g)                                                 # :470
            sigma_h_i = np.diag(1.0 / sigma_h)
            hv = s_prior_vi + sigma_h_i
            hmu = sla.solve(hv, s_prior_vi @ s_prior_mu + sigma_h_i @ h[0])                  # :476
            h_init = hmu + sla.solve_triangular(sla.cholesky(hv, lower=False),
                                                vs.normal("h_init_normal", draw, (k,)), lower=False)
            sig_t = np.zeros((tt, k, k))
            sig_t[:, np.arange(k), np.arange(k)] = 1.0 / np.exp(h)                           # :494
            if covar:
                omega_diag = (1.0 / np.exp(h)).reshape(-1)                                   # vectorise(trans(h)), :494
                sig_t = np.einsum("ia,ti,ib->tab", Psi, 1.0 / np.exp(h), Psi)                # Psi' Omega_t^-1 Psi, :495-497
        else:
            if use_gamma:
                sse = u @ u.T
                g = vs.gamma("omega_gamma", draw, s_post_shape)
                for i in range(k):
                    omega_i[i, i] = g[i] * (1.0 / (s_prior_rate[i] + sse[i, i] * 0.5))      # :484
                sigma_i = omega_i.copy()                                                     # :507
                if covar:
                    omega_diag = np.tile(np.diag(omega_i), tt)                               # kron(diag_tt, omega_i), :504
                    sigma_i = Psi.T @ omega_i @ Psi                                          # :505
            else:
                Sbar = sla.solve(prior_scale + u @ u.T, diag_k)                              # :511
                sigma_i = wishrnd(Sbar, vs.chi2("wishart_chi2", draw, post_df - np.arange(k)),
                                  vs.normal("wishart_normal", draw, (k * (k - 1) // 2,)))
            sig_t = np.tile(sigma_i[None], (tt, 1, 1))                                       # :514

        if draw >= burnin:
            pos = draw - burnin
            if sv:
                for i in range(tt):
                    draws_sigma[i * k * k:(i + 1) * k * k, pos] = sla.solve(sig_t[i], diag_k).reshape(-1, order="F")
                draws_sigma_sigma[:, pos] = np.diag(sigma_h).reshape(-1, order="F")          # :526
            else:
                draws_sigma[:, pos] = sla.solve(sigma_i, diag_k).reshape(-1, order="F")      # :528
            if use_a:
                draws_coef[:, pos] = a
                if varsel:
                    draws_lambda[:, pos] = lam
            if covar and psi_varsel:
                if draws_psi_lambda is None:
                    draws_psi_lambda = np.zeros((n_psi, iters))
                draws_psi_lambda[:, pos] = psi_lam                                           # :530-532

    posteriors = {"a0": None, "a": None, "b": None, "c": None, "sigma": None}
    if use_a:
        _split_rows(posteriors, draws_coef, draws_lambda, n_a, n_b, n_c)
        if model.get("structural") and k > 1:                                                # :553-557,595-601
            a0_from = n_a + n_b + n_c
            posteriors["a0"] = {"coeffs": draws_coef[a0_from:]}
            if draws_lambda is not None:
                posteriors["a0"]["lambda"] = draws_lambda[a0_from:]
    posteriors["sigma"] = {"coeffs": draws_sigma}
    if sv:
        posteriors["sigma"]["sigma"] = draws_sigma_sigma
    if draws_psi_lambda is not None:
        posteriors["sigma"]["lambda"] = draws_psi_lambda                                     # :604-611
    return {"data": obj["data"], "model": obj["model"], "priors": obj["priors"], "posteriors": posteriors}


def _stochvol_banded(y, h, sigma, h_init, constant, u, eps, table):
    """Same formulas as blocks.stochvol_mixture with the T x T system solved in banded form (structured path)."""
    from .blocks import MIXTURES, _normpdf
    p_i, mu, sigma2 = MIXTURES[table]
    J = p_i.shape[0]
    y = np.array(y, dtype=np.float64)
    h = np.array(h, dtype=np.float64)
    tt, k = y.shape
    for i in range(k):
        ys = np.log(y[:, i] ** 2 + constant[i])
        q = p_i[None, :] * _normpdf(ys[:, None], h[:, i][:, None] + mu[None, :], np.sqrt(sigma2)[None, :])
        q = q / q.sum(axis=1, keepdims=True)
        s = np.minimum(J - (u[i][:, None] < np.cumsum(q, axis=1)).sum(axis=1), J - 1)
        dg = np.full(tt, 2.0 / sigma[i])
        dg[-1] = 1.0 / sigma[i]
        dg += 1.0 / sigma2[s]
        ab = np.zeros((2, tt))
        ab[0, 1:] = -1.0 / sigma[i]
        ab[1] = dg
        rhs = (ys - mu[s]) / sigma2[s]
        rhs[0] += h_init[i] / sigma[i]
        c = sla.cholesky_banded(ab, lower=False)
        mean = sla.cho_solve_banded((c, False), rhs)
        w = sla.solve_banded((0, 1), c, eps[i])          # R^-1 eps with R upper bidiagonal
        h[:, i] = mean + w
    return h


def bvartvpalg(obj, variates: VariateSource | None = None, literal=True, table="ksc1998"):
    data, model, priors, initial, y, k, tt, z, n_tot, n_a, n_b, n_c = _unpack_common(obj)
    vs = variates or VariateSource()
    yvec = y.reshape(-1, order="F")
    if n_tot == 0:
        raise NotImplementedError("TVP model without regressors")
    sv = bool(model["sv"])
    diag_k = np.eye(k)
    n = n_tot
    pc = priors["coefficients"]
    init_mu = np.asarray(pc["mu"], dtype=np.float64).reshape(-1)
    init_vi = np.asarray(pc["v_i"], dtype=np.float64)
    v_post_shape = np.asarray(pc["shape"], dtype=np.float64).reshape(-1) + 0.5 * tt
    v_prior_rate = np.asarray(pc["rate"], dtype=np.float64).reshape(-1)
    sigma_v_i = 1.0 / v_prior_rate                                    # :111-112 (diagonal)
    a_init = np.asarray(initial["coefficients"]["draw"], dtype=np.float64).reshape(-1).copy()
    bvs = "bvs" in pc
    if bvs:
        sel = pc["bvs"]
        lprior0 = np.log(1.0 - np.asarray(sel["inprior"], dtype=np.float64).reshape(-1))
        lprior1 = np.log(np.asarray(sel["inprior"], dtype=np.float64).reshape(-1))
        include = np.asarray(sel["include"], dtype=np.int64).reshape(-1) - 1
        lam = np.ones(n)

    # zz: block diagonal kT x nT (:95-98);  a_hh: first-difference matrix (:116-117)
    Zt = [z[i * k:(i + 1) * k] for i in range(tt)]
    Zarr = np.stack(Zt)

    sp = priors["sigma"]
    use_gamma = False
    if sv:
        s_prior_mu = np.asarray(sp["mu"], dtype=np.float64).reshape(-1)
        s_prior_vi = np.asarray(sp["v_i"], dtype=np.float64)
        s_post_shape = np.asarray(sp["shape"], dtype=np.float64).reshape(-1) + 0.5 * tt
        s_prior_rate = np.asarray(sp["rate"], dtype=np.float64).reshape(-1)
        h = np.array(initial["sigma"]["h"], dtype=np.float64)
        sigma_h = np.asarray(initial["sigma"]["sigma_h"], dtype=np.float64).reshape(-1).copy()
        h_const = np.asarray(initial["sigma"]["constant"], dtype=np.float64).reshape(-1)
        h_init = h[0].copy()
        sigma_u_i = np.diag(1.0 / np.exp(h_init))
    else:
        if "df" in sp:
            post_df = float(sp["df"]) + tt
            prior_scale = np.asarray(sp["scale"], dtype=np.float64)
        if "shape" in sp:
            use_gamma = True
            s_post_shape = np.asarray(sp["shape"], dtype=np.float64).reshape(-1) + 0.5 * tt
            s_prior_rate = np.asarray(sp["rate"], dtype=np.float64).reshape(-1)
        omega_i = np.array(initial["sigma"]["sigma_i"], dtype=np.float64)
        sigma_u_i = omega_i.copy()
    sig_t = np.tile(np.diag(np.diag(sigma_u_i))[None], (tt, 1, 1))    # :230

    # covariance (psi) block of the TVP sampler (:143-158, 165-186)
    covar = "psi" in priors
    if covar:
        pp = priors["psi"]
        n_psi = k * (k - 1) // 2
        psi_bvs = "bvs" in pp                                                                    # :152-157,177-185
        if psi_bvs:
            psel = pp["bvs"]
            psi_lprior0 = np.log(1.0 - np.asarray(psel["inprior"], dtype=np.float64).reshape(-1))
            psi_lprior1 = np.log(np.asarray(psel["inprior"], dtype=np.float64).reshape(-1))
            psi_include = np.asarray(psel["include"], dtype=np.int64).reshape(-1) - 1
            psi_lam = np.ones(n_psi)
        psi_prior_mu = np.asarray(pp["mu"], dtype=np.float64).reshape(-1)
        psi_prior_vi = np.asarray(pp["v_i"], dtype=np.float64)
        psi_post_shape = np.asarray(pp["shape"], dtype=np.float64).reshape(-1) + 0.5 * tt      # :149
        psi_prior_rate = np.asarray(pp["rate"], dtype=np.float64).reshape(-1)
        psi_init = np.asarray(initial["psi"]["draw"], dtype=np.float64).reshape(-1).copy()      # :169
        psi_sigma_v_i = 1.0 / psi_prior_rate                                                     # :175-176
        Hpsi = np.eye(n_psi * tt)
        ip = np.arange(n_psi * (tt - 1))
        Hpsi[ip + n_psi, ip] = -1.0                                                              # :172-173
        # diag_omega_i = diag_sigma_u_i at the start (:231-233); only the SV branch ever refreshes it (:446) -- with
        # the inverse-gamma prior the drawn omega_i (:436) never reaches it, so the measurement variances of a
        # gamma + covar TVP model stay at their initial values in the reference.  Restated as is.
        omega_diag = np.tile(np.diag(sigma_u_i), tt).astype(np.float64)

    iters, burnin = int(model["iterations"]), int(model["burnin"])
    draws_coef = np.zeros((n * tt, iters))
    draws_sigma_v = np.zeros((n, iters))
    per_t_sigma = sv or covar
    draws_sigma_u = np.zeros((k * k * tt if per_t_sigma else k * k, iters))
    draws_sigma_sigma = np.zeros((k * k, iters)) if sv else None
    draws_lambda = np.zeros((n, iters)) if bvs else None
    draws_psi_lambda = np.zeros((n_psi, iters)) if covar and psi_bvs else None

    if literal:
        H = np.eye(n * tt)
        idx = np.arange(n * (tt - 1))
        H[idx + n, idx] = -1.0
        zz_full = sla.block_diag(*Zt)

    for draw in range(iters + burnin):
        lam_d = lam if bvs else np.ones(n)
        if literal:
            zz = zz_full * np.tile(lam_d, tt)[None, :] if bvs else zz_full               # :291
            Dm = sla.block_diag(*sig_t)
            zzss = zz.T @ Dm                                                              # :293
            hsh = H.T @ np.kron(np.eye(tt), np.diag(sigma_v_i)) @ H                       # :294
            post_v = hsh + zzss @ zz                                                      # :295
            rhs = hsh @ np.tile(a_init, tt) + zzss @ yvec
            post_mu = sla.solve(post_v, rhs)                                              # :296
            eps = vs.normal("state_normal", draw, (n * tt,))
            a = post_mu + sla.solve_triangular(sla.cholesky(post_v, lower=False), eps, lower=False)
        else:
            a = _tvp_state_banded(Zt, sig_t, sigma_v_i, a_init, y, lam_d,
                                  vs.normal("state_normal", draw, (n * tt,)))
        a_mat = a.reshape(n, tt, order="F")
        if bvs:
            a_AG = lam[:, None] * a_mat                                                   # :303 (stale)
            g1 = np.log(vs.uniform("bvs_u1", draw, (n,)))
            g2 = np.log(vs.uniform("bvs_u2", draw, (n,)))
            new_lam = lam.copy()
            for j in include:
                if lam[j] == 1 and g1[j] >= lprior1[j]:
                    continue
                if lam[j] == 0 and g1[j] >= lprior0[j]:
                    continue
                th0 = a_AG.copy()
                th1 = a_AG.copy()
                th0[j] = 0.0
                th1[j] = a_mat[j]
                q = []
                for th in (th0, th1):
                    r = np.stack([y[:, t] - Zt[t] @ th[:, t] for t in range(tt)], axis=1)  # k x T
                    q.append(float(np.einsum("at,tab,bt->", r, sig_t, r)))
                l0 = -q[0] * 0.5 + lprior0[j]
                l1 = -q[1] * 0.5 + lprior1[j]
                new_lam[j] = 1.0 if (l1 - l0) >= g2[j] else 0.0
            lam = new_lam
            a_mat = lam[:, None] * a_mat                                                  # :328
            a = a_mat.reshape(-1, order="F")
        u = y - np.einsum("tkn,nt->kt", Zarr, a_mat)                                      # :332-336

        if covar:                                                                         # :339-407
            psi_y = u[1:, :].reshape(-1, order="F")
            psi_z = np.zeros(((k - 1) * tt, n_psi * tt))
            dco = np.zeros((k - 1) * tt)
            for i in range(1, k):
                for j in range(tt):
                    psi_z[j * (k - 1) + i - 1, j * n_psi + i * (i - 1) // 2: j * n_psi + (i + 1) * i // 2] = -u[:i, j]
                    dco[j * (k - 1) + i - 1] = omega_diag[j * k + i]                      # :350
            if psi_bvs:
                psi_z_bvs = psi_z
                psi_z = psi_z_bvs * np.tile(psi_lam, tt)[None, :]                         # :354-357
            pzz = psi_z.T * dco[None, :]                                                  # :359
            phsh = Hpsi.T @ np.kron(np.eye(tt), np.diag(psi_sigma_v_i)) @ Hpsi            # :360
            ppv = phsh + pzz @ psi_z
            ppm = sla.solve(ppv, phsh @ np.tile(psi_init, tt) + pzz @ psi_y)              # :362
            psi = ppm + sla.solve_triangular(sla.cholesky(ppv, lower=False),
                                             vs.normal("psi_normal", draw, (n_psi * tt,)), lower=False)   # :363
            if psi_bvs:                                                                   # :365-398
                psi_mat = psi.reshape(n_psi, tt, order="F")
                psi_AG = psi_lam[:, None] * psi_mat                                       # :372 (stale inside the loop)
                pg1 = np.log(vs.uniform("psi_u1", draw, (n_psi,)))
                pg2 = np.log(vs.uniform("psi_u2", draw, (n_psi,)))
                new_plam = psi_lam.copy()
                for j in psi_include:
                    if psi_lam[j] == 1 and pg1[j] >= psi_lprior1[j]:
                        continue
                    if psi_lam[j] == 0 and pg1[j] >= psi_lprior0[j]:
                        continue
                    th0 = psi_AG.copy()
                    th1 = psi_AG.copy()
                    th0[j] = 0.0
                    th1[j] = psi_mat[j]
                    r0 = psi_y - psi_z_bvs @ th0.reshape(-1, order="F")
                    r1 = psi_y - psi_z_bvs @ th1.reshape(-1, order="F")
                    l0 = -float(r0 @ (dco * r0)) / 2 + psi_lprior0[j]
                    l1 = -float(r1 @ (dco * r1)) / 2 + psi_lprior1[j]
                    new_plam[j] = 1.0 if (l1 - l0) >= pg2[j] else 0.0
                psi_lam = new_plam
                psi = (psi_lam[:, None] * psi_mat).reshape(-1, order="F")                 # :396
            Psi_t = np.tile(np.eye(k)[None], (tt, 1, 1))
            for j in range(tt):
                for i in range(1, k):
                    Psi_t[j, i, :i] = psi[j * n_psi + i * (i - 1) // 2: j * n_psi + (i + 1) * i // 2]      # :400-404
            u = np.einsum("tab,bt->at", Psi_t, u)                                         # :406

        if sv:
            args = (u.T, h, sigma_h, h_init, h_const, vs.uniform("sv_u", draw, (k, tt)),
                    vs.normal("sv_normal", draw, (k, tt)))
            h = stochvol_mixture(*args, table=table) if literal else _stochvol_banded(*args, table)
            h_lag = np.vstack([h_init[None, :], h[:-1]])
            dh = h - h_lag
            post_scale = 1.0 / (s_prior_rate + (dh ** 2).sum(axis=0) * 0.5)
            g = vs.gamma("sigma_h_gamma", draw, s_post_shape)
            sigma_h = 1.0 / (post_scale * g)                                              # :422
            sigma_h_i = np.diag(1.0 / sigma_h)
            hv = s_prior_vi + sigma_h_i
            hmu = sla.solve(hv, s_prior_vi @ s_prior_mu + sigma_h_i @ h[0])
            h_init = hmu + sla.solve_triangular(sla.cholesky(hv, lower=False),
                                                vs.normal("h_init_normal", draw, (k,)), lower=False)
            sig_t = np.zeros((tt, k, k))
            sig_t[:, np.arange(k), np.arange(k)] = 1.0 / np.exp(h)                        # :446
            if covar:
                omega_diag = (1.0 / np.exp(h)).reshape(-1)                                # period-major: vectorise(h.t())
                sig_t = np.einsum("tba,tb,tbc->tac", Psi_t, 1.0 / np.exp(h), Psi_t)       # :448
        else:
            if use_gamma:
                sse = u @ u.T
                g = vs.gamma("omega_gamma", draw, s_post_shape)
                for i in range(k):
                    omega_i[i, i] = g[i] * (1.0 / (s_prior_rate[i] + sse[i, i] * 0.5))   # :436
                # NOTE (:453-459): without covar the reference sets sigma_u_i = omega_i but diag_sigma_u_i =
                # diag_omega_i, which is only defined when covar || sv (:231-233) -> an empty sp_mat.  The gamma
                # prior without covar is therefore not usable in the reference TVP sampler; we follow the evident
                # intent (diag_sigma_u_i = kron(I_T, omega_i)).
                sigma_u_i = omega_i.copy()
            elif covar:
                raise ValueError("the covariance block needs a gamma or stochastic-volatility prior")
            else:
                Sbar = sla.solve(prior_scale + u @ u.T, diag_k)                           # :461
                sigma_u_i = wishrnd(Sbar, vs.chi2("wishart_chi2", draw, post_df - np.arange(k)),
                                    vs.normal("wishart_normal", draw, (k * (k - 1) // 2,)))
            sig_t = np.tile(sigma_u_i[None], (tt, 1, 1))
            if covar:                                                                     # :454-455 (frozen diag_omega_i)
                sig_t = np.einsum("tba,tb,tbc->tac", Psi_t, omega_diag.reshape(tt, k), Psi_t)

        # state variances and initial state (:469-483)
        a_lag = np.concatenate([a_init, a[: (tt - 1) * n]])
        a_v = (a - a_lag).reshape(n, tt, order="F")
        v_post_scale = 1.0 / (v_prior_rate + (a_v ** 2).sum(axis=1) * 0.5)
        sigma_v_i = vs.gamma("sigma_v_gamma", draw, v_post_shape) * v_post_scale          # :476
        iv = init_vi + np.diag(sigma_v_i)
        imu = sla.solve(iv, init_vi @ init_mu + sigma_v_i * a[:n])
        a_init = imu + sla.solve_triangular(sla.cholesky(iv, lower=False),
                                            vs.normal("a_init_normal", draw, (n,)), lower=False)

        if covar:                                                                         # :485-499
            psi_lag = np.concatenate([psi_init, psi[: (tt - 1) * n_psi]])
            psi_v = (psi - psi_lag).reshape(n_psi, tt, order="F")
            psc = 1.0 / (psi_prior_rate + (psi_v ** 2).sum(axis=1) * 0.5)
            psi_sigma_v_i = vs.gamma("psi_sigma_v_gamma", draw, psi_post_shape) * psc     # :492
            piv = psi_prior_vi + np.diag(psi_sigma_v_i)
            pim = sla.solve(piv, psi_prior_vi @ psi_prior_mu + psi_sigma_v_i * psi[:n_psi])
            psi_init = pim + sla.solve_triangular(sla.cholesky(piv, lower=False),
                                                  vs.normal("psi_init_normal", draw, (n_psi,)), lower=False)   # :498

        if draw >= burnin:
            pos = draw - burnin
            if per_t_sigma:
                for i in range(tt):
                    draws_sigma_u[i * k * k:(i + 1) * k * k, pos] = sla.solve(sig_t[i], diag_k).reshape(-1, order="F")
                if sv:
                    draws_sigma_sigma[:, pos] = np.diag(sigma_h).reshape(-1, order="F")
            else:
                draws_sigma_u[:, pos] = sla.solve(sigma_u_i, diag_k).reshape(-1, order="F")
            draws_coef[:, pos] = a
            draws_sigma_v[:, pos] = 1.0 / sigma_v_i                                       # :525
            if bvs:
                draws_lambda[:, pos] = lam
            if covar and psi_bvs:
                draws_psi_lambda[:, pos] = psi_lam                                        # :517-519

    posteriors = {"a0": None, "a": None, "b": None, "c": None, "sigma": None}
    _split_rows(posteriors, draws_coef, draws_lambda, n_a, n_b, n_c, tvp_T=tt, sig=draws_sigma_v)
    if model.get("structural") and k > 1:                                                 # :513-516,590-603
        a0_from = n_a + n_b + n_c
        cube = draws_coef.reshape(tt, n, -1)
        posteriors["a0"] = {"coeffs": cube[:, a0_from:, :].reshape(tt * (n - a0_from), -1), "sigma": draws_sigma_v[a0_from:]}
        if draws_lambda is not None:
            posteriors["a0"]["lambda"] = draws_lambda[a0_from:]
    posteriors["sigma"] = {"coeffs": draws_sigma_u}
    if sv:
        posteriors["sigma"]["sigma"] = draws_sigma_sigma
    if draws_psi_lambda is not None:
        posteriors["sigma"]["lambda"] = draws_psi_lambda                                  # :605-613
    return {"data": obj["data"], "model": obj["model"], "priors": obj["priors"], "posteriors": posteriors}


def _tvp_state_banded(Zt, sig_t, sigma_v_i, a_init, y, lam, eps):
    """mean + R^-1 eps of the (nT) x (nT) block-tridiagonal precision, via banded LAPACK (structured path)."""
    tt = len(Zt)
    n = Zt[0].shape[1]
    N = n * tt
    bw = n                                             # upper bandwidth: dense diagonal block + diagonal off-block
    ab = np.zeros((bw + 1, N))
    rhs = np.zeros(N)
    rr, cc = np.triu_indices(n)
    for t in range(tt):
        Z = Zt[t] * lam[None, :]
        blk = Z.T @ sig_t[t] @ Z + np.diag((2.0 if t < tt - 1 else 1.0) * sigma_v_i)
        ab[bw + rr - cc, t * n + cc] = blk[rr, cc]
        rhs[t * n:(t + 1) * n] = Z.T @ sig_t[t] @ y[:, t]
        if t < tt - 1:
            ab[bw - n, (t + 1) * n + np.arange(n)] = -sigma_v_i
    rhs[:n] += sigma_v_i * a_init
    cb = sla.cholesky_banded(ab, lower=False)
    mean = sla.cho_solve_banded((cb, False), rhs)
    w = sla.solve_banded((0, bw), cb, eps)             # R w = eps, R upper banded
    return mean + w
