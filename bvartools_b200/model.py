"""Model set-up: the *input contract* of the Gibbs-sampler hot path.

The reference builds the `bvarmodel` nested list in R (`gen_var` + `add_priors`); the C++ samplers discover
everything by name from that list (reference src/bvaralg.cpp:9-249).  R is not available here, so this module
restates exactly the parts of the set-up the hot path consumes, producing the same nested structure (a dict with
the reference's element names) so the parity tests read like calls into the reference.  It is bookkeeping that
runs once per model in O(ms) and is not part of the accelerated path.

References (file:line into the reference tree):
  gen_var            R/gen_var.R:74-305
  add_priors         R/add_priors.bvarmodel.R:160-621
  minnesota_prior    R/minnesota_prior.R:61-247
  ssvs_prior         R/ssvs_prior.R:37-101
  inclusion_prior    R/inclusion_prior.R:63-204
"""
from __future__ import annotations

import copy
from typing import Iterable, Sequence

import numpy as np

__all__ = ["gen_var", "add_priors", "minnesota_prior", "ssvs_prior", "inclusion_prior", "BvarModel"]


class BvarModel(dict):
    """A `bvarmodel`: dict with keys data / model / priors / initial (R/gen_var.R:288-303)."""


def _as_2d(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    return a


def _lagname(base: Sequence[str], i: int, width: int, sep: str = ".") -> list[str]:
    return [f"{b}{sep}{str(i).zfill(max(2, width))}" for b in base]


def gen_var(data, p=2, exogen=None, s=None, deterministic="const", seasonal=False, structural=False,
            tvp=False, sv=False, fcst=None, iterations=50000, burnin=5000, names=None, exo_names=None,
            frequency=1, start_cycle=1):
    """VAR model generator (R/gen_var.R:74-305).

    `data` is a (rows x k) array (the R version takes a `ts`).  `p` (and `s`) may be a sequence: all models then
    share the sample of `max(p)` (R/gen_var.R:109,116-128,172).  Returns one BvarModel or a list of them.
    """
    data = _as_2d(data)
    nobs, k = data.shape
    if names is None:
        names = [f"y{i + 1}" for i in range(k)] if k > 1 else ["y"]
    names = list(names)
    if deterministic not in ("none", "const", "trend", "both"):
        raise ValueError("Argument 'deterministic' must be 'none', 'const', 'trend' or 'both'.")
    if seasonal and deterministic not in ("const", "both"):                       # R/gen_var.R:91-93
        raise ValueError("Argument 'deterministic' must be either 'const' or 'both' when using 'seasonal = TRUE'.")
    if k == 1 and structural:
        raise ValueError("Model must contain at least two endogenous variables for structural analysis.")

    p_list = [int(v) for v in (p if isinstance(p, Iterable) else [p])]
    p_max = max(p_list)
    use_exo = exogen is not None
    if use_exo:
        exogen = _as_2d(exogen)
        if exogen.shape[0] != nobs:
            raise ValueError("'exogen' must cover the same periods as 'data'.")
        m = exogen.shape[1]
        if exo_names is None:
            exo_names = [f"x{i + 1}" for i in range(m)] if m > 1 else ["x"]
        s_list = [int(v) for v in (s if isinstance(s, Iterable) else [0 if s is None else s])]
        s_max = max(s_list)
    else:
        m, s_list, s_max = 0, [0], 0

    # cbind(data, lag(data,-1), ...) then na.omit (R/gen_var.R:116-128,172): rows drop = max(p_max, s_max)
    drop = max(p_max, s_max)
    tt = nobs - drop
    cols, colnames = [data[drop:]], list(names)
    for i in range(1, p_max + 1):
        cols.append(data[drop - i: nobs - i])
        colnames += _lagname(names, i, len(str(p_max)))
    if use_exo:
        for i in range(0, s_max + 1):
            cols.append(exogen[drop - i: nobs - i])
            colnames += _lagname(exo_names, i, len(str(s_max)), sep=".l")
    det_name = []
    if deterministic in ("const", "both"):
        cols.append(np.ones((tt, 1)))
        colnames.append("const")
        det_name.append("const")
    if deterministic in ("trend", "both"):
        cols.append(np.arange(1, tt + 1, dtype=np.float64)[:, None])
        colnames.append("trend")
        det_name.append("trend")
    if seasonal and int(frequency) > 1:
        # R/gen_var.R:188-203: frequency - 1 dummies; the R version reads frequency / cycle from the `ts` attributes,
        # here `frequency` and `start_cycle` (cycle, 1-based, of the first row of `data`) are arguments
        freq = int(frequency)
        cyc = (int(start_cycle) - 1 + drop + np.arange(tt)) % freq + 1
        first = int(np.argmax(cyc == 1)) + 1                                       # which(cycle(temp) == 1)[1]
        pos = (list(range(1, freq + 1)) * 2)[first - 1: first + freq - 2]
        for i in range(freq - 1):
            s_temp = np.zeros(freq)
            s_temp[pos[i] - 1] = 1.0
            cols.append(np.resize(s_temp, tt)[:, None])                            # rep(s_temp, length.out = tt)
            colnames.append(f"season.{i + 1}")
            det_name.append(f"season.{i + 1}")
    temp = np.hstack(cols)

    model = {"type": "VAR", "endogen": {"variables": names, "lags": 0}}
    if use_exo:
        model["exogen"] = {"variables": list(exo_names), "lags": 0}
    if det_name:
        model["deterministic"] = det_name
    model.update(structural=bool(structural), tvp=bool(tvp), sv=bool(sv),
                 iterations=int(iterations), burnin=int(burnin))

    y = temp[:, :k].copy()
    fcst_y = None
    n_keep = tt
    if fcst is not None:
        fcst_y = y[tt - fcst:]
        y = y[: tt - fcst]
        n_keep = tt - fcst

    y_a0 = None
    if structural and k > 1:
        full = np.kron(-y, np.eye(k))                      # R/gen_var.R:243-251
        drop_cols = [j * k + i for j in range(k) for i in range(j + 1)]
        y_a0 = np.delete(full, drop_cols, axis=1)

    result = []
    for pi in p_list:
        for sj in s_list:
            pos = []
            model_i = copy.deepcopy(model)
            if pi >= 1:
                pos += list(range(k, k + k * pi))
                model_i["endogen"]["lags"] = pi
            if use_exo:
                pos += list(range(k + k * p_max, k + k * p_max + m * (sj + 1)))
                model_i["exogen"]["lags"] = sj
            if det_name:
                base = k + k * p_max + (m * (s_max + 1) if use_exo else 0)
                pos += list(range(base, base + len(det_name)))
            x = z = None
            xnames = None
            if pos:
                x = temp[:n_keep, pos].copy()
                xnames = [colnames[c] for c in pos]
                z = np.kron(x, np.eye(k))                   # R/gen_var.R:281
            if y_a0 is not None:
                z = y_a0 if z is None else np.hstack([z, y_a0])
            result.append(BvarModel(data={"Y": y, "Z": x, "SUR": z, "TEST": fcst_y,
                                          "Y_names": names, "Z_names": xnames},
                                    model=model_i))
    return result[0] if len(result) == 1 else result


def _ols_resid_var(y_row: np.ndarray, x_rows: np.ndarray) -> float:
    """y (I - x'(xx')^-1 x) y' / (T - rows)   (R/minnesota_prior.R:150)."""
    tt = y_row.shape[0]
    beta = np.linalg.solve(x_rows @ x_rows.T, x_rows @ y_row)
    r = y_row - beta @ x_rows
    return float(r @ r) / (tt - x_rows.shape[0])


def minnesota_prior(obj, kappa0=2.0, kappa1=0.5, kappa2=None, kappa3=5.0, max_var=None, coint_var=False,
                    sigma="AR"):
    """Minnesota prior for VAR models (R/minnesota_prior.R:61-247); VEC branch not restated."""
    if any(v is not None and v <= 0 for v in (kappa0, kappa1, kappa2, kappa3)):
        raise ValueError("Kappa arguments must be positive.")
    if obj["model"]["type"] != "VAR":
        raise NotImplementedError("VEC models are out of scope")
    y = obj["data"]["Y"].T
    x = obj["data"]["Z"].T
    k, tt = y.shape
    p = obj["model"]["endogen"]["lags"]
    m = sx = 0
    if "exogen" in obj["model"]:
        m = len(obj["model"]["exogen"]["variables"])
        sx = obj["model"]["exogen"]["lags"]
    n_det = len(obj["model"].get("deterministic", []))
    V = np.full((k, x.shape[0]), np.nan)

    beta = np.linalg.solve(x @ x.T, x @ y.T).T
    u = y - beta @ x
    ols_sigma = u @ u.T / (tt - x.shape[0])

    pos_det = list(range(k * p + (sx + 1) * m, k * p + (sx + 1) * m + n_det))
    if sigma == "AR":
        s_endo = np.zeros(k)
        if p > 0 or pos_det:
            for i in range(k):
                pos = [i + k * r for r in range(p)] + pos_det
                s_endo[i] = _ols_resid_var(y[i], x[pos])
        else:
            s_endo = y.var(axis=1, ddof=1)
    elif sigma == "VAR":
        s_endo = np.diag(ols_sigma).copy()
    else:
        raise ValueError("Argument 'sigma' must be either 'AR' or 'VAR'.")
    s_endo = np.sqrt(s_endo)

    for r in range(1, p + 1):
        for l in range(k):
            for j in range(k):
                if l == j:
                    V[l, (r - 1) * k + j] = kappa0 / r ** 2
                else:
                    V[l, (r - 1) * k + j] = kappa0 * kappa1 / r ** 2 * s_endo[l] ** 2 / s_endo[j] ** 2
    if m > 0:
        s_exo = np.sqrt(x[p * k: p * k + m].var(axis=1, ddof=1))
        for r in range(1, sx + 2):
            for l in range(k):
                for j in range(m):
                    if kappa2 is None:
                        V[l, p * k + (r - 1) * m + j] = kappa0 * kappa3 * s_endo[l] ** 2
                    else:
                        V[l, p * k + (r - 1) * m + j] = kappa0 * kappa2 / r ** 2 * s_endo[l] ** 2 / s_exo[j] ** 2
    if max_var is not None:
        V = np.where(V > max_var, max_var, V)
    if n_det:
        # the reference indexes with m * s (not m * (s + 1)), R/minnesota_prior.R:206 -- kept as is
        V[:, k * p + m * sx:] = (kappa0 * kappa3 * s_endo ** 2)[:, None]

    mu = np.zeros((k, x.shape[0]))
    if coint_var:
        mu[:k, :k] = np.eye(k)
    mu = mu.reshape(-1, 1, order="F")
    Vv = V.reshape(-1, 1, order="F")
    if obj["model"]["structural"] and k > 1:
        mu = np.vstack([mu, np.zeros((k * (k - 1) // 2, 1))])
        vs = []
        for j in range(k - 1):
            vs += list(kappa0 * kappa1 * s_endo[j + 1:] ** 2 / s_endo[j] ** 2)
        Vv = np.vstack([Vv, np.asarray(vs)[:, None]])
    res = {"mu": mu, "v_i": np.diag(1.0 / Vv[:, 0])}
    if not obj["model"]["structural"]:
        res["sigma_i"] = np.linalg.inv(ols_sigma)
    return res


def ssvs_prior(obj, tau=(0.05, 10.0), semiautomatic=None):
    """SSVS prior scales (R/ssvs_prior.R:37-101), VAR branch."""
    if obj["model"]["sv"]:
        raise ValueError("SSVS cannot be used with models with stochastic volatility.")
    if obj["model"]["tvp"]:
        raise ValueError("SSVS cannot be used with models with time varying parameter.")
    y = obj["data"]["Y"].T
    x = obj["data"]["Z"].T
    z = obj["data"]["SUR"]
    k, tt = y.shape
    tot_par = z.shape[1]
    covar = tot_par > k * x.shape[0]
    if semiautomatic is not None:
        xxi = np.linalg.inv(x @ x.T)
        ols = y @ x.T @ xxi
        u = y - ols @ x
        sigma_ols = u @ u.T / (tt - x.shape[0])
        se_ols = np.sqrt(np.diag(np.kron(xxi, sigma_ols)))[:, None]
        tau0 = se_ols * semiautomatic[0]
        tau1 = se_ols * semiautomatic[1]
        if covar:
            nc = k * (k - 1) // 2
            tau0 = np.vstack([tau0, np.full((nc, 1), tau[0])])
            tau1 = np.vstack([tau1, np.full((nc, 1), tau[1])])
    else:
        tau0 = np.full((tot_par, 1), float(tau[0]))
        tau1 = np.full((tot_par, 1), float(tau[1]))
    return {"tau0": tau0, "tau1": tau1}


def inclusion_prior(obj, prob=0.5, exclude_deterministics=True, minnesota_like=False,
                    kappa=(0.8, 0.5, 0.5, 0.8)):
    """Prior inclusion probabilities and the 1-based `include` index (R/inclusion_prior.R:63-204), VAR branch."""
    if not minnesota_like and not (0 <= prob <= 1):
        raise ValueError("Argument 'prob' must be between 0 and 1.")
    k = len(obj["model"]["endogen"]["variables"])
    p = obj["model"]["endogen"]["lags"]
    m = sx = 0
    if "exogen" in obj["model"]:
        m = len(obj["model"]["exogen"]["variables"])
        sx = obj["model"]["exogen"]["lags"]
    n = len(obj["model"].get("deterministic", []))
    nx = obj["data"]["Z"].shape[1]
    result = np.full((k, nx), np.nan)
    include = np.arange(1, k * nx + 1)
    if minnesota_like:
        for i in range(1, p + 1):
            blk = np.full((k, k), kappa[1] / i)
            np.fill_diagonal(blk, kappa[0] / i)
            result[:, (i - 1) * k: i * k] = blk
        if m > 0:
            result[:, p * k: p * k + m] = kappa[2]
            for i in range(1, sx + 1):
                result[:, p * k + m + (i - 1) * m: p * k + m + i * m] = kappa[2] / (1 + i)
        if n > 0:
            result[:, p * k + (sx + 1) * m:] = kappa[3]
    else:
        result[:, :] = prob
    result = result.reshape(-1, 1, order="F")
    if obj["model"]["structural"]:
        ns = k * (k - 1) // 2
        result = np.vstack([result, np.full((ns, 1), prob)])
        include = np.concatenate([include, k * nx + np.arange(1, ns + 1)])
    if n > 0 and exclude_deterministics:
        start = k * (k * p + (sx + 1) * m)
        pos_det = set(range(start, start + k * n))          # 0-based positions inside `include`
        include = np.asarray([v for i, v in enumerate(include) if i not in pos_det], dtype=np.int64)
    return {"prior": result, "include": include.reshape(-1, 1)}


_DEF_COEF = dict(v_i=1, v_i_det=0.1, shape=3, rate=0.0001, rate_det=0.01)
_DEF_SIGMA = dict(df="k", scale=1, mu=0, v_i=0.01, sigma_h=0.05, constant=0.0001)


def add_priors(obj, coef=None, sigma=None, ssvs=None, bvs=None):
    """Attach priors and initial values (R/add_priors.bvarmodel.R:160-621).

    Restated branches: regular / Minnesota / SSVS / BVS coefficient priors, TVP state-variance priors, Wishart /
    gamma / SV error priors, the covariance (psi) block (`sigma$covar`, constant-parameter models), OLS initial
    values.
    """
    coef = dict(_DEF_COEF) if coef is None else dict(coef)
    sigma = dict(_DEF_SIGMA) if sigma is None else dict(sigma)
    only_one = isinstance(obj, dict)
    objs = [copy.deepcopy(obj)] if only_one else [copy.deepcopy(o) for o in obj]

    if coef.get("v_i") is not None:
        if coef["v_i"] < 0:
            raise ValueError("Argument 'v_i' must be at least 0.")
        if coef.get("v_i_det") is None:
            coef["v_i_det"] = coef["v_i"]
    elif not any(key in coef for key in ("minnesota", "ssvs")):
        raise ValueError("If 'coef$v_i' is not specified, at least 'coef$minnesota' or 'coef$ssvs' must be specified.")
    if len(sigma) < 2:
        raise ValueError("Argument 'sigma' must be at least of length 2.")
    covar = bool(sigma.get("covar"))

    any_sv = any(o["model"]["sv"] for o in objs)
    if any_sv:
        if any(key not in sigma for key in ("mu", "v_i", "shape", "rate")):
            raise ValueError("Missing prior specifications for stochastic volatility prior.")
        error_prior = "sv"
    else:
        error_prior = None
        if all(key in sigma for key in ("shape", "rate")):
            error_prior = "gamma"
        if all(key in sigma for key in ("df", "scale")):
            error_prior = "wishart"
        if error_prior is None:
            raise ValueError("Invalid specification for argument 'sigma'.")
        if error_prior == "wishart" and any(o["model"]["structural"] for o in objs):
            raise ValueError("Structural models may not use a Wishart prior.")
    minnesota = coef.get("minnesota") is not None
    use_ssvs_error = use_bvs_error = False
    use_ssvs = ssvs is not None
    if use_ssvs:
        ssvs = dict(ssvs)
        if ssvs.get("inprior") is None:
            raise ValueError("Argument 'ssvs$inprior' must be specified for SSVS.")
        if ssvs.get("tau") is None and ssvs.get("semiautomatic") is None:
            raise ValueError("Either argument 'ssvs$tau' or 'ssvs$semiautomatic' must be specified for SSVS.")
        ssvs.setdefault("exclude_det", False)
        minnesota = False
        if ssvs.get("covar"):
            if error_prior == "wishart":
                raise ValueError("If SSVS should be applied to error covariances, argument 'sigma$shape' and "
                                 "'sigma$rate' must be specified.")
            if ssvs.get("tau") is None:
                raise ValueError("If SSVS should be applied to error covariances, argument 'ssvs$tau' must be specified.")
            use_ssvs_error = True
    use_bvs = bvs is not None
    if use_bvs:
        bvs = dict(bvs)
        if bvs.get("inprior") is None:
            raise ValueError("If BVS should be applied, argument 'bvs$inprior' must be specified.")
        bvs.setdefault("exclude_det", False)
        if bvs.get("covar"):
            if error_prior == "wishart":
                raise ValueError("If BVS should be applied to error covariances, argument 'sigma$shape' must be specified.")
            use_bvs_error = True
    if use_ssvs and use_bvs:
        raise ValueError("SSVS and BVS cannot be applied at the same time.")

    for o in objs:
        k = len(o["model"]["endogen"]["variables"])
        p = o["model"]["endogen"]["lags"]
        m = sx = 0
        if "exogen" in o["model"]:
            m = len(o["model"]["exogen"]["variables"])
            sx = o["model"]["exogen"]["lags"] + 1
        n_a, n_b = k * k * p, k * m * sx
        n_det = len(o["model"].get("deterministic", [])) * k
        tot_par = n_a + n_b + n_det
        structural = o["model"]["structural"]
        n_struct = 0
        if structural and k > 1:
            n_struct = (k - 1) * k // 2
            tot_par += n_struct
        pri = o.setdefault("priors", {})
        ini = o.setdefault("initial", {})

        if tot_par > 0:
            mu = np.zeros((k, (tot_par - n_struct) // k))
            if coef.get("coint_var") and p > 0:
                mu[:k, :k] = np.eye(k)
            if n_det > 0 and coef.get("const") is not None:
                zn = o["data"]["Z_names"]
                if zn.count("const") == 1:
                    pos = zn.index("const")
                    c = coef["const"]
                    if isinstance(c, str):
                        if c == "first":
                            mu[:, pos] = o["data"]["Y"][0]
                        elif c == "mean":
                            mu[:, pos] = o["data"]["Y"].mean(axis=0)
                        else:
                            raise ValueError("Invalid specificatin of coef$const.")
                    else:
                        mu[:, pos] = c
            mu = mu.reshape(-1, 1, order="F")
            if structural:
                mu = np.vstack([mu, np.zeros((n_struct, 1))])
            pc = pri["coefficients"] = {"mu": mu}
            minn = None
            if minnesota:
                mk = coef["minnesota"]
                minn = minnesota_prior(o, kappa0=mk.get("kappa0", 2), kappa1=mk.get("kappa1", 0.5),
                                       kappa2=mk.get("kappa2"), kappa3=mk.get("kappa3", 5))
                pc["v_i"] = minn["v_i"]
            if use_ssvs:
                st = ssvs_prior(o, tau=ssvs.get("tau", (0.05, 10.0)), semiautomatic=ssvs.get("semiautomatic"))
                ip = inclusion_prior(o, prob=ssvs["inprior"], exclude_deterministics=ssvs["exclude_det"],
                                     minnesota_like=ssvs.get("minnesota") is not None,
                                     kappa=ssvs.get("minnesota") or (0.8, 0.5, 0.5, 0.8))
                o["model"]["varselect"] = "SSVS"
                pc["v_i"] = np.diag(1.0 / st["tau1"][:, 0] ** 2)
                pc["ssvs"] = {"inprior": ip["prior"], "include": ip["include"],
                              "tau0": st["tau0"], "tau1": st["tau1"]}
            if not minnesota and not use_ssvs:
                v_i = np.diag(np.full(tot_par, float(coef["v_i"])))
                if n_det > 0 and coef.get("v_i_det") is not None:
                    idx = np.arange(tot_par - n_struct - n_det, tot_par - n_struct)
                    v_i[idx, idx] = coef["v_i_det"]
                pc["v_i"] = v_i
            if use_bvs:
                o["model"]["varselect"] = "BVS"
                ip = inclusion_prior(o, prob=bvs["inprior"], exclude_deterministics=bvs["exclude_det"],
                                     minnesota_like=bvs.get("minnesota") is not None,
                                     kappa=bvs.get("minnesota") or (0.8, 0.5, 0.5, 0.8))
                pc["bvs"] = {"inprior": ip["prior"], "include": ip["include"]}
            if o["model"]["tvp"]:
                pc["shape"] = np.full((tot_par, 1), float(coef["shape"]))
                pc["rate"] = np.full((tot_par, 1), float(coef["rate"]))
                if n_det > 0 and coef.get("rate_det") is not None:
                    pc["rate"][tot_par - n_struct - n_det: tot_par - n_struct, 0] = coef["rate_det"]

        # covariance priors (R/add_priors.bvarmodel.R:494-503): psi ~ N(0, v_i^-1 I) on the k (k - 1) / 2 free elements
        if covar and structural:
            raise ValueError("Error covariances and structural coefficients cannot be estimated at the same time.")
        if covar and not structural and k > 1:
            if not o["model"]["sv"] and error_prior != "gamma":
                raise ValueError("The covariance block needs a gamma ('sigma$shape', 'sigma$rate') or stochastic-volatility "
                                 "error prior.")
            n_covar = k * (k - 1) // 2
            pri["psi"] = {"mu": np.zeros((n_covar, 1)), "v_i": np.diag(np.full(n_covar, float(coef["v_i"])))}
            if o["model"]["tvp"]:                            # :499-502
                pri["psi"]["shape"] = np.full((n_covar, 1), float(coef["shape"]))
                pri["psi"]["rate"] = np.full((n_covar, 1), float(coef["rate"]))
            if use_ssvs_error:                               # :505-511
                pri["psi"]["ssvs"] = {"inprior": np.full((n_covar, 1), float(ssvs["inprior"])),
                                      "include": np.arange(1, n_covar + 1).reshape(-1, 1),
                                      "tau0": np.full((n_covar, 1), float(ssvs["tau"][0])),
                                      "tau1": np.full((n_covar, 1), float(ssvs["tau"][1]))}
            if use_bvs_error:                                # :514-517
                pri["psi"]["bvs"] = {"inprior": np.full((n_covar, 1), float(bvs["inprior"])),
                                     "include": np.arange(1, n_covar + 1).reshape(-1, 1)}

        ps = pri["sigma"] = {}
        if o["model"]["sv"]:
            ps["mu"] = np.full((k, 1), float(sigma["mu"]))
            ps["v_i"] = np.diag(np.full(k, float(sigma["v_i"])))
            ps["shape"] = np.full((k, 1), float(sigma["shape"]))
            ps["rate"] = np.full((k, 1), float(sigma["rate"]))
        else:
            help_df = sigma["df"] if error_prior == "wishart" else sigma["shape"]
            if isinstance(help_df, str):
                if "k" not in help_df:
                    raise ValueError("Use no other letter than 'k' in 'sigma$df'.")
                help_df = float(eval(help_df, {"__builtins__": {}}, {"k": k}))
            if help_df < 0:
                raise ValueError("Current specification implies a negative prior degree of freedom or shape.")
            if error_prior == "wishart":
                if sigma["scale"] <= 0:
                    raise ValueError("Argument 'sigma$scale' must be larger than 0.")
                ps.update(type="wishart", df=float(help_df), scale=np.diag(np.full(k, float(sigma["scale"]))))
            else:
                if sigma["rate"] <= 0:
                    raise ValueError("Argument 'sigma$rate' must be larger than 0.")
                ps.update(type="gamma", shape=np.full((k, 1), float(help_df)),
                          rate=np.full((k, 1), float(sigma["rate"])))
            if minnesota and tot_par > 0 and minn is not None and "sigma_i" in minn:
                ps["sigma_i"] = minn["sigma_i"]

        # Initial values (R/add_priors.bvarmodel.R:571-613)
        y = o["data"]["Y"].T
        tt = y.shape[1]
        if 0 < tot_par < tt:
            z = o["data"]["SUR"]
            yv = y.reshape(-1, 1, order="F")
            ols = np.linalg.solve(z.T @ z, z.T @ yv)
            u = (yv - z @ ols).reshape(k, tt, order="F")
            ini["coefficients"] = {"draw": ols}
        else:
            if tot_par > 0:
                ini["coefficients"] = {"draw": pri["coefficients"]["mu"].copy()}
            u = y - y.mean(axis=1, keepdims=True)
        if o["model"]["tvp"] and tot_par > 0:
            ini["coefficients"]["sigma_i"] = np.diag(1.0 / pri["coefficients"]["rate"][:, 0])
        if covar and not structural and k > 1:               # :586-598 (psi by OLS of u on -u, then u <- Psi u)
            yc = np.kron(-u.T, np.eye(k))
            yc = np.delete(yc, [j * k + i for j in range(k) for i in range(j + 1)], axis=1)
            psi0 = np.linalg.solve(yc.T @ yc, yc.T @ u.reshape(-1, 1, order="F"))
            ini["psi"] = {"draw": psi0}
            Psi = np.eye(k)
            for j in range(1, k):
                Psi[j, :j] = psi0[(j - 1) * j // 2: (j - 1) * j // 2 + j, 0]
            u = Psi @ u
        uvar = u.var(axis=1, ddof=1)
        if o["model"]["sv"]:
            ini["sigma"] = {"h": np.log(np.tile(uvar[None, :], (tt, 1))),
                            "sigma_h": np.full((k, 1), float(sigma["sigma_h"])),
                            "constant": np.full((k, 1), float(sigma.get("constant", 0.0001)))}
        else:
            ini["sigma"] = {"sigma_i": np.diag(1.0 / uvar)}
    return objs[0] if only_one else objs
