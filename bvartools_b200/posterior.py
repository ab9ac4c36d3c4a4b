"""Driver and output objects, mirrored from the reference so that callers can switch unchanged.

  draw_posterior(object, FUN=None, mc_cores=None)   R/draw_posterior.bvarmodel.R:39-86
  bvarpost(object)                                  R/bvarpost.R:56-121
  bvar(...)                                         R/bvar.R:149-315 (the parts the samplers feed)
  summary(bvar)                                     R/summary.bvar.R:121-140 (means, SD, quantiles)

`draw_posterior` keeps the reference's plugin hook (`FUN`) and its failure semantics (a failing model becomes
`{"error": True}`, R/draw_posterior.bvarmodel.R:81-83).  `mc_cores` is accepted and ignored: forked children must
not inherit a CUDA context, and all chains / models already run concurrently on the GPU.
"""
from __future__ import annotations

import numpy as np

import ctypes as C

from . import sampler
from ._lib import check, lib, require_device

__all__ = ["draw_posterior", "bvarpost", "bvar", "summary", "summary_bvarlist", "information_criteria", "thin", "thin_bvarlist", "irf",
           "fevd", "predict"]


def bvar(y, x=None, A0=None, A=None, B=None, C=None, Sigma=None):
    """Collect posterior draws in a `bvar` object (R/bvar.R:149-315): each entry holds draws as iter x rows
    (the transposed `coda::mcmc` layout of R/bvar_fill_helper.R:7-27); TVP entries are lists over t."""
    y = np.asarray(y)
    tt, k = y.shape
    out = {"y": y, "x": x}
    spec = {"dims": {"K": k}, "lags": {}, "tvp": {}, "structural": A0 is not None}

    def fill(entry, n_rows_const):
        if entry is None:
            return None, None, None, False
        coeffs = entry["coeffs"] if isinstance(entry, dict) else entry
        coeffs = np.asarray(coeffs)
        rows = coeffs.shape[0]
        tvp = rows != n_rows_const and n_rows_const > 0 and rows == n_rows_const * tt
        if tvp:
            draws = [coeffs[t * n_rows_const:(t + 1) * n_rows_const].T for t in range(tt)]
        else:
            draws = coeffs.T
        lam = np.asarray(entry["lambda"]).T if isinstance(entry, dict) and "lambda" in entry else None
        sig = np.asarray(entry["sigma"]).T if isinstance(entry, dict) and "sigma" in entry else None
        return draws, lam, sig, tvp

    if A0 is not None:
        out["A0"], out["A0_lambda"], out["A0_sigma"], spec["tvp"]["A0"] = fill(A0, k * k)
    if A is not None:
        rows = np.asarray(A["coeffs"] if isinstance(A, dict) else A).shape[0]
        n_a = rows if rows % (k * k) == 0 and rows // (k * k) < tt else rows // tt
        if rows % tt == 0 and (rows // tt) % (k * k) == 0 and rows // tt >= k * k and isinstance(A, dict) and "sigma" in A:
            n_a = rows // tt
        out["A"], out["A_lambda"], out["A_sigma"], spec["tvp"]["A"] = fill(A, n_a)
        spec["lags"]["p"] = n_a // (k * k)
    else:
        spec["lags"]["p"] = 0
    if B is not None:
        rows = np.asarray(B["coeffs"] if isinstance(B, dict) else B).shape[0]
        n_b = rows // tt if isinstance(B, dict) and "sigma" in B else rows
        out["B"], out["B_lambda"], out["B_sigma"], spec["tvp"]["B"] = fill(B, n_b)
    if C is not None:
        rows = np.asarray(C["coeffs"] if isinstance(C, dict) else C).shape[0]
        n_c = rows // tt if isinstance(C, dict) and "sigma" in C else rows
        out["C"], out["C_lambda"], out["C_sigma"], spec["tvp"]["C"] = fill(C, n_c)
    if Sigma is not None:
        out["Sigma"], out["Sigma_lambda"], out["Sigma_sigma"], spec["tvp"]["Sigma"] = fill(Sigma, k * k)
    out["specifications"] = spec
    out["class"] = "bvar"
    return out


def bvarpost(obj, seed=None, **kw):
    """Posterior simulation of one model (R/bvarpost.R:56-121): picks the TVP / constant sampler and wraps the
    draws in a `bvar` object."""
    res = sampler.bvartvpalg(obj, seed=seed, **kw) if obj["model"].get("tvp") else sampler.bvaralg(obj, seed=seed, **kw)
    post = res["posteriors"]
    if post["sigma"] is not None and "lambda" in post["sigma"]:
        # psi indicators -> k^2 rows, NA on the diagonal, the same draw in both triangles (R/bvarpost.R:64-71)
        k = np.asarray(obj["data"]["Y"]).shape[1]
        lam = np.asarray(post["sigma"]["lambda"])
        full = np.full((k * k, lam.shape[1]), np.nan)
        lower = [i + k * j for j in range(k) for i in range(k) if i > j]            # which(lower.tri), column-major
        upper = [i + k * j for j in range(k) for i in range(k) if i < j]
        full[lower] = lam
        full[upper] = lam
        post = dict(post)
        post["sigma"] = dict(post["sigma"], **{"lambda": full})
    A0 = None
    if obj["model"].get("structural") and post.get("a0") is not None:
        # k^2 x draws with the identity and the strict lower triangle (column-major) = the a0 draws (R/bvarpost.R:73-108)
        k = np.asarray(obj["data"]["Y"]).shape[1]
        lower = [i + k * j for j in range(k) for i in range(k) if i > j]
        a0 = post["a0"]
        draws = np.asarray(a0["coeffs"]).shape[1]
        if obj["model"].get("tvp"):                                 # k^2 T rows, period-major (:84-87)
            tt = np.asarray(obj["data"]["Y"]).shape[0]
            coeffs = np.tile(np.eye(k).reshape(-1, 1, order="F"), (tt, draws))
            rows = [t * k * k + r for t in range(tt) for r in lower]
            coeffs[rows] = a0["coeffs"]
        else:
            coeffs = np.tile(np.eye(k).reshape(-1, 1, order="F"), (1, draws))
            coeffs[lower] = a0["coeffs"]
        A0 = {"coeffs": coeffs}
        if "sigma" in a0:                                           # :93-96
            sg = np.zeros((k * k, draws))
            sg[lower] = a0["sigma"]
            A0["sigma"] = sg
        if "lambda" in a0:
            lam = np.full((k * k, draws), np.nan)
            lam[lower] = a0["lambda"]
            A0["lambda"] = lam
    return bvar(y=obj["data"]["Y"], x=obj["data"]["Z"], A0=A0, A=post["a"], B=post["b"], C=post["c"],
                Sigma=post["sigma"])


def draw_posterior(obj, FUN=None, mc_cores=None, seed=None, **kw):
    """S3 method draw_posterior.bvarmodel (R/draw_posterior.bvarmodel.R:39-86)."""
    only_one = isinstance(obj, dict)
    models = [obj] if only_one else list(obj)
    out = []
    for i, m in enumerate(models):
        try:
            if FUN is None:
                out.append(bvarpost(m, seed=None if seed is None else seed + i, **kw))
            else:
                out.append(FUN(m))
        except Exception as exc:                      # try(...) -> list(error = TRUE)
            out.append({"error": True, "message": str(exc)})
    return out[0] if only_one else out


def summary(obj, ci=(0.025, 0.5, 0.975), n_chains=1, period=None):
    """Posterior means, SDs, naive and time-series SEs and quantiles of a `bvar` object (R/summary.bvar.R:121-140,
    185-203, i.e. coda::summary.mcmc per draw matrix).  The draw matrices are uploaded once and reduced on the device
    (bvg_summary_draws: exact type-7 quantiles; bvg_tsse_draws: spectrum0.ar).  n_chains: the draws are `n_chains`
    equally long chains stacked chain-major (1 = a single chain, as in the reference).  period: 0-based period for TVP /
    SV entries (lists over t); default the last one (:33-38)."""
    import torch
    require_device()
    dp = C.POINTER(C.c_double)
    probs = np.ascontiguousarray(ci, dtype=np.float64)
    res = {}
    tt = np.asarray(obj["y"]).shape[0]
    if period is not None and not 0 <= int(period) < tt:
        raise ValueError("Implausible specification of argument 'period'.")
    for name in ("A0", "A", "B", "C", "Sigma"):
        d = obj.get(name)
        if d is None:
            continue
        if isinstance(d, list):
            d = d[tt - 1 if period is None else int(period)]
        d = np.ascontiguousarray(d, dtype=np.float64)
        n, rows = d.shape
        dev = torch.from_numpy(d).cuda()
        mean, sd, q = np.empty(rows), np.empty(rows), np.empty((probs.size, rows))
        check(lib().bvg_summary_draws(C.c_void_p(dev.data_ptr()), n, rows, probs.ctypes.data_as(dp), probs.size,
                                      mean.ctypes.data_as(dp), sd.ctypes.data_as(dp), q.ctypes.data_as(dp), None))
        res[name] = {"mean": mean, "sd": sd, "naive_se": sd / np.sqrt(n), "quantiles": q}
        if n % n_chains == 0 and n // n_chains >= 3:
            ts = np.empty(rows)
            check(lib().bvg_tsse_draws(C.c_void_p(dev.data_ptr()), n_chains, n // n_chains, rows,
                                       ts.ctypes.data_as(dp), None))
            res[name]["ts_se"] = ts
    return res


def thin(obj, thin=10):
    """thin.bvar (R/thin.bvar.R:31-67): keep every `thin`-th draw (positions thin, 2 thin, ... in R's 1-based count) of
    every draw matrix of a `bvar` object.  Plain row indexing on the host object -- no device path is needed."""
    out = dict(obj)
    draws = None
    for name in ("A0", "A", "B", "C", "Sigma"):
        d = obj.get(name)
        if d is not None:
            draws = (d[0] if isinstance(d, list) else d).shape[0]
            break
    if draws is None:
        return out
    pos = np.arange(thin, draws + 1, thin) - 1
    for base in ("A", "B", "C", "A0", "Sigma"):
        for name in (base, base + "_sigma", base + "_lambda"):
            d = obj.get(name)
            if d is None:
                continue
            out[name] = [np.asarray(x)[pos] for x in d] if isinstance(d, list) else np.asarray(d)[pos]
    return out


def thin_bvarlist(objects, thin=10):
    """thin.bvarlist (R/thin.bvarlist.R): thin every `bvar` object of a list (error objects pass through)."""
    return [o if o.get("error") else globals()["thin"](o, thin) for o in objects]


def information_criteria(ll, tot_pars, tt):
    """LL -> AIC / BIC / HQ as in R/summary.bvarlist.R:181-184 (tot_pars = NCOL(x) for a VAR, :62)."""
    return {"LL": float(ll), "AIC": 2.0 * tot_pars - 2.0 * ll, "BIC": np.log(tt) * tot_pars - 2.0 * ll,
            "HQ": 2.0 * np.log(np.log(tt)) * tot_pars - 2.0 * ll}


def summary_bvarlist(objects):
    """summary method for a list of `bvar` objects (R/summary.bvarlist.R:22-215): one row per model with p, s, LL, AIC,
    BIC and HQ.  The reference loops over every draw in R and calls loglik_normal per draw (:160-180); here the draws
    go through one fused device reduction (bvg_loglik_draws_host) that returns the per-period sums."""
    require_device()
    dp = C.POINTER(C.c_double)
    rows = []
    for obj in objects:
        if obj.get("error"):
            rows.append(None)                                   # :41-45 `next`
            continue
        y = np.ascontiguousarray(np.asarray(obj["y"], dtype=np.float64))          # T x K: (t, i) at i + K t
        tt, k = y.shape
        x = obj.get("x")
        tvp_flags = obj["specifications"]["tvp"]
        tvp = any(bool(v) for v in tvp_flags.values())
        sv = bool(tvp_flags.get("Sigma"))
        parts = [obj[name] for name in ("A", "B", "C") if obj.get(name) is not None]   # :65-76 cbind(A, B, C)
        if parts and isinstance(parts[0], list):
            coef = np.stack([np.concatenate([p[t] for p in parts], axis=1) for t in range(tt)], axis=1)  # [draw][T][n]
        elif parts:
            coef = np.concatenate(parts, axis=1)                                     # [draw][n]
        else:
            coef = None
        m = 0 if x is None else np.asarray(x).shape[1]
        xs = None if x is None else np.ascontiguousarray(np.asarray(x, dtype=np.float64))   # T x M: (t, j) at j + M t
        sig = obj["Sigma"]
        sig = np.stack(sig, axis=1) if isinstance(sig, list) else np.asarray(sig)    # [draw][T][k k] or [draw][k k]
        sig = np.ascontiguousarray(sig, dtype=np.float64)
        draws = sig.shape[0]
        coef_c = None if coef is None else np.ascontiguousarray(coef, dtype=np.float64)
        coef_is_tvp = coef_c is not None and coef_c.ndim == 3
        row = {"p": obj["specifications"]["lags"].get("p"), "s": obj["specifications"]["lags"].get("s")}
        if coef_c is not None and not coef_is_tvp and not sv:
            # constant model: total log-likelihood straight from cross moments (HBM-bound, no loop over periods)
            tot = C.c_double(0.0)
            check(lib().bvg_loglik_total_host(coef_c.ctypes.data_as(dp), sig.ctypes.data_as(dp), draws, k, m, tt,
                                              y.ctypes.data_as(dp), xs.ctypes.data_as(dp), C.byref(tot)))
            ll = tot.value / draws
            row["ll_t"] = None
        else:
            ll_t = np.empty(tt, dtype=np.float64)
            check(lib().bvg_loglik_draws_host(None if coef_c is None else coef_c.ctypes.data_as(dp),
                                              sig.ctypes.data_as(dp), draws, k, m, tt, int(coef_is_tvp), int(sv),
                                              y.ctypes.data_as(dp), None if xs is None else xs.ctypes.data_as(dp),
                                              ll_t.ctypes.data_as(dp)))
            ll = float(np.sum(ll_t / draws))                                         # sum(rowMeans(LL)), :181
            row["ll_t"] = ll_t / draws
        row.update(information_criteria(ll, m, tt))
        rows.append(row)
    return rows


def _device_draws(obj, period):
    """A (draws x k^2 p) and Sigma (draws x k^2) of a `bvar` object as device tensors (torch is only the allocator)."""
    import torch
    require_device()
    y = np.asarray(obj["y"])
    tt, k = y.shape
    A = obj.get("A")
    if isinstance(A, list):
        A = A[tt - 1 if period is None else period]
    S = obj.get("Sigma")
    if isinstance(S, list):
        S = S[tt - 1 if period is None else period]
    if S is None:
        raise ValueError("Argument 'object' must include draws of the variance-covariance matrix Sigma.")
    S = np.ascontiguousarray(S, dtype=np.float64)
    draws = S.shape[0]
    A = np.zeros((draws, 0)) if A is None else np.ascontiguousarray(A, dtype=np.float64)
    p = A.shape[1] // (k * k)
    dA = torch.from_numpy(A).cuda() if A.size else None
    dS = torch.from_numpy(S).cuda()
    return k, p, draws, dA, dS


def irf(obj, impulse, response, n_ahead=5, ci=0.95, shock=1, type="feir", cumulative=False, keep_draws=False,
        period=None):
    """irf.bvar (R/irf.bvar.R:78-216) for a host `bvar` object: `impulse` / `response` are 0-based variable positions.
    The draws are uploaded once; the per-draw responses and their bands are computed on the device."""
    import torch
    types = {"feir": 0, "oir": 1, "gir": 2, "sir": 3, "sgir": 4}
    if type not in types:
        raise ValueError("Argument 'type' not known.")
    if type in ("sir", "sgir") and obj.get("A0") is None:
        raise ValueError("Structural IR requires that draws of 'A0' are contained in the 'bvar' x.")
    if isinstance(shock, str):
        if shock not in ("sd", "nsd"):
            raise ValueError("Invalid specification of argument 'shock'.")
        kind, val = (1 if shock == "sd" else 2), 0.0
    else:
        kind, val = 0, float(shock)
    k, p, draws, dA, dS = _device_draws(obj, period)
    structural = type in ("sir", "sgir")
    if p == 0 and not structural:                                  # R/irf.bvar.R:96-98
        raise ValueError("Impulse responses only supported for models with p > 0, i.e. argument 'x' must contain "
                         "element 'A', or structural models.")
    out = torch.empty((draws, n_ahead + 1), dtype=torch.float64, device="cuda")
    vp = C.c_void_p
    d0 = None
    if structural:
        A0 = obj["A0"]
        if isinstance(A0, list):
            A0 = A0[np.asarray(obj["y"]).shape[0] - 1 if period is None else period]
        d0 = torch.from_numpy(np.ascontiguousarray(A0, dtype=np.float64)).cuda()      # draws x k^2, column-major matrices
    check(lib().bvg_irf_draws_a0(None if dA is None else vp(dA.data_ptr()), 0 if dA is None else dA.shape[1],
                                 vp(dS.data_ptr()), k * k, None if d0 is None else vp(d0.data_ptr()), k * k, 0, draws, k, p,
                                 int(n_ahead), types[type], int(impulse), int(response), kind, val,
                                 int(bool(cumulative)), vp(out.data_ptr()), None))
    if keep_draws:
        return out.cpu().numpy()
    lo = (1.0 - ci) / 2.0
    probs = np.ascontiguousarray([lo, 0.5, 1.0 - lo])
    q = np.empty((3, n_ahead + 1))
    dp = C.POINTER(C.c_double)
    check(lib().bvg_summary_draws(vp(out.data_ptr()), draws, n_ahead + 1, probs.ctypes.data_as(dp), 3, None, None,
                                  q.ctypes.data_as(dp), None))
    return q.T


def fevd(obj, response, n_ahead=5, type="oir", normalise_gir=False, period=None):
    """fevd.bvar (R/fevd.bvar.R:72-186) for a host `bvar` object: (n_ahead + 1) x k table, mean over the draws."""
    import torch
    types = {"oir": 1, "gir": 2, "sir": 3, "sgir": 4}
    if type not in types:
        raise ValueError("The specified type of the used impulse response is not known.")
    structural = type in ("sir", "sgir")
    if structural and obj.get("A0") is None:
        raise ValueError("Structural FEVD requires that draws of 'A0' are contained in the 'bvar' object.")
    k, p, draws, dA, dS = _device_draws(obj, period)
    if p == 0 and not structural:                                  # R/fevd.bvar.R:93-95
        raise ValueError("Variance decompositions only supported for models with p > 0, i.e. argument 'object' must "
                         "contain element 'A', or structural models.")
    out = np.empty((n_ahead + 1, k))
    vp = C.c_void_p
    d0 = None
    if structural:
        A0 = obj["A0"]
        if isinstance(A0, list):
            A0 = A0[np.asarray(obj["y"]).shape[0] - 1 if period is None else period]
        d0 = torch.from_numpy(np.ascontiguousarray(A0, dtype=np.float64)).cuda()
    check(lib().bvg_fevd_draws_a0(None if dA is None else vp(dA.data_ptr()), 0 if dA is None else dA.shape[1],
                                  vp(dS.data_ptr()), k * k, None if d0 is None else vp(d0.data_ptr()), k * k, 0, draws, k, p,
                                  int(n_ahead), types[type], int(response), int(bool(normalise_gir)),
                                  out.ctypes.data_as(C.POINTER(C.c_double)), None))
    return out


def predict(obj, n_ahead=10, new_x=None, new_d=None, ci=0.95, seed=None, eps=None, keep_draws=False, period=None):
    """predict.bvar (R/predict.bvar.R:36-260) for a host `bvar` object: the prediction matrix is built as in :176-201
    (last p observations of y; `new_x` / `new_d` future exogenous / deterministic values, zeros when omitted as in the
    reference), the draws cbind(A, B, C) and Sigma are uploaded once and every draw is simulated forward on the device
    (bvg_forecast_draws).  Returns {variable index: (n_ahead x 3) bands} or the draws [draws][n_ahead][k]."""
    import torch
    from .sampler import _seed
    require_device()
    y = np.asarray(obj["y"], dtype=np.float64)
    tt, k = y.shape
    pick = lambda d: d[tt - 1 if period is None else period]
    parts = []
    for name in ("A", "B", "C"):
        d = obj.get(name)
        if d is not None:
            parts.append(np.asarray(pick(d) if isinstance(d, list) else d, dtype=np.float64))
    if not parts:
        raise ValueError("Argument 'object' must contain draws of at least one of 'A', 'B' or 'C'.")
    S = obj.get("Sigma")
    if S is None:
        raise ValueError("Argument 'object' must include draws of the variance-covariance matrix Sigma.")
    S = np.ascontiguousarray(pick(S) if isinstance(S, list) else S, dtype=np.float64)
    coef = np.ascontiguousarray(np.concatenate(parts, axis=1))
    draws = coef.shape[0]
    A = obj.get("A")
    p = 0 if A is None else (A[0] if isinstance(A, list) else np.asarray(A)).shape[1] // (k * k)
    n_x = coef.shape[1] // k - k * p
    x = None if obj.get("x") is None else np.asarray(obj["x"], dtype=np.float64)
    n_det = 0 if new_d is None and new_x is not None else n_x
    if new_x is not None:
        new_x = np.atleast_2d(np.asarray(new_x, dtype=np.float64))
        n_det = n_x - new_x.shape[1]
    ylen = k * p if p > 0 else k
    tot = ylen + n_x
    pred = np.zeros((tot, n_ahead + 1))
    pred[:ylen, 0] = y[tt - 1::-1][:p].reshape(-1) if p > 0 else y[tt - 1]
    if n_x > 0:
        pred[ylen:, 0] = x[tt - 1, k * p:]
        fut = np.zeros((n_ahead, n_x))
        if new_x is not None:
            if new_x.shape[0] < n_ahead:
                raise ValueError("Longer time series required for argument 'new_x'.")
            fut[:, :new_x.shape[1]] = new_x[:n_ahead]
        if new_d is not None:
            new_d = np.asarray(new_d, dtype=np.float64).reshape(-1, n_det)
            if new_d.shape[0] != n_ahead:
                raise ValueError("Length of argument 'new_d' must be equal to 'n.ahead'.")
            fut[:, n_x - n_det:] = new_d
        pred[ylen:, 1:] = fut.T
    predf = np.ascontiguousarray(pred.reshape(-1, order="F"))
    dA, dS = torch.from_numpy(coef).cuda(), torch.from_numpy(S).cuda()
    dE = None if eps is None else torch.from_numpy(np.ascontiguousarray(eps, dtype=np.float64)).cuda()
    out = torch.empty((draws, n_ahead, k), dtype=torch.float64, device="cuda")
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    d0 = None
    if obj.get("A0") is not None:                                  # structural: y = A0^-1 (A x + u), :200-209
        A0 = obj["A0"]
        d0 = torch.from_numpy(np.ascontiguousarray(pick(A0) if isinstance(A0, list) else A0, dtype=np.float64)).cuda()
    check(lib().bvg_forecast_draws_a0(vp(dA.data_ptr()), coef.shape[1], vp(dS.data_ptr()), k * k,
                                      None if d0 is None else vp(d0.data_ptr()), k * k, 0, draws, k, p, tot,
                                      int(n_ahead), predf.ctypes.data_as(dp), None if dE is None else vp(dE.data_ptr()),
                                      _seed(seed), vp(out.data_ptr()), None))
    if keep_draws:
        return out.cpu().numpy()
    lo = (1.0 - ci) / 2.0
    probs = np.ascontiguousarray([lo, 0.5, 1.0 - lo])
    q = np.empty((3, n_ahead * k))
    check(lib().bvg_summary_draws(vp(out.data_ptr()), draws, n_ahead * k, probs.ctypes.data_as(dp), 3, None, None,
                                  q.ctypes.data_as(dp), None))
    q = q.reshape(3, n_ahead, k)
    return {i: q[:, :, i].T for i in range(k)}

