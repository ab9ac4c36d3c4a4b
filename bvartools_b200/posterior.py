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

from . import sampler

__all__ = ["draw_posterior", "bvarpost", "bvar", "summary"]


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
    return bvar(y=obj["data"]["Y"], x=obj["data"]["Z"], A0=None, A=post["a"], B=post["b"], C=post["c"],
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


def summary(obj, ci=(0.025, 0.5, 0.975)):
    """Posterior means, SDs and quantiles of a `bvar` object (R/summary.bvar.R:121-140)."""
    res = {}
    for name in ("A", "B", "C", "Sigma"):
        d = obj.get(name)
        if d is None or isinstance(d, list):
            continue
        res[name] = {"mean": d.mean(axis=0), "sd": d.std(axis=0, ddof=1),
                     "quantiles": np.quantile(d, ci, axis=0)}
    return res
