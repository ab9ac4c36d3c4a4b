"""Host-side mirror of the reference's sampler entry points, calling the CUDA library through the C ABI.

  bvaralg(object)     <-> .bvaralg(object)     (R/RcppExports.R:4-6,  src/bvaralg.cpp:6)
  bvartvpalg(object)  <-> .bvartvpalg(object)  (R/RcppExports.R:8-10, src/bvartvpalg.cpp:9)

Same input list, same return value: `{data, model, priors, posteriors}` with `posteriors$a|b|c$coeffs` (rows x
iterations), `$lambda`, TVP `$sigma`, and `posteriors$sigma${coeffs,sigma}` (src/bvaralg.cpp:562-627,
src/bvartvpalg.cpp:555-626).  In addition `run_chains` runs many independent chains of one model in a single
launch (the batched entry of SURVEY.md section 8b) and `ChainSampler` keeps the draws resident in HBM.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi
from ._lib import BvgError, check, lib, require_device

__all__ = ["bvaralg", "bvartvpalg", "run_chains", "ChainSampler", "pinned_empty", "split_posteriors"]


class _Pinned(np.ndarray):
    """ndarray view of cudaHostAlloc memory; frees it when the base array dies."""


def pinned_empty(shape, dtype=np.float64):
    """Page-locked host array (for overlapped device->host copies of the draws)."""
    L = lib()
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = L.bvg_host_alloc(max(nbytes, 8))
    if not ptr:
        raise MemoryError("cudaHostAlloc failed")
    buf = (C.c_char * max(nbytes, 8)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner:
        def __init__(self, p):
            self.p = p

        def __del__(self):
            try:
                L.bvg_host_free(self.p)
            except Exception:
                pass
    out = arr.view(_Pinned)
    out._owner = _Owner(ptr)
    out._buf = buf
    return out


def _seed(seed):
    if seed is None:
        return int.from_bytes(os.urandom(8), "little")
    return int(seed)


class ChainSampler:
    """Handle API: inputs uploaded once; retained draws stay in HBM until fetched (bvg_sampler_*)."""

    def __init__(self, obj, n_chains=1, seed=None, chain_offset=0, device=-1, inject=None, **tuning):
        require_device()
        self.spec = _abi.spec_from_model(obj, n_chains=n_chains, seed=_seed(seed), chain_offset=chain_offset,
                                         device=device, inject=inject, **tuning)
        self.lags = int(obj.get("model", {}).get("endogen", {}).get("lags", 0))     # p: A = first k*k*p coefficients
        self._obj = obj
        self.handle = C.c_void_p()
        check(lib().bvg_sampler_create(C.byref(self.spec.c), C.byref(self.handle)))

    def reset(self):
        check(lib().bvg_sampler_reset(self.handle))

    def run(self, stream=0):
        check(lib().bvg_sampler_run(self.handle, C.c_void_p(stream)))

    def sync(self):
        check(lib().bvg_sampler_sync(self.handle))

    def run_to_host(self, out: _abi.OutBuffers, stream=0):
        check(lib().bvg_sampler_run_to_host(self.handle, C.byref(out.c), C.c_void_p(stream)))

    def fetch(self, out: _abi.OutBuffers | None = None):
        out = out or _abi.OutBuffers(self.spec)
        check(lib().bvg_sampler_fetch(self.handle, C.byref(out.c)))
        return out

    def device_out(self):
        d = _abi.BvgOut()
        check(lib().bvg_sampler_device_out(self.handle, C.byref(d)))
        return d

    @property
    def last_kernel_ms(self):
        return lib().bvg_sampler_last_kernel_ms(self.handle)

    @property
    def launch_count(self):
        return lib().bvg_sampler_launch_count(self.handle)

    @property
    def kernel_name(self):
        return lib().bvg_sampler_kernel_name(self.handle).decode()

    def loglik(self, stream=0):
        """Per-period log-likelihood SUMS over all retained draws resident on the device (chains x iterations):
        the reduction behind summary.bvarlist (R/summary.bvarlist.R:160-186, src/loglik_normal.cpp:37-58).
        Returns (ll_t, n_draws); LL of the reference = ll_t.sum() / n_draws."""
        T = int(self.spec.c.T)
        out = np.empty(T, dtype=np.float64)
        check(lib().bvg_sampler_loglik(self.handle, out.ctypes.data_as(C.POINTER(C.c_double)), C.c_void_p(stream or 0)))
        return out, int(self.spec.c.n_chains) * int(self.spec.c.iterations)

    def loglik_total(self, stream=0):
        """Total log-likelihood SUM over all resident retained draws (all that LL / AIC / BIC / HQ need); constant
        models take the HBM-bound cross-moment kernel (no pass over the periods).  Returns (total, n_draws)."""
        tot = C.c_double(0.0)
        check(lib().bvg_sampler_loglik_total(self.handle, C.byref(tot), C.c_void_p(stream or 0)))
        return tot.value, int(self.spec.c.n_chains) * int(self.spec.c.iterations)

    def loglik_total_device(self, dev_ptr, stream=0):
        """The same sum written asynchronously to one double of DEVICE memory (`dev_ptr`: integer address): no host
        synchronisation, e.g. for one all-reduce over a whole model grid."""
        check(lib().bvg_sampler_loglik_total_device(self.handle, C.c_void_p(int(dev_ptr)), C.c_void_p(stream or 0)))
        return int(self.spec.c.n_chains) * int(self.spec.c.iterations)

    _WHICH = {"coef": 0, "sigma": 1, "lambda": 2, "sigma_h": 3, "sigma_v": 4}

    def summary(self, which="coef", ci=0.95, probs=None, stream=0, ts_se=True):
        """Posterior summary of one output over ALL retained draws resident on the device (chains x iterations), as
        summary.bvar / coda::summary.mcmc report them (R/summary.bvar.R:110-140): mean, sd, naive_se and the quantiles
        (ci_low, 0.5, ci_high) of R's type 7 -- exact order statistics from a radix select in HBM -- and ts_se, coda's
        time-series SE (AR spectral estimate at frequency zero per chain).  For `lambda` the mean is the posterior
        inclusion frequency."""
        if probs is None:
            lo = (1.0 - ci) / 2.0
            probs = (lo, 0.5, 1.0 - lo)
        probs = np.ascontiguousarray(probs, dtype=np.float64)
        rows = int(self.spec.out_rows()[which])
        mean, sd = np.empty(rows), np.empty(rows)
        q = np.empty((probs.size, rows))
        dp = C.POINTER(C.c_double)
        check(lib().bvg_sampler_summary(self.handle, self._WHICH[which], probs.ctypes.data_as(dp), probs.size,
                                        mean.ctypes.data_as(dp), sd.ctypes.data_as(dp), q.ctypes.data_as(dp),
                                        C.c_void_p(stream or 0)))
        n = int(self.spec.c.n_chains) * int(self.spec.c.iterations)
        res = {"mean": mean, "sd": sd, "naive_se": sd / np.sqrt(n), "quantiles": q, "probs": probs, "n": n}
        if ts_se and int(self.spec.c.iterations) >= 3:
            res["ts_se"] = self.time_series_se(which, stream)
        return res

    def time_series_se(self, which="coef", stream=0):
        """coda's "Time-series SE" of one output (R/summary.bvar.R:133,141): spectrum0.ar -- an AIC-selected
        Yule-Walker AR fit -- per chain and row on the device, pooled over the chains as summary.mcmc.list does."""
        rows = int(self.spec.out_rows()[which])
        ts = np.empty(rows)
        check(lib().bvg_sampler_tsse(self.handle, self._WHICH[which], ts.ctypes.data_as(C.POINTER(C.c_double)),
                                     C.c_void_p(stream or 0)))
        return ts

    def irf(self, impulse, response, n_ahead=5, ci=0.95, shock=1, type="feir", cumulative=False, keep_draws=False,
            period=None, p=None, stream=0):
        """Impulse responses over ALL retained draws resident on the device (irf.bvar, R/irf.bvar.R:78-216): `impulse`
        / `response` are 0-based variable indices, `shock` a number or "sd" / "nsd", `type` "feir" / "oir" / "gir", and
        for structural models "sir" / "sgir".
        Returns the (n_ahead + 1) x 3 band matrix (ci_low, median, ci_high) like the reference, or the draws
        (draws x (n_ahead + 1)) with keep_draws."""
        types = {"feir": 0, "oir": 1, "gir": 2, "sir": 3, "sgir": 4}
        if type not in types:
            raise ValueError("Argument 'type' not known.")
        if type in ("sir", "sgir") and not self.spec.c.structural:
            raise ValueError("Structural IR requires that draws of 'A0' are contained in the 'bvar' x.")
        if isinstance(shock, str):
            if shock not in ("sd", "nsd"):
                raise ValueError("Invalid specification of argument 'shock'.")
            kind, val = (1 if shock == "sd" else 2), 0.0
        else:
            kind, val = 0, float(shock)
        k = int(self.spec.c.k)
        if p is None:
            p = self.lags
        lo = (1.0 - ci) / 2.0
        probs = np.ascontiguousarray([lo, 0.5, 1.0 - lo])
        n = int(self.spec.c.n_chains) * int(self.spec.c.iterations)
        q = np.empty((3, n_ahead + 1))
        draws = np.empty((n, n_ahead + 1)) if keep_draws else None
        dp = C.POINTER(C.c_double)
        check(lib().bvg_sampler_irf(self.handle, int(p), int(n_ahead), types[type], int(impulse), int(response), kind, val,
                                    int(bool(cumulative)), -1 if period is None else int(period), probs.ctypes.data_as(dp),
                                    0 if keep_draws else 3, q.ctypes.data_as(dp), None,
                                    None if draws is None else draws.ctypes.data_as(dp), C.c_void_p(stream or 0)))
        return draws if keep_draws else q.T

    def fevd(self, response, n_ahead=5, type="oir", normalise_gir=False, period=None, p=None, stream=0):
        """Forecast error variance decomposition averaged over ALL retained draws resident on the device (fevd.bvar,
        R/fevd.bvar.R:72-186): the (n_ahead + 1) x k table for the 0-based `response` variable; type "oir" / "gir", and
        "sir" / "sgir" for structural models."""
        types = {"oir": 1, "gir": 2, "sir": 3, "sgir": 4}
        if type not in types:
            raise ValueError("The specified type of the used impulse response is not known.")
        if type in ("sir", "sgir") and not self.spec.c.structural:
            raise ValueError("Structural FEVD requires that draws of 'A0' are contained in the 'bvar' object.")
        k = int(self.spec.c.k)
        out = np.empty((n_ahead + 1, k))
        check(lib().bvg_sampler_fevd(self.handle, int(self.lags if p is None else p), int(n_ahead), types[type],
                                     int(response), int(bool(normalise_gir)), -1 if period is None else int(period),
                                     out.ctypes.data_as(C.POINTER(C.c_double)), C.c_void_p(stream or 0)))
        return out

    def predict(self, n_ahead=10, new_x=None, new_d=None, ci=0.95, seed=None, eps=None, keep_draws=False, stream=0):
        """Forecasts from ALL retained draws resident on the device (predict.bvar, R/predict.bvar.R:36-260).  The
        prediction matrix is set up as in the reference (:176-201): last p observations of y, `new_x` (n_ahead x m)
        for exogenous regressors, `new_d` (n_ahead x n) for deterministic terms -- zeros when omitted, exactly like the
        reference (pass ones to keep an intercept).  Returns {variable index: (n_ahead x 3) bands (ci_low, median,
        ci_high)} or the draws [draws][n_ahead][k] with keep_draws.  `eps`: injected N(0,1) of that shape (tests)."""
        Y, Z = np.asarray(self._obj["data"]["Y"]), np.asarray(self._obj["data"]["Z"])
        tt, k = Y.shape
        p = self.lags
        n_x = Z.shape[1] - k * p                            # exogenous + deterministic regressors
        n_det = 0 if new_d is None and new_x is not None else n_x
        if new_x is not None:
            new_x = np.atleast_2d(np.asarray(new_x, dtype=np.float64))
            n_det = n_x - new_x.shape[1]
        tot = (k * p if p > 0 else k) + n_x
        pred = np.zeros((tot, n_ahead + 1))
        ylen = k * p if p > 0 else k
        pred[:ylen, 0] = (Y[tt - 1::-1][:p].reshape(-1) if p > 0 else Y[tt - 1])     # t(y[tt:(tt - p + 1), ])
        if n_x > 0:
            pred[ylen:, 0] = Z[tt - 1, k * p:]                                       # latest observation of x / d
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
        lo = (1.0 - ci) / 2.0
        probs = np.ascontiguousarray([lo, 0.5, 1.0 - lo])
        n = int(self.spec.c.n_chains) * int(self.spec.c.iterations)
        q = np.empty((3, n_ahead * k))
        draws = np.empty((n, n_ahead, k)) if keep_draws else None
        dp = C.POINTER(C.c_double)
        epsc = None if eps is None else np.ascontiguousarray(eps, dtype=np.float64)
        check(lib().bvg_sampler_forecast(self.handle, p, tot, int(n_ahead), predf.ctypes.data_as(dp),
                                         None if epsc is None else epsc.ctypes.data_as(dp), _seed(seed), probs.ctypes.data_as(dp),
                                         0 if keep_draws else 3, q.ctypes.data_as(dp), None,
                                         None if draws is None else draws.ctypes.data_as(dp), C.c_void_p(stream or 0)))
        if keep_draws:
            return draws
        q = q.reshape(3, n_ahead, k)
        return {i: q[:, :, i].T for i in range(k)}

    def close(self):
        if self.handle:
            lib().bvg_sampler_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_chains(obj, n_chains=1, seed=None, chain_offset=0, device=-1, inject=None, pinned=False, **tuning):
    """Run `n_chains` independent chains of one model; returns dict of [chain][iter][rows] arrays + 'status'."""
    require_device()
    spec = _abi.spec_from_model(obj, n_chains=n_chains, seed=_seed(seed), chain_offset=chain_offset, device=device,
                                inject=inject, **tuning)
    out = _abi.OutBuffers(spec, alloc=pinned_empty if pinned else None)
    fn = lib().bvg_bvartvpalg if spec.c.tvp else lib().bvg_bvaralg
    check(fn(C.byref(spec.c), C.byref(out.c)))
    res = dict(out.arrays)
    res["status"] = out.status
    return res


def split_posteriors(obj, arrays, chain=0):
    """[iter][rows] draws of one chain -> the reference's `posteriors` list (rows x iterations matrices)."""
    model = obj["model"]
    Y = np.asarray(obj["data"]["Y"])
    T, k = Y.shape
    p = int(model["endogen"]["lags"])
    n_a = k * k * p
    n_b = 0
    if "exogen" in model:
        n_b = k * len(model["exogen"]["variables"]) * (int(model["exogen"]["lags"]) + 1)
    n_c = k * len(model.get("deterministic", []))
    n_a0 = k * (k - 1) // 2 if model.get("structural") and k > 1 else 0       # last rows (bvaralg.cpp:262-263)
    n = n_a + n_b + n_c + n_a0
    tvp = bool(model.get("tvp"))
    post = {"a0": None, "a": None, "b": None, "c": None, "sigma": None}
    coef = arrays.get("coef")
    lam = arrays.get("lambda")
    # the lambda buffer holds [coefficient indicators (n, if selection on coefficients) | psi indicators (k (k - 1) / 2)]
    n_psi = k * (k - 1) // 2
    pc = obj.get("priors", {}).get("coefficients", {})
    coef_sel = ("ssvs" in pc) or ("bvs" in pc)
    pp = obj.get("priors", {}).get("psi", {})
    psi_sel = ("ssvs" in pp) or ("bvs" in pp)
    psi_lam = None
    if lam is not None and psi_sel:
        psi_lam = lam[chain][:, (n if coef_sel else 0):(n if coef_sel else 0) + n_psi].T      # draws_lambda_a0, bvaralg.cpp:604-611
    if not coef_sel:
        lam = None
    off = 0
    for name, cnt in (("a", n_a), ("b", n_b), ("c", n_c), ("a0", n_a0)):
        if cnt > 0 and coef is not None:
            d = {}
            c_ = coef[chain]                                        # [iter][rows]
            if tvp:
                it = c_.shape[0]
                cube = c_.reshape(it, T, n)                         # [iter][t][n]
                d["coeffs"] = cube[:, :, off:off + cnt].reshape(it, T * cnt).T      # (cnt*T) x iter, time-major
                d["sigma"] = arrays["sigma_v"][chain][:, off:off + cnt].T
            else:
                d["coeffs"] = c_[:, off:off + cnt].T
            if lam is not None:
                d["lambda"] = lam[chain][:, off:off + cnt].T
            post[name] = d
        off += cnt
    sig = {"coeffs": arrays["sigma"][chain].T}
    if "sigma_h" in arrays:
        sh = arrays["sigma_h"][chain]                               # [iter][k] -> vec(diag(sigma_h)) k*k x iter
        full = np.zeros((sh.shape[0], k, k))
        full[:, np.arange(k), np.arange(k)] = sh
        sig["sigma"] = full.reshape(sh.shape[0], k * k).T
    if psi_lam is not None:
        sig["lambda"] = psi_lam
    post["sigma"] = sig
    return post


def _one(obj, want_tvp, seed, **kw):
    if bool(obj["model"].get("tvp")) != want_tvp:
        raise BvgError(-10, "model$tvp does not match the sampler called")
    arrays = run_chains(obj, n_chains=1, seed=seed, **kw)
    if arrays["status"][0] != 0:
        raise BvgError(1, "a posterior precision matrix was not positive definite")
    return {"data": obj["data"], "model": obj["model"], "priors": obj["priors"],
            "posteriors": split_posteriors(obj, arrays, 0)}


def bvaralg(obj, seed=None, **kw):
    """Drop-in for `.bvaralg(object)`: one chain of the constant-parameter BVAR Gibbs sampler."""
    return _one(obj, False, seed, **kw)


def bvartvpalg(obj, seed=None, **kw):
    """Drop-in for `.bvartvpalg(object)`: one chain of the TVP BVAR Gibbs sampler."""
    return _one(obj, True, seed, **kw)
