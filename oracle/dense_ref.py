"""ctypes wrapper of oracle/bvar_dense_ref.c -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline "port")."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_DIR = pathlib.Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libbvar_dense_ref.so"
_lib = None


def load():
    global _lib
    if _lib is None:
        src = _DIR / "bvar_dense_ref.c"
        if not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
            subprocess.check_call(["make", "-s", "-C", str(_DIR)])
        _lib = C.CDLL(str(_SO))
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _lib.bvar_dense_ref.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, dp, dp, C.c_int, dp, dp,
                                        dp, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                        dp, dp, dp, dp, dp, dp, dp]
        _lib.bvar_dense_ref.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def bvar_dense(obj, n_chains=1, seed=0, n_threads=1, dense_gemm=True, inject=None, iterations=None, burnin=None):
    """Run the dense C restatement (Wishart prior, optional SSVS).  Returns dict of [chain][iter][rows] arrays."""
    lib = load()
    Y = np.asarray(obj["data"]["Y"], dtype=np.float64)
    T, k = Y.shape
    Z = np.asfortranarray(obj["data"]["SUR"], dtype=np.float64)
    n = Z.shape[1]
    M = n // k
    pc, ps = obj["priors"]["coefficients"], obj["priors"]["sigma"]
    if obj["model"].get("sv") or obj["model"].get("tvp") or "df" not in ps or "bvs" in pc:
        raise NotImplementedError("dense C port covers the Wishart prior with optional SSVS")
    it = int(obj["model"]["iterations"] if iterations is None else iterations)
    bi = int(obj["model"]["burnin"] if burnin is None else burnin)
    yv = np.ascontiguousarray(Y.reshape(-1))                       # time-major vec(y), y = t(Y)
    mu = np.ascontiguousarray(pc["mu"], dtype=np.float64).reshape(-1)
    Vi = np.asfortranarray(pc["v_i"], dtype=np.float64)
    S0 = np.asfortranarray(ps["scale"], dtype=np.float64)
    sig0 = np.asfortranarray(obj["initial"]["sigma"]["sigma_i"], dtype=np.float64)
    ssvs = "ssvs" in pc
    tau0 = tau1 = inprior = None
    inc = np.zeros(0, dtype=np.int32)
    if ssvs:
        tau0 = np.ascontiguousarray(pc["ssvs"]["tau0"], dtype=np.float64).reshape(-1)
        tau1 = np.ascontiguousarray(pc["ssvs"]["tau1"], dtype=np.float64).reshape(-1)
        inprior = np.ascontiguousarray(pc["ssvs"]["inprior"], dtype=np.float64).reshape(-1)
        inc = np.ascontiguousarray(np.asarray(pc["ssvs"]["include"]).reshape(-1) - 1, dtype=np.int32)
    inj = inject or {}
    arrs = {kname: (np.ascontiguousarray(inj[kname], dtype=np.float64) if kname in inj else None)
            for kname in ("coef_normal", "wishart_chi2", "wishart_normal", "ssvs_u")}
    out_coef = np.empty((n_chains, it, n))
    out_sigma = np.empty((n_chains, it, k * k))
    out_lambda = np.empty((n_chains, it, n)) if ssvs else None
    rc = lib.bvar_dense_ref(k, T, M, _p(yv), _p(Z), _p(mu), _p(Vi), float(ps["df"]) + T, _p(S0), _p(sig0), int(ssvs),
                            _p(tau0), _p(tau1), _p(inprior), inc.ctypes.data_as(C.POINTER(C.c_int)), int(inc.shape[0]),
                            it, bi, int(n_chains), int(seed), int(n_threads), int(bool(dense_gemm)),
                            _p(arrs["coef_normal"]), _p(arrs["wishart_chi2"]), _p(arrs["wishart_normal"]),
                            _p(arrs["ssvs_u"]), _p(out_coef), _p(out_sigma), _p(out_lambda))
    if rc != 0:
        raise RuntimeError(f"bvar_dense_ref failed with status {rc}")
    res = {"coef": out_coef, "sigma": out_sigma}
    if ssvs:
        res["lambda"] = out_lambda
    return res
