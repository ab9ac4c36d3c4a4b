"""Shared helpers of the test-suite: model builders for the BASELINE configs, injected-variate plumbing, and
loaders for (a) the test-only kernel-logic emulator and (b) the product library."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from bvartools_b200 import _abi                                          # noqa: E402
from bvartools_b200.data import e1_readme_sample                         # noqa: E402
from bvartools_b200.model import add_priors, gen_var                     # noqa: E402
from oracle import samplers as osamp                                     # noqa: E402
from oracle.variates import VariateSource, draw_variates                 # noqa: E402

EMU_SRC = ROOT / "tests" / "emu" / "bvg_emu.cpp"
EMU_SO = ROOT / "tests" / "emu" / "_bvgemu.so"


def load_emu():
    """Build (if stale) and load the TEST-ONLY host emulation of the kernel logic."""
    srcs = [EMU_SRC] + list((ROOT / "bvartools_b200" / "csrc").glob("*.cuh")) + \
           list((ROOT / "bvartools_b200" / "csrc").glob("*.hpp")) + [ROOT / "include" / "bvargpu.h"]
    if not EMU_SO.exists() or any(s.stat().st_mtime > EMU_SO.stat().st_mtime for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC",
                               "-ffp-contract=off", "-x", "c++", str(EMU_SRC), "-o", str(EMU_SO)])
    lib = C.CDLL(str(EMU_SO))
    for fn in ("emu_bvaralg", "emu_bvartvpalg"):
        getattr(lib, fn).argtypes = [C.POINTER(_abi.BvgSpec), C.POINTER(_abi.BvgOut)]
        getattr(lib, fn).restype = C.c_int
    lib.emu_last_error.restype = C.c_char_p
    return lib


# ---------------------------------------------------------------- models of the BASELINE configs
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


def model_structural(K=3, p=1, nobs=50, iterations=12, burnin=4, seed=9, sv=False, bvs=False, tvp=False):
    """Structural VAR (A0 coefficients: equation i contains -y_j, j < i) with inverse-gamma or SV error variances."""
    y = synth_var(K, p, nobs, seed)
    m = gen_var(y, p=p, deterministic="const", iterations=iterations, burnin=burnin, structural=True, sv=sv, tvp=tvp)
    sigma = dict(shape=3, rate=1e-3)
    if sv:
        sigma.update(mu=0, v_i=0.01, sigma_h=0.05, constant=1e-4)
    return add_priors(m, coef=dict(v_i=1, v_i_det=0.1, shape=3, rate=1e-4, rate_det=0.01), sigma=sigma,
                      bvs=dict(inprior=0.5, exclude_det=True) if bvs else None)


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


# ---------------------------------------------------------------- injected variates
def variate_kinds(obj):
    model, priors = obj["model"], obj["priors"]
    kinds = []
    if model.get("tvp"):
        kinds += ["state_normal", "sigma_v_gamma", "a_init_normal"]
    elif obj["data"].get("SUR") is not None:
        kinds += ["coef_normal"]
    pc = priors.get("coefficients", {})
    if "ssvs" in pc:
        kinds += ["ssvs_u"]
    if "bvs" in pc:
        kinds += ["bvs_u1", "bvs_u2"]
    if model.get("sv"):
        kinds += ["sv_u", "sv_normal", "sigma_h_gamma", "h_init_normal"]
    elif "df" in priors["sigma"]:
        kinds += ["wishart_chi2", "wishart_normal"]
    else:
        kinds += ["omega_gamma"]
    if "psi" in priors:
        kinds += ["psi_normal"]
        if model.get("tvp"):
            kinds += ["psi_sigma_v_gamma", "psi_init_normal"]
        if "ssvs" in priors["psi"]:
            kinds += ["psi_u1"]
        if "bvs" in priors["psi"]:
            kinds += ["psi_u1", "psi_u2"]
    return kinds


def make_variates(obj, n_chains, seed):
    """Per chain injected arrays; returns (list of per-chain dicts for the oracle, dict of stacked arrays for the ABI)."""
    Y = obj["data"]["Y"]
    T, k = Y.shape
    n = 0 if obj["data"].get("SUR") is None else obj["data"]["SUR"].shape[1]
    sweeps = obj["model"]["iterations"] + obj["model"]["burnin"]
    ps = obj["priors"]["sigma"]
    dims = dict(n=n, k=k, T=T)
    if "shape" in ps:
        sh = np.asarray(ps["shape"], dtype=np.float64).reshape(-1) + 0.5 * T
        dims["shape_h"] = sh
        dims["shape_omega"] = sh
    if "df" in ps:
        dims["df"] = float(ps["df"]) + T
    if obj["model"].get("tvp"):
        dims["shape_v"] = np.asarray(obj["priors"]["coefficients"]["shape"], dtype=np.float64).reshape(-1) + 0.5 * T
        if "psi" in obj["priors"]:
            dims["tvp_psi"] = True
            dims["shape_psi"] = np.asarray(obj["priors"]["psi"]["shape"], dtype=np.float64).reshape(-1) + 0.5 * T
    rng = np.random.default_rng(seed)
    per_chain = [draw_variates(dims, sweeps, rng, variate_kinds(obj)) for _ in range(n_chains)]
    stacked = {kname: np.stack([pc[kname].reshape(sweeps, -1) for pc in per_chain]) for kname in per_chain[0]}
    if obj["model"].get("structural") and k > 1:
        # the library runs structural models on the regressors [x; -y] (M + k of them) with the positions
        # (variable j, equation i <= j) masked: coefficient-level injected arrays are indexed by the internal position
        Mu = obj["data"]["Z"].shape[1]
        pos = list(range(k * Mu)) + [k * (Mu + j) + i for j in range(k) for i in range(j + 1, k)]
        na = k * (Mu + k)
        for kname in ("coef_normal", "bvs_u1", "bvs_u2", "sigma_v_gamma", "a_init_normal"):
            if kname in stacked:
                full = np.full(stacked[kname].shape[:2] + (na,), 0.5)
                full[:, :, pos] = stacked[kname]
                stacked[kname] = full
        if "state_normal" in stacked:                         # [chain][sweep][T * n], period-major
            sn = stacked["state_normal"]
            src = sn.reshape(sn.shape[0], sn.shape[1], T, -1)
            full = np.full(sn.shape[:2] + (T, na), 0.5)
            full[:, :, :, pos] = src
            stacked["state_normal"] = full.reshape(sn.shape[0], sn.shape[1], T * na)
    return per_chain, stacked


def run_abi(fn, obj, n_chains, stacked=None, seed=0, **kw):
    """Call a bvg_bvaralg-shaped entry point (product or emulator); returns dict of [chain][iter][rows] arrays."""
    spec = _abi.spec_from_model(obj, n_chains=n_chains, seed=seed, inject=stacked, **kw)
    out = _abi.OutBuffers(spec)
    rc = fn(C.byref(spec.c), C.byref(out.c))
    return rc, out


def oracle_stack(obj, per_chain, literal=True, table="ksc1998"):
    """Run the oracle per chain with the injected variates; returns dict of [chain][iter][rows] arrays."""
    fn = osamp.bvartvpalg if obj["model"].get("tvp") else osamp.bvaralg
    res = {"coef": [], "sigma": [], "lambda": [], "sigma_h": [], "sigma_v": []}
    k = obj["data"]["Y"].shape[1]
    for inj in per_chain:
        r = fn(obj, VariateSource(injected=inj), literal=literal, table=table)["posteriors"]
        parts = [r[nm] for nm in ("a", "b", "c", "a0") if r.get(nm) is not None]
        if obj["model"].get("tvp"):
            T = obj["data"]["Y"].shape[0]
            it = parts[0]["coeffs"].shape[1]
            cube = np.concatenate([p["coeffs"].reshape(T, -1, it) for p in parts], axis=1)   # [T][n][iter]
            res["coef"].append(cube.reshape(-1, it).T)
            res["sigma_v"].append(np.concatenate([p["sigma"] for p in parts], axis=0).T)
        elif parts:
            res["coef"].append(np.concatenate([p["coeffs"] for p in parts], axis=0).T)
        lam_rows = [p["lambda"] for p in parts] if parts and "lambda" in parts[0] else []
        if "lambda" in r["sigma"]:                          # psi inclusion draws follow the coefficient ones in the buffer
            lam_rows.append(r["sigma"]["lambda"])
        if lam_rows:
            res["lambda"].append(np.concatenate(lam_rows, axis=0).T)
        res["sigma"].append(r["sigma"]["coeffs"].T)
        if "sigma" in r["sigma"]:
            res["sigma_h"].append(r["sigma"]["sigma"].reshape(k, k, -1)[np.arange(k), np.arange(k)].T)
    return {kname: np.stack(v) for kname, v in res.items() if v}


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300 + 1e-3 * np.max(np.abs(b)))))
