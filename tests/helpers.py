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
from bvartools_b200.workloads import (model_c1, model_grid_c2b, synth_var, model_c3, model_sv, model_covar_gamma, model_tvp_covar, model_structural, model_c5)  # noqa: E402,F401


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
        for kname in ("coef_normal", "ssvs_u", "bvs_u1", "bvs_u2", "sigma_v_gamma", "a_init_normal"):
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


# ---------------------------------------------------------------- oracle runs spread over the host cores (large tests)
def _oracle_chain_job(args):
    obj, inj, seed, literal, table = args
    src = VariateSource(injected=inj) if inj is not None else VariateSource(rng=np.random.default_rng(seed))
    if inj is not None:
        return oracle_stack(obj, [inj], literal=literal, table=table)
    fn = osamp.bvartvpalg if obj["model"].get("tvp") else osamp.bvaralg
    return fn(obj, src, literal=literal, table=table)["posteriors"]


def oracle_stack_parallel(obj, per_chain, literal=False, table="ksc1998", workers=None):
    """oracle_stack with one process per chain (the oracle is single-threaded NumPy); same result, chain order kept."""
    import os
    from concurrent.futures import ProcessPoolExecutor
    workers = workers or min(len(per_chain), os.cpu_count() or 1)
    with ProcessPoolExecutor(max_workers=workers) as ex:
        parts = list(ex.map(_oracle_chain_job, [(obj, inj, 0, literal, table) for inj in per_chain]))
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}


def oracle_sampler_parallel(obj, seeds, literal=False, table="ksc1998", workers=None):
    """The oracle SAMPLER (its own NumPy generator) for a list of seeds, one process per chain; list of `posteriors`."""
    import os
    from concurrent.futures import ProcessPoolExecutor
    workers = workers or min(len(seeds), os.cpu_count() or 1)
    with ProcessPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(_oracle_chain_job, [(obj, None, s, literal, table) for s in seeds]))
