"""ctypes mirror of include/bvargpu.h and the list <-> POD glue.

This is the Python twin of the Rcpp glue shown in INTEGRATION.md: it unpacks the `bvarmodel` nested list by name
exactly the way the reference samplers do (src/bvaralg.cpp:9-249, src/bvartvpalg.cpp:11-279) and fills the flat
`bvg_spec` the C ABI takes.  No numerics happen here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

ABI_VERSION = 2
SIGMA_WISHART, SIGMA_GAMMA, SIGMA_SV = 0, 1, 2
VARSEL_NONE, VARSEL_SSVS, VARSEL_BVS = 0, 1, 2
RESID_AUTO, RESID_DIRECT, RESID_MOMENT = 0, 1, 2
MIX_KSC1998, MIX_OCSN2007 = 7, 10

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

INJECT_FIELDS = ["coef_normal", "ssvs_u", "bvs_u1", "bvs_u2", "sv_u", "sv_normal", "sigma_h_gamma",
                 "h_init_normal", "omega_gamma", "wishart_chi2", "wishart_normal", "state_normal",
                 "sigma_v_gamma", "a_init_normal", "psi_normal", "psi_u1", "psi_u2", "psi_sigma_v_gamma",
                 "psi_init_normal"]


class BvgInject(C.Structure):
    _fields_ = [(f, _dp) for f in INJECT_FIELDS]


class BvgSpec(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("k", C.c_int32), ("T", C.c_int32), ("M", C.c_int32), ("tvp", C.c_int32),
        ("iterations", C.c_int32), ("burnin", C.c_int32),
        ("n_chains", C.c_int64), ("chain_offset", C.c_int64), ("seed", C.c_uint64),
        ("Y", _dp), ("X", _dp),
        ("mu", _dp), ("Vi", _dp), ("vi_dense", C.c_int32), ("varsel", C.c_int32),
        ("tau0", _dp), ("tau1", _dp), ("inprior", _dp), ("include", _ip), ("n_include", C.c_int32),
        ("sigma_type", C.c_int32), ("wishart_df", C.c_double), ("wishart_scale", _dp),
        ("sigma_shape", _dp), ("sigma_rate", _dp), ("sv_mu", _dp), ("sv_vi", _dp), ("sv_sigma_h", _dp),
        ("sv_constant", _dp), ("sv_h", _dp), ("sv_mixture", C.c_int32), ("sigma_i", _dp),
        ("tvp_shape", _dp), ("tvp_rate", _dp), ("a_init", _dp), ("psi_mu", _dp), ("psi_vi", _dp),
        ("psi_varsel", C.c_int32), ("psi_tau0", _dp), ("psi_tau1", _dp), ("psi_inprior", _dp), ("psi_include", _ip),
        ("psi_n_include", C.c_int32), ("psi_shape", _dp), ("psi_rate", _dp), ("psi_init", _dp),
        ("structural", C.c_int32),
        ("resid_mode", C.c_int32), ("threads_per_chain", C.c_int32), ("sweeps_per_launch", C.c_int32),
        ("n_gpus", C.c_int32),
        ("inject", C.POINTER(BvgInject)),
    ]


class BvgOut(C.Structure):
    _fields_ = [("coef", _dp), ("sigma", _dp), ("lambda_", _dp), ("sigma_h", _dp), ("sigma_v", _dp),
                ("status", _ip)]


def _f(a, order="F"):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, order=order))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


class Spec:
    """Owns the NumPy buffers a BvgSpec points to."""

    def __init__(self):
        self.c = BvgSpec()
        self.keep = {}
        self.inject_struct = None

    def set(self, name, arr):
        if arr is None:
            setattr(self.c, name, None)
            return
        self.keep[name] = arr
        if arr.dtype == np.int32:
            setattr(self.c, name, arr.ctypes.data_as(_ip))
        else:
            setattr(self.c, name, arr.ctypes.data_as(_dp))

    @property
    def n(self):
        """coefficient rows per draw: k M, plus the k (k - 1) / 2 structural (A0) coefficients"""
        k = self.c.k
        return k * self.c.M + (k * (k - 1) // 2 if self.c.structural else 0)

    def out_rows(self):
        """rows per retained draw of (coef, sigma, lambda, sigma_h, sigma_v) -- mirrors bvg_out_rows()."""
        k, T, n = self.c.k, self.c.T, self.n
        sv = self.c.sigma_type == SIGMA_SV
        tv_sigma = sv or (bool(self.c.tvp) and bool(self.c.psi_mu))      # TVP + covariance block: Sigma_t per period
        return {
            "coef": n * T if self.c.tvp else n,
            "sigma": k * k * T if tv_sigma else k * k,
            "lambda": (n if self.c.varsel != VARSEL_NONE else 0) +
                      (k * (k - 1) // 2 if self.c.psi_varsel != VARSEL_NONE else 0),
            "sigma_h": k if sv else 0,
            "sigma_v": n if self.c.tvp else 0,
        }


def spec_from_model(obj, n_chains=1, seed=0, chain_offset=0, device=-1, resid_mode=RESID_AUTO,
                    threads_per_chain=0, sweeps_per_launch=0, mixture="ksc1998", inject=None, n_gpus=0) -> Spec:
    """bvarmodel list -> bvg_spec (the unpacking of src/bvaralg.cpp:9-249 / src/bvartvpalg.cpp:11-279)."""
    data, model, priors, initial = obj["data"], obj["model"], obj["priors"], obj["initial"]
    structural = bool(model.get("structural"))
    Y = np.asarray(data["Y"], dtype=np.float64)          # T x k
    T, k = Y.shape
    Z = data.get("Z")
    M = 0 if Z is None else np.asarray(Z).shape[1]
    n = k * M + (k * (k - 1) // 2 if structural and k > 1 else 0)
    sp = Spec()
    c = sp.c
    c.structural = 1 if structural and k > 1 else 0
    c.abi_version, c.device = ABI_VERSION, device
    c.k, c.T, c.M = k, T, M
    c.tvp = 1 if model.get("tvp") else 0
    c.iterations, c.burnin = int(model["iterations"]), int(model["burnin"])
    c.n_chains, c.chain_offset, c.seed = int(n_chains), int(chain_offset), int(seed) & (2 ** 64 - 1)
    sp.set("Y", _f(Y.T))                                  # k x T col-major
    sp.set("X", None if M == 0 else _f(np.asarray(Z, dtype=np.float64).T))
    c.varsel = VARSEL_NONE
    if n > 0:
        pc = priors["coefficients"]
        sp.set("mu", _f(pc["mu"]))
        vi = np.asarray(pc["v_i"], dtype=np.float64)
        if vi.ndim == 2 and vi.shape == (n, n):
            if np.count_nonzero(vi - np.diag(np.diag(vi))) == 0:
                sp.set("Vi", _f(np.diag(vi)))
                c.vi_dense = 0
            else:
                sp.set("Vi", _f(vi))
                c.vi_dense = 1
        else:
            sp.set("Vi", _f(vi))
            c.vi_dense = 0
        sel = None
        if "ssvs" in pc:
            c.varsel = VARSEL_SSVS
            sel = pc["ssvs"]
            sp.set("tau0", _f(sel["tau0"]))
            sp.set("tau1", _f(sel["tau1"]))
        if "bvs" in pc:
            c.varsel = VARSEL_BVS
            sel = pc["bvs"]
        if sel is not None:
            sp.set("inprior", _f(sel["inprior"]))
            inc = np.ascontiguousarray(np.asarray(sel["include"], dtype=np.int64).reshape(-1) - 1, dtype=np.int32)
            sp.set("include", inc)
            c.n_include = inc.shape[0]
        if c.tvp:
            sp.set("tvp_shape", _f(pc["shape"]))
            sp.set("tvp_rate", _f(pc["rate"]))
            sp.set("a_init", _f(initial["coefficients"]["draw"]))
    ps = priors["sigma"]
    if model.get("sv"):
        c.sigma_type = SIGMA_SV
        sp.set("sv_mu", _f(ps["mu"]))
        sp.set("sv_vi", _f(ps["v_i"]))
        sp.set("sigma_shape", _f(ps["shape"]))
        sp.set("sigma_rate", _f(ps["rate"]))
        sp.set("sv_h", _f(initial["sigma"]["h"]))          # T x k col-major
        sp.set("sv_sigma_h", _f(initial["sigma"]["sigma_h"]))
        sp.set("sv_constant", _f(initial["sigma"]["constant"]))
        c.sv_mixture = MIX_OCSN2007 if mixture == "ocsn2007" else MIX_KSC1998
    else:
        sp.set("sigma_i", _f(initial["sigma"]["sigma_i"]))
        if "df" in ps:
            c.sigma_type = SIGMA_WISHART
            c.wishart_df = float(ps["df"])
            sp.set("wishart_scale", _f(ps["scale"]))
        if "shape" in ps:
            c.sigma_type = SIGMA_GAMMA
            sp.set("sigma_shape", _f(ps["shape"]))
            sp.set("sigma_rate", _f(ps["rate"]))
    if "psi" in priors:                                   # covariance block (bvaralg.cpp:157-162)
        sp.set("psi_mu", _f(priors["psi"]["mu"]))
        sp.set("psi_vi", _f(np.asarray(priors["psi"]["v_i"], dtype=np.float64)))
        if model.get("tvp"):                              # bvartvpalg.cpp:147-150,169
            sp.set("psi_shape", _f(priors["psi"]["shape"]))
            sp.set("psi_rate", _f(priors["psi"]["rate"]))
            sp.set("psi_init", _f(initial["psi"]["draw"]))
        psel = None
        if "ssvs" in priors["psi"]:
            c.psi_varsel = VARSEL_SSVS
            psel = priors["psi"]["ssvs"]
            sp.set("psi_tau0", _f(psel["tau0"]))
            sp.set("psi_tau1", _f(psel["tau1"]))
        if "bvs" in priors["psi"]:
            c.psi_varsel = VARSEL_BVS
            psel = priors["psi"]["bvs"]
        if psel is not None:
            sp.set("psi_inprior", _f(psel["inprior"]))
            pinc = np.ascontiguousarray(np.asarray(psel["include"], dtype=np.int64).reshape(-1) - 1, dtype=np.int32)
            sp.set("psi_include", pinc)
            c.psi_n_include = pinc.shape[0]
    c.resid_mode, c.threads_per_chain, c.sweeps_per_launch = resid_mode, threads_per_chain, sweeps_per_launch
    c.n_gpus = int(n_gpus)
    if inject:
        st = BvgInject()
        for name in INJECT_FIELDS:
            if name in inject and inject[name] is not None:
                arr = np.ascontiguousarray(inject[name], dtype=np.float64)
                sp.keep["inj_" + name] = arr
                setattr(st, name, arr.ctypes.data_as(_dp))
        sp.inject_struct = st
        c.inject = C.pointer(st)
    return sp


class OutBuffers:
    """Caller-allocated outputs of one call: arrays shaped [chain][iter][rows] (C order)."""

    def __init__(self, spec: Spec, alloc=None):
        rows = spec.out_rows()
        nc, it = spec.c.n_chains, spec.c.iterations
        self.c = BvgOut()
        self.arrays = {}
        for name, field in (("coef", "coef"), ("sigma", "sigma"), ("lambda", "lambda_"), ("sigma_h", "sigma_h"),
                            ("sigma_v", "sigma_v")):
            r = rows[name]
            if r == 0:
                continue
            shape = (nc, it, r)
            arr = alloc(shape) if alloc is not None else np.empty(shape, dtype=np.float64)
            self.arrays[name] = arr
            setattr(self.c, field, arr.ctypes.data_as(_dp))
        self.status = np.zeros(nc, dtype=np.int32)
        self.c.status = self.status.ctypes.data_as(_ip)
