"""ctypes wrapper of oracle/blocks_ref.c -- TEST INFRASTRUCTURE ONLY (the second, plain-C restatement of the SV step, the
BVS step and the TVP precision sampler; see the header of the C file)."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_DIR = pathlib.Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libblocks_ref.so"
_lib = None


def load():
    global _lib
    if _lib is None:
        src = _DIR / "blocks_ref.c"
        if not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
            subprocess.check_call(["make", "-s", "-C", str(_DIR)])
        _lib = C.CDLL(str(_SO))
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _lib.stochvol_ref.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, C.c_int, ip]
        _lib.bvs_ref.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, C.c_int, dp, dp, C.c_int, dp, dp, dp, ip, C.c_int]
        _lib.tvp_state_ref.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp]
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def stochvol(y, h, sigma, h_init, constant, u, eps, table="ksc1998"):
    """Same arguments as oracle.blocks.stochvol_mixture; returns (h_new T x k, s T x k)."""
    y, h = _f(y), _f(h).copy(order="F")
    T, k = y.shape
    sigma, h_init, constant = (np.ascontiguousarray(v, dtype=np.float64) for v in (sigma, h_init, constant))
    u, eps = np.ascontiguousarray(u, dtype=np.float64), np.ascontiguousarray(eps, dtype=np.float64)   # k x T, row-major
    s = np.zeros((T, k), dtype=np.int32, order="F")
    rc = load().stochvol_ref(T, k, _d(y), _d(h), _d(sigma), _d(h_init), _d(constant), _d(u), _d(eps),
                             7 if table == "ksc1998" else 10, s.ctypes.data_as(C.POINTER(C.c_int)))
    if rc:
        raise RuntimeError(f"stochvol_ref failed with status {rc}")
    return np.array(h), np.array(s, dtype=np.int64)


def bvs(y, z, a, lam, sigma_i, prob_prior, u1, u2, include=None):
    """Same arguments as oracle.blocks.bvs; returns the n x n diagonal indicator matrix."""
    y, z = _f(y), _f(z)
    k, T = y.shape
    a = np.asarray(a, dtype=np.float64)
    a = _f(a[:, None] if a.ndim == 1 else a)
    n, na = a.shape
    lamd = np.ascontiguousarray(np.diag(np.asarray(lam, dtype=np.float64))).copy()
    sg = _f(sigma_i)
    prior, u1, u2 = (np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1)) for v in (prob_prior, u1, u2))
    inc = np.arange(n, dtype=np.int32) if include is None else \
        np.ascontiguousarray(np.asarray(include, dtype=np.int64).reshape(-1) - 1, dtype=np.int32)
    load().bvs_ref(k, T, n, _d(y), _d(z), _d(a), na, _d(lamd), _d(sg), sg.shape[0], _d(prior), _d(u1), _d(u2),
                   inc.ctypes.data_as(C.POINTER(C.c_int)), inc.shape[0])
    return np.diag(lamd)


def tvp_state(y, z, sig_stack, sigma_v_i, a_init, eps):
    """a (n T, time-major) = P^-1 rhs + chol(P)^-1 eps of bvartvpalg.cpp:290-297.  y k x T, z kT x n (masked), sig_stack
    kT x k, sigma_v_i / a_init n, eps nT."""
    y, z, sg = _f(y), _f(z), _f(sig_stack)
    k, T = y.shape
    n = z.shape[1]
    sv, a0, e = (np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1)) for v in (sigma_v_i, a_init, eps))
    out = np.empty(n * T)
    rc = load().tvp_state_ref(k, T, n, _d(y), _d(z), _d(sg), _d(sv), _d(a0), _d(e), _d(out))
    if rc:
        raise RuntimeError(f"tvp_state_ref failed with status {rc}")
    return out
