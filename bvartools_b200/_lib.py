"""Loader of libbvargpu.so (the CUDA product library).  There is no fallback: if the library has not been built
or no CUDA device is present the compute entry points raise."""
from __future__ import annotations

import ctypes as C
import pathlib

from . import _abi

_HERE = pathlib.Path(__file__).resolve().parent
import os as _os
# BVG_LIB: alternative build of the same library (A/B timing experiments)
LIB_PATH = pathlib.Path(_os.environ["BVG_LIB"]) if _os.environ.get("BVG_LIB") else _HERE / "csrc" / "libbvargpu.so"
_lib = None


class BvgError(RuntimeError):
    """Non-zero return code of a libbvargpu entry point (the glue's Rcpp::stop analogue)."""

    def __init__(self, code, message):
        super().__init__(f"libbvargpu error {code}: {message}")
        self.code = code
        self.message = message


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C bvartools_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
    sp, op = C.POINTER(_abi.BvgSpec), C.POINTER(_abi.BvgOut)
    L.bvg_version.restype = C.c_char_p
    L.bvg_last_error.restype = C.c_char_p
    L.bvg_device_count.restype = C.c_int
    L.bvg_host_alloc.restype = vp
    L.bvg_host_alloc.argtypes = [C.c_size_t]
    L.bvg_host_free.argtypes = [vp]
    L.bvg_out_rows.argtypes = [sp] + [C.POINTER(C.c_int64)] * 5
    for fn in ("bvg_bvaralg", "bvg_bvartvpalg"):
        getattr(L, fn).argtypes = [sp, op]
        getattr(L, fn).restype = C.c_int
    L.bvg_sampler_create.argtypes = [sp, C.POINTER(vp)]
    L.bvg_sampler_reset.argtypes = [vp]
    L.bvg_sampler_run.argtypes = [vp, vp]
    L.bvg_sampler_run_to_host.argtypes = [vp, op, vp]
    L.bvg_sampler_fetch.argtypes = [vp, op]
    L.bvg_sampler_device_out.argtypes = [vp, op]
    L.bvg_sampler_sync.argtypes = [vp]
    L.bvg_sampler_last_kernel_ms.argtypes = [vp]
    L.bvg_sampler_last_kernel_ms.restype = C.c_double
    L.bvg_sampler_launch_count.argtypes = [vp]
    L.bvg_sampler_launch_count.restype = C.c_int64
    L.bvg_sampler_kernel_name.argtypes = [vp]
    L.bvg_sampler_kernel_name.restype = C.c_char_p
    L.bvg_sampler_destroy.argtypes = [vp]
    L.bvg_sampler_destroy.restype = None
    L.bvg_moments.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp]
    L.bvg_loglik_draws.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp]
    L.bvg_sampler_loglik.argtypes = [vp, dp, vp]
    L.bvg_sampler_loglik_total.argtypes = [vp, dp, vp]
    L.bvg_sampler_loglik_total_device.argtypes = [vp, vp, vp]
    L.bvg_loglik_total_host.argtypes = [dp, dp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, dp, dp, dp]
    L.bvg_loglik_draws_host.argtypes = [dp, dp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp, dp, dp]
    L.bvg_loglik_normal.argtypes = [dp, dp, C.c_int32, C.c_int32, C.c_int32, C.c_int64, dp]
    L.bvg_summary_draws.argtypes = [vp, C.c_int64, C.c_int64, dp, C.c_int32, dp, dp, dp, vp]
    L.bvg_sampler_summary.argtypes = [vp, C.c_int32, dp, C.c_int32, dp, dp, dp, vp]
    L.bvg_tsse_draws.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, dp, vp]
    L.bvg_sampler_tsse.argtypes = [vp, C.c_int32, dp, vp]
    L.bvg_irf_draws.argtypes = [vp, C.c_int64, vp, C.c_int64, C.c_int64] + [C.c_int32] * 7 + [C.c_double, C.c_int32, vp, vp]
    L.bvg_irf_draws_a0.argtypes = [vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, C.c_int32, C.c_int64] + [C.c_int32] * 7 + [C.c_double, C.c_int32, vp, vp]
    L.bvg_sampler_irf.argtypes = [vp] + [C.c_int32] * 6 + [C.c_double, C.c_int32, C.c_int32, dp, C.c_int32, dp, dp, dp, vp]
    L.bvg_fevd_draws.argtypes = [vp, C.c_int64, vp, C.c_int64, C.c_int64] + [C.c_int32] * 6 + [dp, vp]
    L.bvg_fevd_draws_a0.argtypes = [vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, C.c_int32, C.c_int64] + [C.c_int32] * 6 + [dp, vp]
    L.bvg_sampler_fevd.argtypes = [vp] + [C.c_int32] * 6 + [dp, vp]
    L.bvg_forecast_draws.argtypes = [vp, C.c_int64, vp, C.c_int64, C.c_int64] + [C.c_int32] * 4 + [dp, vp, C.c_uint64, vp, vp]
    L.bvg_forecast_draws_a0.argtypes = [vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, C.c_int32, C.c_int64] + [C.c_int32] * 4 + [dp, vp, C.c_uint64, vp, vp]
    L.bvg_sampler_forecast.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, dp, dp, C.c_uint64, dp, C.c_int32, dp, dp, dp, vp]
    L.bvg_fp64_peak.argtypes = [dp, vp]
    L.bvg_dmma_peak.argtypes = [dp, vp]
    L.bvg_post_normal.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, C.c_uint64, C.c_int, dp]
    L.bvg_post_normal_sur.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, dp, dp, dp, C.c_int, dp, dp, dp,
                                      C.c_uint64, C.c_int, dp]
    L.bvg_ssvs.argtypes = [C.c_int64, C.c_int, dp, dp, dp, dp, ip, C.c_int, dp, C.c_uint64, dp, dp]
    L.bvg_bvs.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, dp, dp, dp, C.c_int, dp, dp, C.c_int, dp, ip,
                          C.c_int, dp, dp, C.c_uint64, dp]
    L.bvg_stochvol.argtypes = [C.c_int64, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_int, dp, dp, C.c_uint64, dp, ip]
    L.bvg_kalman_dk.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_int, dp, dp, dp,
                                C.c_uint64, dp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise BvgError(rc, lib().bvg_last_error().decode("utf-8", "replace"))


def require_device():
    if lib().bvg_device_count() < 1:
        raise BvgError(-50, "no CUDA device available: libbvargpu has no CPU fallback")


EXPORTED = ["bvg_version", "bvg_last_error", "bvg_device_count", "bvg_host_alloc", "bvg_host_free", "bvg_out_rows",
            "bvg_bvaralg", "bvg_bvartvpalg", "bvg_sampler_create", "bvg_sampler_reset", "bvg_sampler_run",
            "bvg_sampler_run_to_host", "bvg_sampler_fetch", "bvg_sampler_device_out", "bvg_sampler_sync",
            "bvg_sampler_last_kernel_ms", "bvg_sampler_launch_count", "bvg_sampler_kernel_name", "bvg_sampler_destroy",
            "bvg_set_interrupt_poll", "bvg_moments", "bvg_loglik_draws", "bvg_loglik_draws_host", "bvg_sampler_loglik", "bvg_sampler_loglik_total", "bvg_sampler_loglik_total_device", "bvg_loglik_total_host", "bvg_loglik_normal", "bvg_summary_draws", "bvg_sampler_summary", "bvg_tsse_draws", "bvg_sampler_tsse", "bvg_irf_draws", "bvg_irf_draws_a0", "bvg_sampler_irf", "bvg_fevd_draws", "bvg_fevd_draws_a0", "bvg_sampler_fevd", "bvg_forecast_draws", "bvg_forecast_draws_a0", "bvg_sampler_forecast", "bvg_fp64_peak", "bvg_dmma_peak", "bvg_post_normal", "bvg_post_normal_sur",
            "bvg_ssvs", "bvg_bvs", "bvg_stochvol", "bvg_kalman_dk"]
