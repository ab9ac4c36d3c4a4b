#!/usr/bin/env python3
"""Cycles per phase of the warp-per-chain TVP sweep (chain 0, lane 0).  Needs a measurement build of the library
(tools/build_prof_lib.sh -> build_alt/libbvargpu_prof.so; the product library carries no instrumentation):
    BVG_LIB=build_alt/libbvargpu_prof.so python tools/micro/tvp_prof.py"""
import ctypes as C
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers  # noqa: E402
from bvartools_b200 import ChainSampler  # noqa: E402
from bvartools_b200._lib import lib  # noqa: E402

sweeps = 20
obj = helpers.model_sv(K=3, p=2, nobs=202, iterations=sweeps // 2, burnin=sweeps // 2, tvp=True)
s = ChainSampler(obj, n_chains=1024, seed=1, device=0)
s.run(0)
s.sync()
import torch  # noqa: E402
torch.cuda.synchronize()
out = (C.c_longlong * 16)()
rc = lib().bvg_debug_tvp_prof(out)
T = 200
names = ["build", "factor", "inverse / columns", "publish + stream", "carry (tiled path)", "(unused)", "backward (whole pass)", "normals prologue", "residuals", "sv_step", "sigma_h / h_init / Sigma out", "state variances", "a_init draw", "store", "factor_invert only (tiled path)"]
tot = sum(out[:15])
print("rc", rc, "kernel ms", s.last_kernel_ms)
for i, nm in enumerate(names):
    per_t = out[i] / (sweeps * T)
    print(f"{nm:24s} {per_t:9.0f} cycles per period   {100.0 * out[i] / tot:5.1f} %")
print(f"total {tot / (sweeps * T):.0f} cycles per period, {tot / sweeps / 1e6:.2f} Mcycles per sweep")
