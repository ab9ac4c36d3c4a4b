#!/usr/bin/env python3
"""Time bvg_sampler_loglik on the c2-A draws (4096 chains x 5000 retained) -- run plain or under
`ncu --metrics gpu__time_duration.sum` to split it into its kernels."""
import pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers
from bvartools_b200 import ChainSampler
chains, it = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 5000
obj = helpers.model_c1(iterations=it, burnin=100)
s = ChainSampler(obj, n_chains=chains, seed=1, device=0)
s.run(0); s.sync()
for i in range(3):
    t0 = time.perf_counter(); ll, n = s.loglik(); dt = time.perf_counter() - t0
    print(f"loglik: {dt * 1e3:.2f} ms for {n} draws -> {n / dt / 1e6:.1f} M draws/s, {240 * n / dt / 1e9:.1f} GB/s, LL {ll.sum() / n:.6f}")
for i in range(3):
    t0 = time.perf_counter(); tot, n = s.loglik_total(); dt = time.perf_counter() - t0
    print(f"loglik_total: {dt * 1e3:.2f} ms for {n} draws -> {n / dt / 1e6:.1f} M draws/s, {240 * n / dt / 1e9:.1f} GB/s, LL {tot / n:.6f}")

for which in ("coef", "sigma"):
    for i in range(2):
        t0 = time.perf_counter(); sm = s.summary(which); dt = time.perf_counter() - t0
    rows = sm["mean"].size
    print(f"summary({which}): {dt * 1e3:.2f} ms for {sm['n']} draws x {rows} rows; 10 streaming passes -> {10 * 8 * rows * sm['n'] / dt / 1e9:.1f} GB/s; median[0] {sm['quantiles'][1][0]:.6f}")
for typ in ("feir", "oir"):
    for i in range(2):
        t0 = time.perf_counter(); b = s.irf(0, 1, n_ahead=20, type=typ, shock="sd" if typ == "oir" else 1); dt = time.perf_counter() - t0
    print(f"irf({typ}, h=20): {dt * 1e3:.2f} ms for {chains * it} draws (kernel + 3-quantile bands over 21 horizons); median h=1: {b[1, 1]:.6f}")
