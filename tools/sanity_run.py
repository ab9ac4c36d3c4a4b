#!/usr/bin/env python3
"""Tiny invocation of every kernel family through the public API -- meant to run under compute-sanitizer
(memcheck / racecheck / synccheck) on the GPU box:  compute-sanitizer --tool racecheck python tools/sanity_run.py"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import helpers  # noqa: E402
import bvartools_b200 as bt  # noqa: E402
from bvartools_b200 import ChainSampler  # noqa: E402

which = sys.argv[1:] or ["kron", "general", "cta", "tvp", "large", "loglik", "blocks"]


def run(obj, n_chains, **kw):
    s = ChainSampler(obj, n_chains=n_chains, seed=3, **kw)
    s.run(0)
    s.sync()
    name = s.kernel_name
    tot = s.loglik_total()[0] if not obj["model"].get("tvp") and not obj["model"].get("sv") else s.loglik()[0].sum()
    s.close()
    return name, tot


if "kron" in which:
    print("kron   ", run(helpers.model_c1(iterations=6, burnin=2), 5))
if "general" in which:
    print("general", run(helpers.model_c1(iterations=4, burnin=1, coef=dict(v_i=1, v_i_det=0.1), sigma=dict(df="k", scale=1)), 5))
    print("general", run(helpers.model_c3(K=3, p=2, nobs=60, iterations=4, burnin=1), 3))
    print("general", run(helpers.model_sv(nobs=40, iterations=3, burnin=1), 3))
if "cta" in which:
    print("cta    ", run(helpers.model_c3(iterations=2, burnin=1), 2))
    print("cta    ", run(helpers.model_c3(K=3, p=2, nobs=60, iterations=3, burnin=1), 2, threads_per_chain=128))
if "tvp" in which:
    print("tvp    ", run(helpers.model_sv(tvp=True, nobs=25, iterations=2, burnin=1), 2))
    print("tvp    ", run(helpers.model_sv(tvp=True, nobs=25, iterations=2, burnin=1), 2, threads_per_chain=8))
if "large" in which:
    y = helpers.synth_var(6, 4, 120, 77)
    m = helpers.gen_var(y, p=4, deterministic="const", iterations=2, burnin=1)
    obj = helpers.add_priors(m, coef=dict(minnesota=dict(kappa0=2, kappa1=0.5, kappa3=5)), sigma=dict(df="k", scale=1))
    print("large  ", run(obj, 2, threads_per_chain=-1))
if "loglik" in which:
    u = np.random.default_rng(0).standard_normal((3, 40))
    print("loglik ", bt.loglik_normal(u, np.eye(3)).sum(), bt.loglik_normal(u, np.vstack([np.eye(3)] * 40)).sum())
if "blocks" in which:
    rng = np.random.default_rng(1)
    y = rng.standard_normal((50, 2))
    h = np.zeros((50, 2))
    print("blocks ", bt.stochvol_ksc1998(y, h, np.array([0.05, 0.05]), np.zeros(2), np.full(2, 1e-4), seed=1).shape)
print("done")
