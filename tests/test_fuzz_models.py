"""Randomised differential test: small models drawn at random (k, p, T, deterministic terms, error-variance prior,
variable selection, covariance block, TVP) through the kernel logic against the oracle on injected variates.
`emu` = the group-generic headers compiled with g++ (CPU); `gpu` = the CUDA kernels through the C ABI with a random
lanes-per-chain layout (sub-warp tile, warp, CTA per chain)."""
import numpy as np
import pytest

import helpers


def _random_model(rng):
    k = int(rng.integers(1, 6))
    p = int(rng.integers(0, 4))
    nobs = int(rng.integers(24, 70))
    kind = str(rng.choice(["wishart", "gamma", "sv", "gamma_covar", "sv_covar", "tvp_sv", "tvp_wishart", "tvp_sv_covar",
                           "tvp_gamma_covar", "structural_gamma", "structural_sv"]))
    sel = str(rng.choice(["none", "ssvs", "bvs"]))
    det = str(rng.choice(["const", "none", "both"])) if p > 0 else "const"
    tvp, sv = kind.startswith("tvp"), "sv" in kind
    structural = kind.startswith("structural") and k > 1
    if structural and sel == "ssvs":
        sel = "none"                                       # SSVS is not combined with structural coefficients
    y = helpers.synth_var(k, max(p, 1), nobs, int(rng.integers(1, 1000)))
    m = helpers.gen_var(y, p=p, deterministic=det, iterations=6, burnin=2, tvp=tvp, sv=sv, structural=structural)
    tvp_covar = kind in ("tvp_sv_covar", "tvp_gamma_covar") and k > 1 and k * (k - 1) // 2 <= k * (k * p + (det != "none"))
    if kind in ("gamma", "gamma_covar", "tvp_gamma_covar", "structural_gamma"):
        sigma = dict(shape=3, rate=1e-3, covar=((kind == "gamma_covar" and k > 1) or tvp_covar))
    elif sv:
        sigma = dict(shape=3, rate=1e-4, mu=0, v_i=0.01, sigma_h=0.05, constant=1e-4,
                     covar=((kind == "sv_covar" and k > 1) or tvp_covar))
    else:
        sigma = dict(df="k", scale=1)
    ssvs = dict(inprior=0.5, tau=(0.05, 10.0), exclude_det=bool(rng.integers(0, 2))) if sel == "ssvs" and not sv and not tvp else None
    bvs = dict(inprior=0.5, exclude_det=bool(rng.integers(0, 2))) if sel == "bvs" else None
    if bvs is not None and tvp_covar:
        bvs["covar"] = bool(rng.integers(0, 2))            # BVS on the time-varying covariances as well
    obj = helpers.add_priors(m, coef=dict(v_i=1, v_i_det=0.1, shape=3, rate=1e-4, rate_det=0.01), sigma=sigma, ssvs=ssvs,
                             bvs=bvs)
    return obj, (k, p, nobs, kind, sel, det)


def _check(fn_for, rng, trials, pick_tpc=None):
    done = 0
    for trial in range(trials):
        obj, tag = _random_model(rng)
        tvp = bool(obj["model"].get("tvp"))
        per, stacked = helpers.make_variates(obj, 2, seed=trial)
        kw = {} if pick_tpc is None else dict(threads_per_chain=int(pick_tpc(rng)))
        rc, out = helpers.run_abi(fn_for(tvp), obj, 2, stacked, **kw)
        assert rc == 0, (tag, kw)
        want = helpers.oracle_stack(obj, per, literal=True)
        tol = 1e-8 if tvp else 1e-10
        for key in want:
            assert helpers.relerr(out.arrays[key], want[key]) < tol, (tag, kw, key)
            if key == "lambda":
                assert np.array_equal(out.arrays[key], want[key]), (tag, kw)
        done += 1
    return done


def test_random_models_kernel_logic_cpu(emu):
    rng = np.random.default_rng(20240113)
    assert _check(lambda tvp: emu.emu_bvartvpalg if tvp else emu.emu_bvaralg, rng, 24) == 24


@pytest.mark.gpu
def test_random_models_cuda(gpu):
    rng = np.random.default_rng(7)
    assert _check(lambda tvp: gpu.bvg_bvartvpalg if tvp else gpu.bvg_bvaralg, rng, 30,
                  pick_tpc=lambda r: r.choice([0, 0, 8, 32, 128])) == 30
