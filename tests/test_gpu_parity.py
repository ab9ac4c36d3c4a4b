"""GPU parity tests proper: the CUDA path, called through the C ABI (libbvargpu.so), against the oracle on the
same injected variates (tier 1: continuous <= 1e-10 relative, indicators bit-exact), plus size-independent
properties at larger sizes and the statistical tier (tier 2)."""
import ctypes as C
import json
import pathlib

import numpy as np
import pytest

import helpers
from test_kernel_logic_cpu import CASES

pytestmark = pytest.mark.gpu
TOL = 1e-10
GOLD = pathlib.Path(__file__).resolve().parent / "golden"


def _run(gpu, obj, n_chains, stacked=None, seed=0, **kw):
    fn = gpu.bvg_bvartvpalg if obj["model"].get("tvp") else gpu.bvg_bvaralg
    rc, out = helpers.run_abi(fn, obj, n_chains, stacked, seed=seed, **kw)
    assert rc == 0, gpu.bvg_last_error()
    return out


@pytest.mark.parametrize("tpc", [0, 8, 32, 128])
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_oracle_injected(gpu, name, tpc):
    """Every sampler variant x lanes-per-chain layout (sub-warp tile, full warp, CTA per chain)."""
    builder, kw = CASES[name]
    obj = builder()
    table = kw.get("mixture", "ksc1998")
    n_chains = 5                                        # ragged: not a multiple of chains per CTA
    per, stacked = helpers.make_variates(obj, n_chains, seed=11)
    out = _run(gpu, obj, n_chains, stacked, threads_per_chain=tpc, **kw)
    want = helpers.oracle_stack(obj, per, table=table)
    assert set(want) == set(out.arrays)
    for key in want:
        assert helpers.relerr(out.arrays[key], want[key]) < TOL, (name, tpc, key)
    if "lambda" in want:
        assert np.array_equal(out.arrays["lambda"], want["lambda"])
    assert not out.status.any()


def test_c3_full_size_ssvs_injected(gpu):
    """Config c3 at full width (K=6, p=4, n=150, T=191; CTA per chain), few sweeps, vs the literal dense oracle."""
    obj = helpers.model_c3(iterations=6, burnin=2)
    per, stacked = helpers.make_variates(obj, 3, seed=5)
    out = _run(gpu, obj, 3, stacked)
    want = helpers.oracle_stack(obj, per, literal=True)
    for key in want:
        assert helpers.relerr(out.arrays[key], want[key]) < TOL, key
    assert np.array_equal(out.arrays["lambda"], want["lambda"])


def test_c4_shape_tvp_sv_injected(gpu):
    """Config c4 shape (TVP-VAR(2), K=3, KSC SV) at T=60 against the structured oracle (the literal dense
    (nT)^2 path is checked against the structured one in test_oracle.py)."""
    obj = helpers.model_sv(tvp=True, nobs=62, iterations=5, burnin=2)
    per, stacked = helpers.make_variates(obj, 3, seed=8)
    out = _run(gpu, obj, 3, stacked)
    want = helpers.oracle_stack(obj, per, literal=False)
    for key in want:
        assert helpers.relerr(out.arrays[key], want[key]) < 1e-8, key      # state dim nT = 1260


def test_c4_full_size_tvp_sv_injected(gpu):
    """Config c4 at BASELINE size (T = 200, n = 21, state dimension nT = 4200; bvartvpalg.cpp:290-297): 2 chains x 3
    sweeps against the structured oracle (== the literal dense path at this size: test_oracle.py
    ::test_literal_equals_structured_c4_full_size)."""
    obj = helpers.model_sv(tvp=True, nobs=202, iterations=2, burnin=1)
    per, stacked = helpers.make_variates(obj, 2, seed=8)
    out = _run(gpu, obj, 2, stacked)
    want = helpers.oracle_stack(obj, per, literal=False)
    assert out.arrays["coef"].shape == (2, 2, 4200)
    for key in want:
        assert helpers.relerr(out.arrays[key], want[key]) < 1e-8, key
    assert not out.status.any()


@pytest.mark.parametrize("case", ["n21", "n150", "n189"])
def test_large_path_small_models_injected(gpu, case):
    """The large-VAR path (per-chain matrices in HBM, blocked DMMA Cholesky) forced on models that also fit on chip
    (threads_per_chain = -1), against the literal dense oracle: 1, 3 (ragged: 150 = 2 * 64 + 22) and 3 tiles."""
    if case == "n21":
        obj = helpers.model_c1(iterations=12, burnin=3, coef=dict(v_i=1, v_i_det=0.1), sigma=dict(df="k", scale=1))
    else:
        K, p = (6, 4) if case == "n150" else (7, 3)
        y = helpers.synth_var(K, p, 150, 77)
        m = helpers.gen_var(y, p=p, deterministic="const", iterations=6, burnin=2)
        K = y.shape[1]
        obj = helpers.add_priors(m, coef=dict(minnesota=dict(kappa0=2, kappa1=0.5, kappa3=5)), sigma=dict(df="k", scale=1))
    per, stacked = helpers.make_variates(obj, 3, seed=21)
    out = _run(gpu, obj, 3, stacked, threads_per_chain=-1)
    want = helpers.oracle_stack(obj, per, literal=True)
    for key in want:   # norm-wise per draw (the literal and the structured oracle themselves agree to ~5e-12 norm-wise,
        # cond(XX') ~ 2.5e5 here; the element-wise metric is dominated by coefficients ~1e-4 of the largest one)
        err = np.linalg.norm(out.arrays[key] - want[key], axis=2) / np.linalg.norm(want[key], axis=2)
        assert err.max() < TOL, (case, key, err.max())
    assert not out.status.any()
    # Philox: the large path and the on-chip kernel draw the same chain (same sites, different arithmetic order)
    a = _run(gpu, obj, 2, seed=5, threads_per_chain=-1)
    b = _run(gpu, obj, 2, seed=5, resid_mode=1)
    for key in a.arrays:
        np.testing.assert_allclose(a.arrays[key], b.arrays[key], rtol=1e-8, atol=1e-11)


def test_c5_full_size_large_var_injected(gpu):
    """Config c5 at full width (K = 20, p = 4, n = 1620, T = 200, Minnesota prior): 2 chains x 3 sweeps against the
    structured oracle (the literal dense kT x kT path is checked against the structured one in test_oracle.py);
    norm-wise 1e-10 as in SURVEY.md Appendix D."""
    obj = helpers.model_c5(iterations=2, burnin=1)
    per, stacked = helpers.make_variates(obj, 2, seed=3)
    out = _run(gpu, obj, 2, stacked)
    want = helpers.oracle_stack(obj, per, literal=False)
    for key in want:
        err = np.linalg.norm(out.arrays[key] - want[key]) / np.linalg.norm(want[key])
        assert err < TOL, (key, err)
    assert not out.status.any()


def test_chunking_and_layout_invariance_philox(gpu):
    """Same seed => bit-identical draws whatever the launch chunking or lanes per chain (counter-based RNG)."""
    obj = helpers.model_c3(K=3, p=2, nobs=80, iterations=30, burnin=10)
    ref = _run(gpu, obj, 37, seed=42)
    for kw in (dict(sweeps_per_launch=7), dict(threads_per_chain=8), dict(threads_per_chain=64)):
        other = _run(gpu, obj, 37, seed=42, **kw)
        for key in ref.arrays:
            if key == "lambda":
                assert np.array_equal(ref.arrays[key], other.arrays[key]), (kw, key)
            else:   # reductions are ordered differently per layout: equal to rounding, chunking alone is exact
                np.testing.assert_allclose(other.arrays[key], ref.arrays[key], rtol=1e-9, atol=1e-12)
    exact = _run(gpu, obj, 37, seed=42, sweeps_per_launch=7)
    for key in ref.arrays:
        assert np.array_equal(ref.arrays[key], exact.arrays[key]), key


@pytest.mark.parametrize("chains", [5, 37, 300])
def test_kronecker_kernel_chunking_is_bit_identical(gpu, chains):
    """Flat-prior Kronecker kernel (fixed warp roles, three record generations): launches of 7, 33 and all sweeps give the
    same draws bit for bit -- rounds restart at every launch (tail rounds, chunks with repeated lanes), the chain state
    W crosses launches, the Philox sites are sweep-addressed.  300 chains = 3 chains per CTA on a 148-SM part."""
    obj = helpers.model_c1(iterations=70, burnin=21)
    ref = _run(gpu, obj, chains, seed=7)
    for spl in (7, 33):
        other = _run(gpu, obj, chains, seed=7, sweeps_per_launch=spl)
        for key in ref.arrays:
            assert np.array_equal(ref.arrays[key], other.arrays[key]), (chains, spl, key)
    assert not ref.status.any()


@pytest.mark.parametrize("case", ["c3_cta_tiled", "tvp_warp_tiled", "large"])
def test_run_to_run_determinism_of_the_tile_kernels(gpu, case):
    """Same seed, same launch shape => bit-identical draws on repeated runs (compute-sanitizer is not available on the
    GPU pool: a shared-memory race in the DMMA tile kernels would show up here as run-to-run noise)."""
    kw = {}
    if case == "c3_cta_tiled":
        obj, chains = helpers.model_c3(iterations=4, burnin=2), 5
    elif case == "tvp_warp_tiled":
        obj, chains = helpers.model_sv(tvp=True, nobs=40, iterations=3, burnin=1), 9
    else:
        y = helpers.synth_var(6, 4, 150, 77)
        m = helpers.gen_var(y, p=4, deterministic="const", iterations=3, burnin=1)
        obj = helpers.add_priors(m, coef=dict(minnesota=dict(kappa0=2, kappa1=0.5, kappa3=5)), sigma=dict(df="k", scale=1))
        chains, kw = 3, dict(threads_per_chain=-1)
    first = _run(gpu, obj, chains, seed=123, **kw)
    for _ in range(3):
        again = _run(gpu, obj, chains, seed=123, **kw)
        for key in first.arrays:
            assert np.array_equal(first.arrays[key], again.arrays[key]), (case, key)


def test_sharding_invariance(gpu):
    """Chains [offset, offset+m) of a shard equal the same global chains of the full job, bit for bit."""
    obj = helpers.model_c1(iterations=40, burnin=10)
    full = _run(gpu, obj, 96, seed=20240113)
    for off, cnt in ((0, 48), (48, 48), (17, 5)):
        part = _run(gpu, obj, cnt, seed=20240113, chain_offset=off)
        for key in full.arrays:
            assert np.array_equal(full.arrays[key][off:off + cnt], part.arrays[key]), (off, key)


def test_matches_emulated_logic_philox(gpu, emu):
    """Philox path: device transforms (log, cospi, exp) vs the host build of the same headers."""
    obj = helpers.model_c1(iterations=25, burnin=5)
    dev = _run(gpu, obj, 4, seed=3)
    rc, host = helpers.run_abi(emu.emu_bvaralg, obj, 4, None, seed=3)
    assert rc == 0
    for key in dev.arrays:
        np.testing.assert_allclose(dev.arrays[key], host.arrays[key], rtol=1e-7, atol=1e-9)


def _mcse(x, nb=20):
    n = (x.shape[0] // nb) * nb
    bm = x[:n].reshape(nb, -1, *x.shape[1:]).mean(axis=1)
    return bm.std(axis=0, ddof=1) / np.sqrt(nb)


def test_tier2_c1_posterior_vs_analytic_and_readme(gpu):
    """Tier 2, config c1/c2-A: full GPU sampler (Philox).  Posterior mean of a = OLS and E[Sigma] = SSE/63 within
    4 MCSE; README table (README.md:240-299) means within 4 combined MCSE, SD ratio within 5 %, quantiles within 4 MCSE
    of a sample quantile."""
    from scipy import stats
    obj = helpers.model_c1(iterations=5000, burnin=1000)
    out = _run(gpu, obj, 64, seed=1234567)
    coef = out.arrays["coef"]                      # [chain][iter][21]
    sig = out.arrays["sigma"]
    assert np.isfinite(coef).all() and np.isfinite(sig).all() and not out.status.any()
    ols = obj["initial"]["coefficients"]["draw"].reshape(-1)
    chain_means = coef.mean(axis=1)                # independent chains -> exact MCSE across chains
    mc = chain_means.std(axis=0, ddof=1) / np.sqrt(coef.shape[0])
    assert np.all(np.abs(chain_means.mean(axis=0) - ols) < 4.0 * mc)
    Y, Z = obj["data"]["Y"], obj["data"]["Z"]
    U = Y - Z @ np.linalg.lstsq(Z, Y, rcond=None)[0]
    expect = ((U.T @ U + 1e-4 * np.eye(3)) / 63.0).reshape(-1, order="F")
    sm = sig.mean(axis=1)
    smc = sm.std(axis=0, ddof=1) / np.sqrt(sig.shape[0])
    assert np.all(np.abs(sm.mean(axis=0) - expect) < 4.0 * smc)
    table = json.loads((GOLD / "readme_table_c1.json").read_text())
    names = ["invest.01", "income.01", "cons.01", "invest.02", "income.02", "cons.02", "const"]
    flat = coef.reshape(-1, 21)
    for i, var in enumerate(["invest", "income", "cons"]):
        for j, reg in enumerate(names):
            row = table["coef"][var][reg]
            col = j * 3 + i
            assert abs(flat[:, col].mean() - row["Mean"]) < 4.0 * np.hypot(mc[col], row["TSSD"])
            assert abs(flat[:, col].std(ddof=1) / row["SD"] - 1.0) < 0.05
            # README quantiles: 4 Monte-Carlo standard errors of a sample quantile of the README's own run (asymptotic
            # sqrt(p (1 - p)) / phi(z_p) = 1.25 / 2.67 times the SE of the mean, = the table's time-series SD) plus the
            # table's 4-significant-digit rounding; the GPU side (64 x 5000 draws) contributes nothing visible
            for qn, qv in (("q2.5", 0.025), ("q50", 0.5), ("q97.5", 0.975)):
                tol = 4.0 * row["TSSD"] * (1.25 if qv == 0.5 else 2.67) * 1.1 + 5e-4 * abs(row[qn])
                assert abs(np.quantile(flat[:, col], qv) - row[qn]) < tol, (var, reg, qn)
    # (KS of the marginals: against the analytic matrix-t posterior and the oracle's own sampler, see
    #  test_tier2_c1_marginals_vs_analytic_matrix_t)


def test_tier2_ssvs_vs_oracle_sampler(gpu):
    """Tier 2 for SSVS: inclusion frequencies and posterior means of the GPU sampler (Philox) vs the dense C
    restatement of the reference sweep run with its own generator, 96 independent chains each; difference of the
    across-chain means within 4.5 standard errors for every coefficient (42 comparisons)."""
    import os
    from oracle.dense_ref import bvar_dense
    obj = helpers.model_c3(K=3, p=2, nobs=120, iterations=1500, burnin=500)
    dev = _run(gpu, obj, 96, seed=7)
    ref = bvar_dense(obj, n_chains=96, seed=99, n_threads=os.cpu_count() or 1, dense_gemm=False)
    assert 0.05 < dev.arrays["lambda"][:, :, :18].mean() < 0.95
    for name in ("lambda", "coef"):
        d, o = dev.arrays[name].mean(axis=1), ref[name].mean(axis=1)
        se = np.hypot(d.std(axis=0, ddof=1) / np.sqrt(d.shape[0]), o.std(axis=0, ddof=1) / np.sqrt(o.shape[0]))
        assert np.all(np.abs(d.mean(axis=0) - o.mean(axis=0)) < 4.5 * se + 1e-9), name


def test_tier2_covariance_block_vs_oracle_sampler(gpu):
    """Tier 2 for the covariance (psi) block with BVS on the covariances (gamma variances): posterior means of the
    coefficients, of Sigma and the psi inclusion frequencies of the GPU sampler (Philox) vs the oracle sampler with its
    own NumPy generator; across-chain means within 4.5 standard errors."""
    from oracle.samplers import bvaralg
    from oracle.variates import VariateSource
    obj = helpers.model_covar_gamma(K=3, p=1, nobs=90, iterations=500, burnin=300, psi_sel="bvs")
    dev = _run(gpu, obj, 48, seed=21)
    o = {"coef": [], "sigma": [], "lambda": []}
    for c in range(12):
        r = bvaralg(obj, VariateSource(rng=np.random.default_rng(900 + c)), literal=False)["posteriors"]
        o["coef"].append(np.concatenate([r[nm]["coeffs"] for nm in ("a", "c")]).mean(axis=1))
        o["sigma"].append(r["sigma"]["coeffs"].mean(axis=1))
        o["lambda"].append(np.concatenate([np.concatenate([r[nm]["lambda"] for nm in ("a", "c")]),
                                           r["sigma"]["lambda"]]).mean(axis=1))
    assert 0.02 < dev.arrays["lambda"][:, :, -3:].mean() < 0.98
    for name in ("coef", "sigma", "lambda"):
        d, oo = dev.arrays[name].mean(axis=1), np.array(o[name])
        se = np.hypot(d.std(axis=0, ddof=1) / np.sqrt(d.shape[0]), oo.std(axis=0, ddof=1) / np.sqrt(oo.shape[0]))
        assert np.all(np.abs(d.mean(axis=0) - oo.mean(axis=0)) < 4.5 * se + 1e-9), name


def test_tier2_tvp_sv_vs_oracle_sampler(gpu):
    """Tier 2 for the TVP + KSC-SV sampler at a small shape: posterior means of the state path, state variances and
    log-volatility driven Sigma_t vs the oracle sampler, within 4.5 MCSE (many marginals)."""
    obj = helpers.model_sv(tvp=True, nobs=42, iterations=600, burnin=400)
    dev = _run(gpu, obj, 32, seed=11)
    from oracle.samplers import bvartvpalg
    from oracle.variates import VariateSource
    o_coef, o_sv, o_sig = [], [], []
    for c in range(8):
        r = bvartvpalg(obj, VariateSource(rng=np.random.default_rng(500 + c)), literal=False)["posteriors"]
        per = [{"x": 0}]
        stack = helpers.oracle_stack.__wrapped__ if hasattr(helpers.oracle_stack, "__wrapped__") else None
        T = obj["data"]["Y"].shape[0]
        parts = [r[nm] for nm in ("a", "c")]
        it = parts[0]["coeffs"].shape[1]
        cube = np.concatenate([p["coeffs"].reshape(T, -1, it) for p in parts], axis=1).reshape(-1, it)
        o_coef.append(cube.mean(axis=1))
        o_sv.append(np.concatenate([p["sigma"] for p in parts]).mean(axis=1))
        o_sig.append(np.log(r["sigma"]["coeffs"].reshape(T, 9, it)[:, [0, 4, 8], :]).mean(axis=2).reshape(-1))
    T = obj["data"]["Y"].shape[0]
    d_coef = dev.arrays["coef"].mean(axis=1)
    d_sv = dev.arrays["sigma_v"].mean(axis=1)
    d_sig = np.log(dev.arrays["sigma"].reshape(32, -1, T, 9)[:, :, :, [0, 4, 8]]).mean(axis=1).reshape(32, -1)
    for d, o in ((d_coef, np.array(o_coef)), (d_sv, np.array(o_sv)), (d_sig, np.array(o_sig))):
        se = np.hypot(d.std(axis=0, ddof=1) / np.sqrt(d.shape[0]), o.std(axis=0, ddof=1) / np.sqrt(o.shape[0]))
        assert np.all(np.abs(d.mean(axis=0) - o.mean(axis=0)) < 4.5 * se + 1e-9)


def test_handle_api_resident_draws_and_moments(gpu):
    """bvg_sampler_*: draws stay in HBM; bvg_moments reduces them on the device; fetch equals the one-shot call."""
    from bvartools_b200 import ChainSampler, run_chains
    from bvartools_b200._lib import check
    obj = helpers.model_c1(iterations=200, burnin=50)
    s = ChainSampler(obj, n_chains=130, seed=5)
    s.run()
    s.sync()
    assert s.last_kernel_ms > 0 and s.launch_count >= 1
    got = s.fetch()
    one = run_chains(obj, n_chains=130, seed=5)
    assert np.array_equal(got.arrays["coef"], one["coef"]) and np.array_equal(got.arrays["sigma"], one["sigma"])
    import torch
    d = s.device_out()
    sums = torch.zeros(21, dtype=torch.float64, device="cuda")
    sq = torch.zeros(21, dtype=torch.float64, device="cuda")
    check(gpu.bvg_moments(C.cast(d.coef, C.c_void_p), 130, 200, 21, C.c_void_p(sums.data_ptr()),
                          C.c_void_p(sq.data_ptr()), None))
    torch.cuda.synchronize()
    flat = one["coef"].reshape(-1, 21)
    np.testing.assert_allclose(sums.cpu().numpy(), flat.sum(axis=0), rtol=1e-10)
    np.testing.assert_allclose(sq.cpu().numpy(), (flat ** 2).sum(axis=0), rtol=1e-10)
    s.reset()
    s.run()
    again = s.fetch()
    assert np.array_equal(again.arrays["coef"], one["coef"])
    s.close()


def test_failure_paths(gpu):
    from bvartools_b200 import run_chains
    from bvartools_b200._lib import BvgError
    obj = helpers.model_c1(iterations=5, burnin=1)
    bad = json.loads(json.dumps({}))  # noqa: F841
    obj_nan = helpers.model_c1(iterations=5, burnin=1)
    obj_nan["data"]["Y"] = obj_nan["data"]["Y"].copy()
    obj_nan["data"]["Y"][2, 1] = np.nan
    with pytest.raises(BvgError, match="NAs"):
        run_chains(obj_nan, 2)
    obj_bad = helpers.model_c1(iterations=5, burnin=1, coef=dict(v_i=1, v_i_det=0.1))
    obj_bad["priors"]["coefficients"]["v_i"] = -1e6 * np.eye(21)
    res = run_chains(obj_bad, 3, seed=1)
    assert res["status"].all()                          # flagged, call itself survives
    good = run_chains(obj, 3, seed=1)
    assert not good["status"].any()


def test_reference_shaped_results(gpu):
    """bvaralg / bvartvpalg / bvarpost / draw_posterior return the reference's structures (src/bvaralg.cpp:562-627,
    src/bvartvpalg.cpp:555-626, R/bvarpost.R:111-119)."""
    import bvartools_b200 as b
    obj = helpers.model_c3(K=3, p=2, nobs=80, iterations=20, burnin=5)
    r = b.bvaralg(obj, seed=1)
    post = r["posteriors"]
    assert post["a"]["coeffs"].shape == (18, 20) and post["a"]["lambda"].shape == (18, 20)
    assert post["c"]["coeffs"].shape == (3, 20) and post["b"] is None and post["a0"] is None
    assert post["sigma"]["coeffs"].shape == (9, 20)
    tv = helpers.model_sv(tvp=True, nobs=30, iterations=6, burnin=2)
    rt = b.bvartvpalg(tv, seed=2)["posteriors"]
    T = tv["data"]["Y"].shape[0]
    assert rt["a"]["coeffs"].shape == (18 * T, 6) and rt["a"]["sigma"].shape == (18, 6)
    assert rt["c"]["coeffs"].shape == (3 * T, 6) and rt["sigma"]["coeffs"].shape == (9 * T, 6)
    assert rt["sigma"]["sigma"].shape == (9, 6)
    est = b.draw_posterior([obj, helpers.model_c1(iterations=10, burnin=2)], seed=3)
    assert len(est) == 2 and est[0]["class"] == "bvar" and est[0]["A"].shape == (20, 18)
    assert est[1]["Sigma"].shape == (10, 9) and est[0]["specifications"]["lags"]["p"] == 2


@pytest.mark.gpu
def test_tvp_covariance_block_through_the_mirror(gpu):
    """TVP-VAR with time-varying covariances and SV through draw_posterior: Sigma_t per period, positive definite,
    genuinely non-diagonal, identical for a repeated seed, and usable by the post-estimation calls."""
    import bvartools_b200 as bt
    obj = helpers.model_tvp_covar(K=3, p=1, nobs=40, iterations=30, burnin=10)
    fit = bt.draw_posterior(obj, seed=3)
    assert not fit.get("error"), fit.get("message")
    T, k = obj["data"]["Y"].shape
    assert isinstance(fit["Sigma"], list) and len(fit["Sigma"]) == T and fit["Sigma"][0].shape == (30, k * k)
    S = np.stack(fit["Sigma"], axis=1).reshape(30, T, k, k)
    assert np.allclose(S, np.swapaxes(S, -1, -2), rtol=1e-12, atol=0)
    assert np.all(np.linalg.eigvalsh(S) > 0)
    assert np.abs(S[..., 1, 0]).max() > 0
    again = bt.draw_posterior(obj, seed=3)
    assert np.array_equal(np.stack(again["Sigma"]), np.stack(fit["Sigma"]))
    ll = bt.summary_bvarlist([fit])
    assert np.isfinite(ll[0]["LL"])
    bands = bt.irf(fit, 0, 1, n_ahead=4, type="oir")
    assert bands.shape == (5, 3) and np.isfinite(bands).all()


@pytest.mark.gpu
def test_structural_model_through_the_mirror(gpu):
    """Structural VAR through draw_posterior: A0 draws (identity + strict lower triangle), reproducible, and the
    lower-triangular elements agree with the oracle sampler's posterior means within Monte-Carlo error."""
    import bvartools_b200 as bt
    from oracle import samplers as osamp
    from oracle.variates import VariateSource
    obj = helpers.model_structural(K=3, p=1, nobs=80, iterations=400, burnin=100)
    fit = bt.draw_posterior(obj, seed=21)
    assert not fit.get("error"), fit.get("message")
    k = 3
    A0 = np.asarray(fit["A0"])
    assert A0.shape == (400, k * k)
    M = A0.reshape(400, k, k, order="F")
    assert np.all(M[:, np.arange(k), np.arange(k)] == 1.0) and np.all(np.triu(M, 1) == 0.0)
    assert np.abs(M[:, 1, 0]).max() > 0
    again = bt.draw_posterior(obj, seed=21)
    assert np.array_equal(np.asarray(again["A0"]), A0)
    ref = osamp.bvaralg(obj, VariateSource(rng=np.random.default_rng(5)), literal=True)["posteriors"]["a0"]["coeffs"]
    lower = [i + k * j for j in range(k) for i in range(k) if i > j]
    got_m, ref_m = A0[:, lower].mean(axis=0), ref.mean(axis=1)
    se = np.sqrt(A0[:, lower].var(axis=0, ddof=1) / 400 + ref.var(axis=1, ddof=1) / 400)
    assert np.all(np.abs(got_m - ref_m) < 6.0 * se), (got_m, ref_m, se)


def test_structural_tvp_guard_padded_buffers(gpu):
    """Structural TVP models store k M + k (k - 1) / 2 rows of sigma_v / lambda per draw, not the k (M + k) rows of the
    augmented model the kernels run on: bvg_bvartvpalg must fill exactly the buffers bvg_out_rows announces (guard words
    behind every output stay intact) and run_to_host == fetch == the oracle."""
    from bvartools_b200 import _abi
    obj = helpers.model_structural(tvp=True, sv=True, bvs=True, nobs=24, iterations=6, burnin=2)
    per, stacked = helpers.make_variates(obj, 3, seed=17)
    spec = _abi.spec_from_model(obj, n_chains=3, seed=0, inject=stacked)
    vals = [C.c_int64() for _ in range(5)]
    assert gpu.bvg_out_rows(C.byref(spec.c), *[C.byref(v) for v in vals]) == 0
    rows = spec.out_rows()
    assert [v.value for v in vals] == [rows[nm] for nm in ("coef", "sigma", "lambda", "sigma_h", "sigma_v")]
    GUARD = 64
    sentinel = -7.25e300

    def alloc(shape):
        flat = np.full(int(np.prod(shape)) + GUARD, sentinel)
        alloc.raw.append(flat)
        return flat[:-GUARD].reshape(shape)
    alloc.raw = []
    out = _abi.OutBuffers(spec, alloc=alloc)
    assert gpu.bvg_bvartvpalg(C.byref(spec.c), C.byref(out.c)) == 0, gpu.bvg_last_error()
    for flat in alloc.raw:
        assert np.all(flat[-GUARD:] == sentinel)
        assert not np.any(flat[:-GUARD] == sentinel)
    want = helpers.oracle_stack(obj, per)
    assert set(want) == set(out.arrays)
    for key in want:
        assert helpers.relerr(out.arrays[key], want[key]) < TOL, key
    assert np.array_equal(out.arrays["lambda"], want["lambda"])
    # handle API: fetch() and the on-device summary of sigma_v use the same row count
    from bvartools_b200 import ChainSampler
    s = ChainSampler(obj, n_chains=3, seed=0, inject=stacked)
    s.run()
    got = s.fetch()
    for key in want:
        assert np.array_equal(got.arrays[key], out.arrays[key]), key
    sm = s.summary("sigma_v", ts_se=False)
    np.testing.assert_allclose(sm["mean"], out.arrays["sigma_v"].reshape(-1, rows["sigma_v"]).mean(axis=0), rtol=1e-12)
    s.close()


@pytest.mark.parametrize("case", ["kron", "ssvs_injected", "tvp_sv"])
def test_library_internal_sharding_is_bit_identical(gpu, case):
    """bvg_spec.n_gpus: the one-shot entry points shard the chains over host threads / devices themselves (shard i on
    device (device + i) mod count -- on a one-GPU box all shards share the device, which exercises the same slicing of
    outputs, injected variates and global chain ids); results equal the single-device call bit for bit."""
    if case == "kron":
        obj, n, kw = helpers.model_c1(iterations=60, burnin=20), 37, {}
        stacked = None
    elif case == "ssvs_injected":
        obj, n, kw = helpers.model_c3(K=3, p=2, nobs=80, iterations=12, burnin=4), 7, {}
        _, stacked = helpers.make_variates(obj, n, seed=3)
    else:
        obj, n, kw = helpers.model_sv(tvp=True, nobs=30, iterations=6, burnin=2), 5, {}
        stacked = None
    one = _run(gpu, obj, n, stacked, seed=99, **kw)
    for g in (2, 3, -1):
        many = _run(gpu, obj, n, stacked, seed=99, n_gpus=g, **kw)
        for key in one.arrays:
            assert np.array_equal(one.arrays[key], many.arrays[key]), (case, g, key)
        assert np.array_equal(one.status, many.status)
    # more shards than chains, and an offset shard of a larger job
    few = _run(gpu, obj, 2, None if stacked is None else {k_: v[:2] for k_, v in stacked.items()}, seed=99, n_gpus=8, **kw)
    for key in one.arrays:
        assert np.array_equal(one.arrays[key][:2], few.arrays[key]), key


def test_interrupt_poll_aborts_the_run(gpu):
    """bvg_set_interrupt_poll (Rcpp::checkUserInterrupt analogue, bvaralg.cpp:294-296): a non-zero return of the callback
    aborts the call with -100, on the single-device path and -- polled from the calling thread -- on the sharded one."""
    from bvartools_b200 import _abi
    obj = helpers.model_c3(iterations=3000, burnin=0)         # ~0.5 s of sweeps: the calling thread polls every 20 ms
    calls = {"n": 0}
    CB = C.CFUNCTYPE(C.c_int)

    def poll():
        calls["n"] += 1
        return 1 if calls["n"] >= 2 else 0
    cb = CB(poll)
    gpu.bvg_set_interrupt_poll.argtypes = [CB]
    gpu.bvg_set_interrupt_poll(cb)
    try:
        for ng in (0, 2):
            calls["n"] = 0
            spec = _abi.spec_from_model(obj, n_chains=296, seed=1, sweeps_per_launch=50, n_gpus=ng)
            out = _abi.OutBuffers(spec)
            rc = gpu.bvg_bvaralg(C.byref(spec.c), C.byref(out.c))
            assert rc == -100, (ng, rc, gpu.bvg_last_error())
            assert b"interrupted" in gpu.bvg_last_error()
            assert calls["n"] >= 2
    finally:
        gpu.bvg_set_interrupt_poll(CB(0))
    ok = _run(gpu, helpers.model_c1(iterations=20, burnin=5), 4, seed=1)
    assert not ok.status.any()


# ------------------------------------------------------------------------------------------- tier 2 at BASELINE shapes
def _across_chain_check(dev_means, ref_means, z=4.5, what=""):
    """|mean_dev - mean_ref| < z * SE, SE from the spread of the per-chain means of both samplers (independent chains)."""
    d, o = np.asarray(dev_means), np.asarray(ref_means)
    se = np.hypot(d.std(axis=0, ddof=1) / np.sqrt(d.shape[0]), o.std(axis=0, ddof=1) / np.sqrt(o.shape[0]))
    bad = np.abs(d.mean(axis=0) - o.mean(axis=0)) > z * se + 1e-12
    assert not bad.any(), (what, int(bad.sum()), np.abs(d.mean(axis=0) - o.mean(axis=0))[bad][:5], se[bad][:5])


def test_tier2_c1_marginals_vs_analytic_matrix_t(gpu):
    """Config c1 / c2-A against the ANALYTIC posterior.  Flat coefficient prior + Wishart prior: integrating a out of the joint
    posterior gives Sigma^-1 ~ W((S0 + SSE)^-1, T + df - M) and A | y matrix-t, i.e. every coefficient is Student-t with
    nu = T + df - M - k + 1 = 65 degrees of freedom, location OLS and scale^2 = (S0 + SSE)_ii (X X')^-1_jj / nu.
    One-sample KS of the (thinned) GPU draws against that t distribution for all 21 coefficients, Bonferroni 1e-3; plus a
    two-sample KS against the draws of the oracle's C restatement run with its own generator."""
    import os
    from scipy import stats
    from oracle.dense_ref import bvar_dense
    obj = helpers.model_c1(iterations=5000, burnin=1000)
    out = _run(gpu, obj, 64, seed=424242)
    coef = out.arrays["coef"][:, ::10, :].reshape(-1, 21)         # thinned: the chain mixes in a few sweeps
    Y, Z = obj["data"]["Y"], obj["data"]["Z"]
    T, k = Y.shape
    M = Z.shape[1]
    B = np.linalg.lstsq(Z, Y, rcond=None)[0]                      # M x k
    S = (Y - Z @ B).T @ (Y - Z @ B) + 1e-4 * np.eye(k)
    XXi = np.linalg.inv(Z.T @ Z)
    nu = T + 1.0 - M - k + 1.0
    pvals = []
    for j in range(M):
        for i in range(k):
            scale = np.sqrt(S[i, i] * XXi[j, j] / nu)
            pvals.append(stats.kstest(coef[:, j * k + i], stats.t(df=nu, loc=B[j, i], scale=scale).cdf).pvalue)
    assert min(pvals) > 1e-3 / 21, sorted(pvals)[:3]
    assert np.median(pvals) > 0.05
    ref = bvar_dense(obj, n_chains=16, seed=777, n_threads=os.cpu_count() or 1, dense_gemm=False)
    rc = ref["coef"][:, ::10, :].reshape(-1, 21)
    rs = ref["sigma"][:, ::10, :].reshape(-1, 9)
    ds = out.arrays["sigma"][:, ::10, :].reshape(-1, 9)
    p2 = [stats.ks_2samp(coef[:, c], rc[:, c]).pvalue for c in range(21)] + \
         [stats.ks_2samp(ds[:, c], rs[:, c]).pvalue for c in (0, 1, 2, 4, 5, 8)]
    assert min(p2) > 1e-3 / len(p2), sorted(p2)[:3]


def test_tier2_c3_full_size_ssvs_vs_c_port(gpu):
    """Tier 2 at the c3 shape (K = 6, p = 4, n = 150, T = 191, SSVS): GPU sampler (Philox) against the structure-exploiting C
    restatement with its own generator; posterior means of all 150 coefficients, their inclusion frequencies and Sigma
    within 4.5 standard errors of the across-chain means."""
    import os
    from oracle.dense_ref import bvar_dense
    obj = helpers.model_c3(iterations=500, burnin=300)
    dev = _run(gpu, obj, 64, seed=31)
    ref = bvar_dense(obj, n_chains=12, seed=5, n_threads=os.cpu_count() or 1, dense_gemm=False)
    assert not dev.status.any()
    assert 0.02 < dev.arrays["lambda"][:, :, :144].mean() < 0.98
    for name in ("coef", "lambda", "sigma"):
        _across_chain_check(dev.arrays[name].mean(axis=1), ref[name].mean(axis=1), what=name)


def test_tier2_c4_full_size_tvp_sv_vs_oracle_sampler(gpu):
    """Tier 2 at the c4 shape (T = 200, n = 21, KSC SV): posterior means of the whole state path (4200 marginals), of the
    state variances and of log Sigma_t of the GPU sampler against the oracle sampler (own NumPy generator, structured path
    == literal dense path: test_oracle.py), 8 oracle chains."""
    obj = helpers.model_sv(tvp=True, nobs=202, iterations=200, burnin=200)
    dev = _run(gpu, obj, 32, seed=77)
    assert not dev.status.any()
    posts = helpers.oracle_sampler_parallel(obj, [900 + c for c in range(8)], literal=False)
    T = obj["data"]["Y"].shape[0]
    o_coef, o_sv, o_sig = [], [], []
    for r in posts:
        parts = [r[nm] for nm in ("a", "c")]
        it = parts[0]["coeffs"].shape[1]
        cube = np.concatenate([p["coeffs"].reshape(T, -1, it) for p in parts], axis=1).reshape(-1, it)
        o_coef.append(cube.mean(axis=1))
        o_sv.append(np.log(np.concatenate([p["sigma"] for p in parts])).mean(axis=1))
        o_sig.append(np.log(r["sigma"]["coeffs"].reshape(T, 9, it)[:, [0, 4, 8], :]).mean(axis=2).reshape(-1))
    d_coef = dev.arrays["coef"].mean(axis=1)
    d_sv = np.log(dev.arrays["sigma_v"]).mean(axis=1)
    d_sig = np.log(dev.arrays["sigma"].reshape(32, -1, T, 9)[:, :, :, [0, 4, 8]]).mean(axis=1).reshape(32, -1)
    _across_chain_check(d_coef, o_coef, z=5.0, what="state path")          # 4200 comparisons
    _across_chain_check(d_sv, o_sv, what="log state variances")
    _across_chain_check(d_sig, o_sig, z=5.0, what="log Sigma_t")


def test_tier2_large_path_vs_c_port(gpu):
    """Tier 2 for the large-VAR path (config c5's kernels: blocked DMMA Cholesky in HBM) at K = 6, p = 4 (n = 150, three
    block columns, Minnesota prior) -- the c5 shape itself is beyond what the CPU restatement samples in test time (one
    dense sweep at n = 1620 takes seconds); c5 at full size is covered by tier 1 (test_c5_full_size_large_var_injected)."""
    import os
    from oracle.dense_ref import bvar_dense
    y = helpers.synth_var(6, 4, 150, 77)
    m = helpers.gen_var(y, p=4, deterministic="const", iterations=400, burnin=200)
    obj = helpers.add_priors(m, coef=dict(minnesota=dict(kappa0=2, kappa1=0.5, kappa3=5)), sigma=dict(df="k", scale=1))
    dev = _run(gpu, obj, 48, seed=13, threads_per_chain=-1)
    ref = bvar_dense(obj, n_chains=12, seed=6, n_threads=os.cpu_count() or 1, dense_gemm=False)
    assert not dev.status.any()
    for name in ("coef", "sigma"):
        _across_chain_check(dev.arrays[name].mean(axis=1), ref[name].mean(axis=1), what=name)


def test_bvs_decisions_match_over_a_million_sites(gpu):
    """the BVS inclusion decisions of the CUDA path (rank-1 likelihood ratio, stale A_G) against the
    reference-ordered evaluation of the oracle on the same injected uniforms / normals: 64 chains x 900 sweeps x 18 positions
    = 1.04e6 decisions, zero mismatches allowed; continuous draws <= 1e-10."""
    obj = helpers.model_c3(K=3, p=2, nobs=80, iterations=900, burnin=0, bvs=True)
    n_chains = 64
    per, stacked = helpers.make_variates(obj, n_chains, seed=2024)
    out = _run(gpu, obj, n_chains, stacked)
    want = helpers.oracle_stack_parallel(obj, per, literal=False)
    n_inc = int(np.asarray(obj["priors"]["coefficients"]["bvs"]["include"]).size)
    assert n_chains * 900 * n_inc >= 1_000_000
    mism = int((out.arrays["lambda"] != want["lambda"]).sum())
    assert mism == 0, mism
    assert 0.05 < want["lambda"][:, :, :n_inc].mean() < 0.95          # the decisions are not trivial
    for key in ("coef", "sigma"):
        assert helpers.relerr(out.arrays[key], want[key]) < TOL, key
