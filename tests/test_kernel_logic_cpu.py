"""CPU tests of the kernel LOGIC: the group-generic headers the CUDA kernels are built from, compiled with g++
as a single-thread group (tests/emu -- test-only, never loaded by the product), against the oracle on injected
variates.  Tier-1 bar (BASELINE.json north_star): continuous draws <= 1e-10 relative, indicators bit-exact."""
import ctypes as C

import numpy as np
import pytest

import helpers

TOL = 1e-10

CASES = {
    "c1_wishart_moment": (lambda: helpers.model_c1(iterations=40, burnin=10), dict(resid_mode=2)),
    "c1_wishart_direct": (lambda: helpers.model_c1(iterations=40, burnin=10), dict(resid_mode=1)),
    "c1_informative_prior": (lambda: helpers.model_c1(iterations=30, burnin=5, coef=dict(v_i=1, v_i_det=0.1),
                                                     sigma=dict(df="k", scale=1)), {}),
    "ssvs_moment": (lambda: helpers.model_c3(K=3, p=2, nobs=80), dict(resid_mode=2)),
    "ssvs_direct": (lambda: helpers.model_c3(K=3, p=2, nobs=80), dict(resid_mode=1)),
    "bvs_moment": (lambda: helpers.model_c3(K=3, p=2, nobs=80, bvs=True), dict(resid_mode=2)),
    "bvs_direct": (lambda: helpers.model_c3(K=3, p=2, nobs=80, bvs=True), dict(resid_mode=1)),
    "gamma": (lambda: helpers.model_c1(iterations=30, burnin=5, sigma=dict(shape=3, rate=1e-4),
                                       coef=dict(v_i=1, v_i_det=0.1)), {}),
    "sv_ksc": (lambda: helpers.model_sv(), {}),
    "sv_ocsn": (lambda: helpers.model_sv(), dict(mixture="ocsn2007")),
    "sv_bvs": (lambda: helpers.model_sv(bvs=True), {}),
    "covar_gamma": (lambda: helpers.model_covar_gamma(), {}),
    "covar_gamma_k4_bvs": (lambda: helpers.model_covar_gamma(K=4, p=1, bvs=True), {}),
    "covar_gamma_psi_ssvs": (lambda: helpers.model_covar_gamma(psi_sel="ssvs"), {}),
    "covar_gamma_psi_bvs": (lambda: helpers.model_covar_gamma(K=4, p=1, psi_sel="bvs"), {}),
    "covar_sv_psi_bvs": (lambda: helpers.model_sv(covar=True, psi_bvs=True), {}),
    "covar_sv": (lambda: helpers.model_sv(covar=True), {}),
    "covar_sv_k2": (lambda: helpers.model_sv(K=2, p=1, covar=True), {}),
    "tvp_sv": (lambda: helpers.model_sv(tvp=True, nobs=30), {}),
    "tvp_sv_bvs": (lambda: helpers.model_sv(tvp=True, nobs=30, bvs=True), {}),
}


def _grid_model(p, K=3, **kw):
    """One member of the p = 0:4 grid of config c2-B (flat prior, df = k) -> Kronecker sweep kernel when p > 0."""
    x, names = helpers.e1_readme_sample()
    ms = helpers.gen_var(x[:, :K], p=range(0, 5), deterministic="const", iterations=25, burnin=5, names=names[:K])
    return helpers.add_priors(ms[p], coef=dict(v_i=0, v_i_det=0), sigma=dict(df="k", scale=1e-4))


# resid_mode left to the library + flat prior => flat-prior Kronecker kernel (bvg_kron_sweep.cuh)
CASES["kron_c1"] = (lambda: helpers.model_c1(iterations=40, burnin=10), {})
for _p in range(0, 5):
    CASES["kron_grid_p%d" % _p] = ((lambda p=_p: _grid_model(p)), {})
CASES["kron_k2"] = (lambda: _grid_model(2, K=2), {})
CASES["kron_k1"] = (lambda: _grid_model(3, K=1), {})


def _kron_k4(p):
    """Four variables (synthetic): the generic k x k factor of the state recursion, and the M = 2 k + 1 / run-time-M paths."""
    y = helpers.synth_var(4, 2, 90, 31)
    m = helpers.gen_var(y, p=p, deterministic="const", iterations=20, burnin=5)
    return helpers.add_priors(m, coef=dict(v_i=0, v_i_det=0), sigma=dict(df="k", scale=1e-3))


CASES["kron_k4_p2"] = (lambda: _kron_k4(2), {})
CASES["kron_k4_p5"] = (lambda: _kron_k4(5), {})


def _tvp_wishart():
    y = helpers.synth_var(2, 1, 28, 11)
    m = helpers.gen_var(y, p=1, deterministic="const", iterations=8, burnin=3, tvp=True)
    return helpers.add_priors(m, coef=dict(v_i=1, v_i_det=0.1, shape=3, rate=1e-4, rate_det=0.01),
                              sigma=dict(df="k", scale=1))


CASES["tvp_wishart"] = (_tvp_wishart, {})
CASES["tvp_covar_sv"] = (lambda: helpers.model_tvp_covar(), {})
CASES["tvp_covar_sv_k4"] = (lambda: helpers.model_tvp_covar(K=4, p=1, nobs=20, iterations=6, burnin=2), {})
CASES["tvp_covar_sv_k2"] = (lambda: helpers.model_tvp_covar(K=2, p=2, nobs=22), {})
CASES["tvp_covar_gamma"] = (lambda: helpers.model_tvp_covar(sv=False), {})
CASES["tvp_covar_sv_psi_bvs"] = (lambda: helpers.model_tvp_covar(K=4, p=1, nobs=20, iterations=8, burnin=2, psi_bvs=True), {})
CASES["structural_gamma"] = (lambda: helpers.model_structural(), {})
CASES["structural_gamma_k4_p2"] = (lambda: helpers.model_structural(K=4, p=2, nobs=60), {})
CASES["structural_sv"] = (lambda: helpers.model_structural(sv=True), {})
CASES["structural_gamma_bvs"] = (lambda: helpers.model_structural(bvs=True), {})
CASES["structural_sv_bvs_k2"] = (lambda: helpers.model_structural(K=2, p=2, sv=True, bvs=True), {})
# SSVS on a structural model (bvaralg.cpp:107-135 with the structural z): the static mask travels beside lambda
CASES["structural_gamma_ssvs"] = (lambda: helpers.model_structural(ssvs=True, iterations=16, burnin=4), {})
CASES["structural_gamma_ssvs_k4_p2"] = (lambda: helpers.model_structural(K=4, p=2, nobs=60, ssvs=True), {})
CASES["tvp_covar_gamma_bvs"] = (lambda: helpers.model_tvp_covar(sv=False, bvs=True), {})


def _exogen(sv=False, ssvs=False):
    """VARX: two exogenous regressors with one lag (B coefficients, R/gen_var.R:143-161) next to lags and a constant."""
    y = helpers.synth_var(3, 2, 80, 5)
    x = np.random.default_rng(1).standard_normal((80, 2))
    m = helpers.gen_var(y, p=2, exogen=x, s=1, deterministic="const", iterations=12, burnin=3, sv=sv)
    if sv:
        return helpers.add_priors(m, coef=dict(v_i=1, v_i_det=0.1),
                                  sigma=dict(shape=3, rate=1e-4, mu=0, v_i=0.01, sigma_h=0.05, constant=1e-4, covar=True))
    return helpers.add_priors(m, coef=dict(v_i=1, v_i_det=0.1), sigma=dict(df="k", scale=1),
                              ssvs=dict(inprior=0.5, tau=(0.05, 10.0), exclude_det=True) if ssvs else None)


CASES["exogen_wishart"] = (_exogen, {})
CASES["exogen_ssvs"] = (lambda: _exogen(ssvs=True), {})
CASES["exogen_sv_covar"] = (lambda: _exogen(sv=True), {})


# Structural TVP models (A0_t random walks; bvartvpalg.cpp:44-47,544,594)
CASES["structural_tvp_gamma"] = (lambda: helpers.model_structural(tvp=True, nobs=24, iterations=6, burnin=2), {})
CASES["structural_tvp_sv"] = (lambda: helpers.model_structural(tvp=True, sv=True, nobs=24, iterations=6, burnin=2), {})
CASES["structural_tvp_gamma_bvs"] = (lambda: helpers.model_structural(tvp=True, bvs=True, nobs=24, iterations=8, burnin=2), {})
CASES["structural_tvp_sv_bvs_k4"] = (lambda: helpers.model_structural(tvp=True, K=4, p=1, sv=True, bvs=True, nobs=20,
                                                                      iterations=6, burnin=2), {})
ALL_CASES = CASES


@pytest.mark.parametrize("name", sorted(ALL_CASES))
def test_kernel_logic_matches_oracle(emu, name):
    builder, kw = ALL_CASES[name]
    obj = builder()
    table = kw.get("mixture", "ksc1998")
    per, stacked = helpers.make_variates(obj, 2, seed=11)
    fn = emu.emu_bvartvpalg if obj["model"].get("tvp") else emu.emu_bvaralg
    rc, out = helpers.run_abi(fn, obj, 2, stacked, **kw)
    assert rc == 0, emu.emu_last_error()
    want = helpers.oracle_stack(obj, per, table=table)
    assert set(want) == set(out.arrays)
    for key in want:
        assert helpers.relerr(out.arrays[key], want[key]) < TOL, (name, key)
    if "lambda" in want:
        assert np.array_equal(out.arrays["lambda"], want["lambda"])
    assert not out.status.any()


@pytest.mark.parametrize("G", [4, 32])
@pytest.mark.parametrize("name", ["kron_c1", "kron_grid_p4", "kron_k2", "c1_wishart_moment", "c1_informative_prior", "ssvs_moment", "bvs_direct", "gamma",
                                  "sv_bvs", "tvp_sv_bvs", "tvp_wishart", "tvp_covar_sv", "tvp_covar_gamma",
                                  "structural_gamma_bvs", "structural_tvp_sv"])
def test_kernel_logic_spmd_lanes_match_oracle(emu, name, G):
    """Same check with the chain owned by G real threads (barriers + emulated shuffles): exercises the
    one-row-per-lane Cholesky / back substitution, the register Wishart block and the work distribution."""
    builder, kw = ALL_CASES[name]
    obj = builder()
    per, stacked = helpers.make_variates(obj, 1, seed=13)
    fn = emu.emu_bvartvpalg if obj["model"].get("tvp") else emu.emu_bvaralg
    rc, out = helpers.run_abi(fn, obj, 1, stacked, threads_per_chain=G, **kw)
    assert rc == 0, emu.emu_last_error()
    want = helpers.oracle_stack(obj, per, table=kw.get("mixture", "ksc1998"))
    for key in want:
        assert helpers.relerr(out.arrays[key], want[key]) < TOL, (name, G, key)
    if "lambda" in want:
        assert np.array_equal(out.arrays["lambda"], want["lambda"])


def test_chunked_launches_are_bit_identical(emu):
    """State save/restore between kernel launches must not change anything (Philox sites are sweep-addressed)."""
    obj = helpers.model_c3(K=3, p=2, nobs=80)
    rc1, a = helpers.run_abi(emu.emu_bvaralg, obj, 3, None, seed=42)
    rc2, b = helpers.run_abi(emu.emu_bvaralg, obj, 3, None, seed=42, sweeps_per_launch=7)
    assert rc1 == 0 and rc2 == 0
    for key in a.arrays:
        assert np.array_equal(a.arrays[key], b.arrays[key]), key


def test_sharding_invariance(emu):
    """Chains are keyed by GLOBAL id: chains [2,5) computed alone equal chains 2..4 of a 6-chain call."""
    obj = helpers.model_c1(iterations=15, burnin=5)
    _, full = helpers.run_abi(emu.emu_bvaralg, obj, 6, None, seed=99)
    _, part = helpers.run_abi(emu.emu_bvaralg, obj, 3, None, seed=99, chain_offset=2)
    for key in full.arrays:
        assert np.array_equal(full.arrays[key][2:5], part.arrays[key]), key
    _, other = helpers.run_abi(emu.emu_bvaralg, obj, 6, None, seed=100)
    assert not np.array_equal(full.arrays["coef"], other.arrays["coef"])


def test_philox_known_answer(emu):
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors)."""
    out = (C.c_uint32 * 4)()
    emu.emu_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
    emu.emu_philox(0, 0, 0, 0, 0, 0, out)
    assert list(out) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    emu.emu_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out)
    assert list(out) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    emu.emu_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out)
    assert list(out) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_variate_transform_arithmetic(emu):
    """Range-specific log / sin-cos polynomials of bvg_common.cuh against libm (host build: the device replaces the
    division / sqrt by MUFU seeds + Newton steps, covered by the device-vs-emulator test), the uniform construction, and
    the committed coefficient tables against tools/fit_variate_polys.py."""
    rng = np.random.default_rng(3)
    u = np.concatenate([rng.random(200000), np.ldexp(rng.random(2000), -rng.integers(1, 52, 2000)),
                        1.0 - np.ldexp(rng.random(2000), -rng.integers(1, 50, 2000))])
    u = u[(u > 0) & (u < 1)]
    n = u.shape[0]
    outs = [np.empty(n) for _ in range(4)]
    dp = C.POINTER(C.c_double)
    emu.emu_transforms.argtypes = [C.c_int] + [dp] * 5
    emu.emu_transforms(n, u.ctypes.data_as(dp), *[o.ctypes.data_as(dp) for o in outs])
    lg, sq, sn, cs = outs
    assert np.max(np.abs(lg - np.log(u)) / np.abs(np.log(u))) < 4.5e-16
    assert np.max(np.abs(sq - np.sqrt(u)) / np.sqrt(u)) < 2.3e-16
    assert np.max(np.abs(sn - np.sin(2 * np.pi * u))) < 1.5e-15 and np.max(np.abs(cs - np.cos(2 * np.pi * u))) < 1.5e-15
    assert np.max(np.abs(sn * sn + cs * cs - 1.0)) < 5e-16
    big = np.array([3.0, 75.0, 1e-30, 5e10])
    outs = [np.empty(4) for _ in range(4)]
    emu.emu_transforms(4, big.ctypes.data_as(dp), *[o.ctypes.data_as(dp) for o in outs])
    assert np.max(np.abs(outs[0] - np.log(big)) / np.abs(np.log(big))) < 4.5e-16
    emu.emu_u53.restype = C.c_double
    emu.emu_u53.argtypes = [C.c_uint32, C.c_uint32]
    assert emu.emu_u53(0, 0) == 2.0 ** -53 and emu.emu_u53(0xffffffff, 0xffffffff) == 1.0 - 2.0 ** -53
    assert emu.emu_u53(0x80000000, 0) == 0.5 + 2.0 ** -53
    # tables in the header == the fit script (first 12 significant digits: the fit is deterministic, the header is a paste)
    import importlib.util
    import pathlib
    import re
    root = pathlib.Path(__file__).resolve().parents[1]
    hdr = (root / "bvartools_b200" / "csrc" / "bvg_common.cuh").read_text()
    spec = importlib.util.spec_from_file_location("fit_variate_polys", root / "tools" / "fit_variate_polys.py")
    fit = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fit)
    mp = fit.mp
    zmax = (mp.pi / 4) ** 2
    want = {"kSinTab": fit.cheb_fit(lambda z: mp.sin(mp.sqrt(z)) / mp.sqrt(z) if z > 0 else mp.mpf(1), mp.mpf(0), zmax, 6),
            "kCosTab": fit.cheb_fit(lambda z: mp.cos(mp.sqrt(z)), mp.mpf(0), zmax, 7)}
    for name, coef in want.items():
        m = re.search(r"BVG_TABLE\(" + name + r",\s*\d+,([^)]*)\)", hdr)
        got = [float(v) for v in m.group(1).replace("\n", " ").split(",")]
        np.testing.assert_allclose(got[:len(coef)], [float(c) for c in coef], rtol=1e-12)


def test_variate_transforms_are_distributionally_right(emu):
    """Box-Muller normals, uniforms, Marsaglia-Tsang gammas / chi-squares from the Philox stream: KS tests."""
    from scipy import stats
    emu.emu_variates.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.POINTER(C.c_double)]
    n = 200000
    buf = np.empty(n)
    ptr = buf.ctypes.data_as(C.POINTER(C.c_double))
    emu.emu_variates(123, 5, 0, 3, n, 0, 0.0, ptr)
    assert stats.kstest(buf, "uniform").pvalue > 1e-3 and buf.min() > 0 and buf.max() < 1
    emu.emu_variates(123, 5, 0, 3, n, 1, 0.0, ptr)
    assert stats.kstest(buf, "norm").pvalue > 1e-3
    for shape in (0.4, 1.0, 3.5, 37.0):
        emu.emu_variates(7, 1, 6, 0, n, 2, shape, ptr)
        assert stats.kstest(buf, "gamma", args=(shape,)).pvalue > 1e-3, shape
    emu.emu_variates(7, 1, 9, 0, n, 3, 74.0, ptr)
    assert stats.kstest(buf, "chi2", args=(74.0,)).pvalue > 1e-3


def test_invalid_specs_are_rejected(emu):
    obj = helpers.model_c1()
    spec = helpers._abi.spec_from_model(obj, n_chains=1)
    out = helpers._abi.OutBuffers(spec)
    spec.c.abi_version = 99
    assert emu.emu_bvaralg(C.byref(spec.c), C.byref(out.c)) == -2
    spec = helpers._abi.spec_from_model(obj, n_chains=1)
    spec.keep["Y"][3] = np.nan
    assert emu.emu_bvaralg(C.byref(spec.c), C.byref(out.c)) == -6
    assert b"NAs" in emu.emu_last_error()
    # SSVS + SV is illegal (src/bvaralg.cpp:116-118)
    spec = helpers._abi.spec_from_model(helpers.model_sv(), n_chains=1)
    spec.c.varsel = helpers._abi.VARSEL_SSVS
    out = helpers._abi.OutBuffers(spec)
    assert emu.emu_bvaralg(C.byref(spec.c), C.byref(out.c)) == -7
    assert b"SSVS with stochastic volatility" in emu.emu_last_error()


def test_unsupported_combinations_of_the_new_branches_are_rejected(emu):
    """Error behaviour of the structural / TVP-psi branches: what the reference forbids, or what is not built, is an
    error code with the reference's message where there is one -- never a silent fallback."""
    # structural + Wishart (R/add_priors.bvarmodel.R: "Structural models may not use a Wishart prior.")
    spec = helpers._abi.spec_from_model(helpers.model_structural(), n_chains=1)
    spec.c.sigma_type = 0
    out = helpers._abi.OutBuffers(spec)
    assert emu.emu_bvaralg(C.byref(spec.c), C.byref(out.c)) == -7
    assert b"Wishart" in emu.emu_last_error()
    # structural + SSVS without its prior arrays
    spec = helpers._abi.spec_from_model(helpers.model_structural(), n_chains=1)
    spec.c.varsel = helpers._abi.VARSEL_SSVS
    assert emu.emu_bvaralg(C.byref(spec.c), C.byref(out.c)) == -4
    # structural TVP model together with the covariance block
    cov = helpers._abi.spec_from_model(helpers.model_tvp_covar(), n_chains=1)
    spec = helpers._abi.spec_from_model(helpers.model_structural(tvp=True, sv=True, nobs=24), n_chains=1)
    spec.c.psi_mu, spec.c.psi_vi = cov.c.psi_mu, cov.c.psi_vi
    out = helpers._abi.OutBuffers(spec)
    assert emu.emu_bvartvpalg(C.byref(spec.c), C.byref(out.c)) == -7
    # TVP psi block without its state-variance prior / initial draw
    spec = helpers._abi.spec_from_model(helpers.model_tvp_covar(), n_chains=1)
    spec.c.psi_rate = None
    out = helpers._abi.OutBuffers(spec)
    assert emu.emu_bvartvpalg(C.byref(spec.c), C.byref(out.c)) == -4
    # SSVS on the covariances of a TVP-SV model: rejected like SSVS + SV (src/bvaralg.cpp:173-175)
    spec = helpers._abi.spec_from_model(helpers.model_tvp_covar(), n_chains=1)
    spec.c.psi_varsel = helpers._abi.VARSEL_SSVS
    assert emu.emu_bvartvpalg(C.byref(spec.c), C.byref(out.c)) == -11
    assert b"SSVS" in emu.emu_last_error()
    # covariance block together with structural coefficients (R/add_priors.bvarmodel.R)
    cov = helpers._abi.spec_from_model(helpers.model_covar_gamma(), n_chains=1)
    spec = helpers._abi.spec_from_model(helpers.model_structural(nobs=70), n_chains=1)
    spec.c.psi_mu, spec.c.psi_vi = cov.c.psi_mu, cov.c.psi_vi
    out = helpers._abi.OutBuffers(spec)
    assert emu.emu_bvaralg(C.byref(spec.c), C.byref(out.c)) == -7
    assert b"cannot be estimated at the same time" in emu.emu_last_error()


def test_non_spd_prior_flags_chain_status(emu):
    obj = helpers.model_c1(iterations=5, burnin=1, coef=dict(v_i=1, v_i_det=0.1))
    obj["priors"]["coefficients"]["v_i"] = -1e6 * np.eye(21)
    rc, out = helpers.run_abi(emu.emu_bvaralg, obj, 2, None, seed=1)
    assert rc == 0 and out.status.all()
