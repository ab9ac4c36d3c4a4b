"""CPU tests of the boundary: the shared library loads, exports every symbol include/bvargpu.h declares, the
ctypes mirror matches the C structs, and the product fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import pathlib
import re
import subprocess

import numpy as np
import pytest

import helpers
from bvartools_b200 import _abi
from bvartools_b200._lib import EXPORTED, BvgError, lib

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "bvargpu.h").read_text()
    declared = set(re.findall(r"\b(bvg_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(EXPORTED), declared ^ set(EXPORTED)
    L = lib()
    for sym in declared:
        assert hasattr(L, sym), sym
    assert b"sm_100a" in L.bvg_version()


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof of the ctypes mirror == what the C compiler sees."""
    src = tmp_path / "lay.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bvargpu.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(bvg_spec),sizeof(bvg_out),sizeof(bvg_inject),offsetof(bvg_spec,seed),offsetof(bvg_spec,sigma_i),'
                   'offsetof(bvg_spec,inject),offsetof(bvg_out,status),offsetof(bvg_spec,n_gpus));return 0;}\n')
    exe = tmp_path / "lay"
    subprocess.check_call(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_abi.BvgSpec), C.sizeof(_abi.BvgOut), C.sizeof(_abi.BvgInject), _abi.BvgSpec.seed.offset,
            _abi.BvgSpec.sigma_i.offset, _abi.BvgSpec.inject.offset, _abi.BvgOut.status.offset, _abi.BvgSpec.n_gpus.offset]
    assert got == want


def test_out_rows():
    L = lib()
    for obj, want in ((helpers.model_c1(), (21, 9, 0, 0, 0)),
                      (helpers.model_c3(K=3, p=2, nobs=60), (21, 9, 21, 0, 0)),
                      (helpers.model_sv(nobs=40, tvp=True), (21 * 38, 9 * 38, 0, 3, 21)),
                      # structural: k M + k (k - 1) / 2 stored rows (constant and TVP), BVS indicators on the same rows
                      (helpers.model_structural(K=3, p=1, nobs=30), (15, 9, 0, 0, 0)),
                      (helpers.model_structural(K=3, p=1, nobs=30, bvs=True), (15, 9, 15, 0, 0)),
                      (helpers.model_structural(K=3, p=1, nobs=30, ssvs=True), (15, 9, 15, 0, 0)),
                      (helpers.model_structural(K=3, p=1, nobs=30, tvp=True), (15 * 29, 9, 0, 0, 15)),
                      (helpers.model_structural(K=3, p=1, nobs=30, tvp=True, sv=True, bvs=True),
                       (15 * 29, 9 * 29, 15, 3, 15)),
                      # covariance (psi) block: Sigma_t per period for TVP models even with gamma variances; psi
                      # inclusion draws appended to lambda
                      (helpers.model_tvp_covar(K=3, p=1, nobs=24, sv=False), (12 * 23, 9 * 23, 0, 0, 12)),
                      (helpers.model_tvp_covar(K=3, p=1, nobs=24, sv=True, bvs=True, psi_bvs=True),
                       (12 * 23, 9 * 23, 12 + 3, 3, 12)),
                      (helpers.model_covar_gamma(K=3, p=2, nobs=40, psi_sel="ssvs"), (21, 9, 24, 0, 0)),
                      (helpers.model_covar_gamma(K=3, p=2, nobs=40, bvs=True, psi_sel="bvs"), (21, 9, 24, 0, 0))):
        spec = _abi.spec_from_model(obj)
        vals = [C.c_int64() for _ in range(5)]
        assert L.bvg_out_rows(C.byref(spec.c), *[C.byref(v) for v in vals]) == 0
        assert tuple(v.value for v in vals) == want
        rows = spec.out_rows()
        assert (rows["coef"], rows["sigma"], rows["lambda"], rows["sigma_h"], rows["sigma_v"]) == want


def test_no_cpu_fallback_without_device():
    L = lib()
    if L.bvg_device_count() > 0:
        pytest.skip("a GPU is present")
    import bvartools_b200 as b
    with pytest.raises(BvgError, match="no CUDA device"):
        b.bvaralg(helpers.model_c1())
    with pytest.raises(BvgError, match="no CUDA device"):
        b.draw_posterior([helpers.model_c1()], FUN=lambda m: b.bvaralg(m))[0] if False else b.run_chains(helpers.model_c1(), 4)
    res = b.draw_posterior(helpers.model_c1())            # try() semantics: failure -> error object
    assert res.get("error") is True


def test_post_estimation_calls_fail_loudly_without_device():
    """summary / irf / fevd / predict of host objects go through the device or raise: no NumPy fallback."""
    L = lib()
    if L.bvg_device_count() > 0:
        pytest.skip("a GPU is present")
    import bvartools_b200 as b
    import numpy as np
    rng = np.random.default_rng(0)
    fit = {"y": rng.standard_normal((20, 2)), "x": rng.standard_normal((20, 5)), "A": rng.standard_normal((30, 4)),
           "C": rng.standard_normal((30, 2)), "Sigma": np.tile(np.eye(2).reshape(-1), (30, 1))}
    for call in (lambda: b.summary(fit), lambda: b.irf(fit, 0, 1), lambda: b.fevd(fit, 0), lambda: b.predict(fit, 3)):
        with pytest.raises(BvgError, match="no CUDA device"):
            call()
    from bvartools_b200.distributed import bind_to_gpu_numa
    assert bind_to_gpu_numa(0) is False


def test_product_never_imports_the_oracle_or_the_emulator():
    for path in list((ROOT / "bvartools_b200").rglob("*.py")) + list((ROOT / "bvartools_b200" / "csrc").glob("*")):
        if path.is_file() and path.suffix in (".py", ".cu", ".cuh", ".hpp", ".cpp"):
            txt = path.read_text()
            assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), path
            assert "_bvgemu" not in txt, path


def test_gen_var_grid_shares_sample():
    """p = 0:4 grid on the common sample (R/gen_var.R:109,116-128,172; SURVEY.md config c2-B: T = 71)."""
    x, names = helpers.e1_readme_sample()
    models = helpers.gen_var(x, p=range(0, 5), deterministic="const", iterations=10, burnin=2, names=names)
    assert [m["model"]["endogen"]["lags"] for m in models] == [0, 1, 2, 3, 4]
    assert all(m["data"]["Y"].shape == (71, 3) for m in models)
    assert [m["data"]["Z"].shape[1] for m in models] == [1, 4, 7, 10, 13]
    pri = helpers.add_priors(models, coef=dict(v_i=0, v_i_det=0), sigma=dict(df="k", scale=1e-4))
    assert all(p["priors"]["sigma"]["df"] == 3.0 for p in pri)


def test_seasonal_dummies_follow_the_cycle():
    """R/gen_var.R:188-203: frequency - 1 dummies, the first one on the first cycle-1 row of the trimmed sample."""
    from bvartools_b200 import model as M
    y = helpers.synth_var(2, 1, 30, 3)
    m = M.gen_var(y, p=1, deterministic="const", seasonal=True, frequency=4, start_cycle=2, iterations=5, burnin=2)
    assert m["model"]["deterministic"] == ["const", "season.1", "season.2", "season.3"]
    Z = m["data"]["Z"]
    # first row of the trimmed sample is cycle 3: cycle 1 arrives at row 3 (1-based) -> dummies on cycles 1, 2, 3
    assert Z[:6, -3].tolist() == [0, 0, 1, 0, 0, 0] and Z[:6, -2].tolist() == [0, 0, 0, 1, 0, 0]
    assert Z[:6, -1].tolist() == [1, 0, 0, 0, 1, 0]
    assert np.all(Z[:, -3:].sum(axis=1) <= 1)
    obj = M.add_priors(m, coef=dict(v_i=1, v_i_det=0.1), sigma=dict(df="k", scale=1))
    assert obj["priors"]["coefficients"]["v_i"].shape == (2 * (2 + 4), 2 * (2 + 4))
    with pytest.raises(ValueError, match="seasonal"):
        M.gen_var(y, p=1, deterministic="trend", seasonal=True, frequency=4)


def test_minnesota_and_ssvs_priors():
    obj = helpers.gen_var(helpers.synth_var(4, 2, 90, 3), p=2, deterministic="const", iterations=5, burnin=1)
    m = helpers.add_priors(obj, coef=dict(minnesota=dict(kappa0=2, kappa1=0.5, kappa3=5)), sigma=dict(df="k", scale=1))
    v = 1.0 / np.diag(m["priors"]["coefficients"]["v_i"])
    V = v.reshape(4, 9, order="F")
    assert np.allclose(np.diag(V[:, :4]), 2.0) and np.allclose(np.diag(V[:, 4:8]), 0.5)     # kappa0 / r^2
    s = helpers.add_priors(obj, sigma=dict(df=0, scale=1e-5),
                           ssvs=dict(inprior=0.5, semiautomatic=(0.01, 10.0), exclude_det=True))
    sel = s["priors"]["coefficients"]["ssvs"]
    assert sel["include"].shape[0] == 32 and sel["include"].max() == 32
    assert np.allclose(sel["tau1"] / sel["tau0"], 1000.0)
    assert np.allclose(np.diag(s["priors"]["coefficients"]["v_i"]), 1.0 / sel["tau1"][:, 0] ** 2)
