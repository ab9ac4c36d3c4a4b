"""Regression lock of the ORACLE itself (not a reference fixture: the reference ships no numeric fixtures for these
branches): for a few small models with seeded injected variates, store the last retained draw of every output of the
literal restatement.  tests/test_oracle_regression_cpu.py recomputes them; a change means the oracle's arithmetic changed.
Run from the repo root:  python tools/make_oracle_regression.py"""
import json
import sys

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import helpers  # noqa: E402

MODELS = {
    "structural_gamma_bvs": lambda: helpers.model_structural(bvs=True),
    "structural_sv": lambda: helpers.model_structural(sv=True),
    "tvp_covar_sv": lambda: helpers.model_tvp_covar(),
    "tvp_covar_gamma_psi_bvs": lambda: helpers.model_tvp_covar(sv=False, psi_bvs=True),
    "covar_gamma_psi_ssvs": lambda: helpers.model_covar_gamma(psi_sel="ssvs"),
}


def compute():
    out = {}
    for name, build in MODELS.items():
        obj = build()
        per, _ = helpers.make_variates(obj, 1, seed=123)
        res = helpers.oracle_stack(obj, per)
        out[name] = {k: [float(v) for v in np.asarray(a)[0, -1].reshape(-1)[:12]] for k, a in res.items()}
    return out


if __name__ == "__main__":
    with open("tests/golden/oracle_regression.json", "w") as f:
        json.dump(compute(), f, indent=1)
    print("written tests/golden/oracle_regression.json")
