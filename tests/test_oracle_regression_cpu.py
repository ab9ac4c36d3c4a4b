"""The oracle's own arithmetic for the branches without reference fixtures (structural, covariance blocks) is locked
against tests/golden/oracle_regression.json (tools/make_oracle_regression.py): a deliberate change of the restatement
has to regenerate the file, an accidental one fails here."""
import importlib.util
import json
import pathlib

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_oracle_outputs_are_unchanged():
    spec = importlib.util.spec_from_file_location("make_oracle_regression", ROOT / "tools" / "make_oracle_regression.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.loads((ROOT / "tests" / "golden" / "oracle_regression.json").read_text())
    got = mod.compute()
    assert set(got) == set(want)
    for name in want:
        assert set(got[name]) == set(want[name]), name
        for key in want[name]:
            np.testing.assert_allclose(got[name][key], want[name][key], rtol=1e-9, atol=1e-12, err_msg=f"{name}/{key}")
