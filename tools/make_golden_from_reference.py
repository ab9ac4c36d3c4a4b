#!/usr/bin/env python3
"""Generate the small committed fixtures from the reference tree (run in the build container only).

/root/reference does not exist on the GPU box, so everything derived from it is written once into
tests/golden/ and bvartools_b200/data/ and committed together with this script.

  * e1.csv                 <- data-raw/e1.dat  (West German investment / income / consumption, 92 x 3)
  * readme_table_c1.json   <- README.md:249-299, the printed summary(bvar_est) table of the README example
                              (the only numeric known-answer in the reference tree)
"""
import json, re, sys, pathlib

REF = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
ROOT = pathlib.Path(__file__).resolve().parents[1]

def make_e1():
    rows, names = [], None
    in_comment = False
    for ln in (REF / "data-raw" / "e1.dat").read_text().splitlines():
        s = ln.strip()
        if not s:
            continue
        if s.startswith("/*"):
            in_comment = True
        if in_comment:
            if s.endswith("*/"):
                in_comment = False
            continue
        if s.startswith("<"):
            continue
        toks = s.split()
        if names is None:
            names = toks
            continue
        rows.append([int(t) for t in toks])
    assert names == ["invest", "income", "cons"] and len(rows) == 92, (names, len(rows))
    txt = "# e1: quarterly, seasonally adjusted West German fixed investment, disposable income, consumption\n" \
          "# expenditures, billions of DM, 1960Q1-1982Q4 (Deutsche Bundesbank; Luetkepohl 2006, file E1)\n"
    txt += ",".join(names) + "\n" + "\n".join(",".join(map(str, r)) for r in rows) + "\n"
    for dst in (ROOT / "tests" / "golden" / "e1.csv", ROOT / "bvartools_b200" / "data" / "e1.csv"):
        dst.write_text(txt)

def make_readme_table():
    lines = (REF / "README.md").read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if "Bayesian VAR model with p = 2" in l)
    block = [l[6:] if l.startswith("    ## ") else "" for l in lines[start:start + 70]]
    out = {"source": "README.md (summary(bvar_est) of the R-level post_normal/rWishart loop, set.seed(1234567), "
                     "10000 kept / 5000 burn-in, priors v_i=0, df=1, scale=1e-4)",
           "coef": {}, "sigma": {}}
    cur, hdr = None, None
    tail = {}            # the 'cons' block wraps its 97.5% column onto a second sub-table
    for l in block:
        m = re.match(r"\s*Variable: (\w+)", l)
        if m:
            cur, hdr = m.group(1), None
            out["coef"][cur] = {}
            continue
        if "Variance-covariance matrix" in l:
            cur, hdr = "__sigma__", None
            continue
        if cur is None or not l.strip():
            continue
        toks = l.split()
        if toks[0] in ("Mean", "97.5%"):
            hdr = ["Mean", "SD", "NaiveSD", "TSSD", "q2.5", "q50", "q97.5"] if toks[0] == "Mean" else ["q97.5"]
            if toks[0] == "Mean" and "97.5%" not in toks:
                hdr = hdr[:-1]
            continue
        if hdr is None:
            continue
        name, vals = toks[0], [float(t) for t in toks[1:]]
        if len(vals) != len(hdr):
            continue
        tgt = out["sigma"] if cur == "__sigma__" else out["coef"][cur]
        tgt.setdefault(name, {}).update(dict(zip(hdr, vals)))
    assert len(out["sigma"]) == 9 and all(len(v) == 7 for v in out["coef"].values())
    assert all(len(s) == 7 for v in out["coef"].values() for s in v.values()), out["coef"]["cons"]
    (ROOT / "tests" / "golden" / "readme_table_c1.json").write_text(json.dumps(out, indent=1) + "\n")

if __name__ == "__main__":
    make_e1()
    make_readme_table()
    print("fixtures written")
