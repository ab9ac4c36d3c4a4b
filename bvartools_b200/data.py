"""Data sets shipped with the package (the reference ships them as .rda, R/data.R:25-93)."""
from __future__ import annotations

import pathlib

import numpy as np

_DIR = pathlib.Path(__file__).resolve().parent / "data"


def load_e1():
    """West German investment / income / consumption, 1960Q1-1982Q4 (92 x 3).  Returns (array, names)."""
    arr = np.loadtxt(_DIR / "e1.csv", delimiter=",", comments="#", skiprows=3)
    return arr, ["invest", "income", "cons"]


def e1_readme_sample():
    """README.md:110-116 -- diff(log(e1)) * 100 windowed to 1978Q4 (75 rows)."""
    arr, names = load_e1()
    x = np.diff(np.log(arr), axis=0) * 100.0
    return x[:75], names
