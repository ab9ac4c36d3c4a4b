#!/usr/bin/env python3
"""Coefficient tables of bvartools_b200/csrc/bvg_common.cuh (kSinTab, kCosTab, kLogTab): interpolation at the Chebyshev
nodes of the argument interval with 60-digit arithmetic (near-minimax), printed as doubles together with the maximum
error of the exact polynomial.  usage: python tools/fit_variate_polys.py"""
import mpmath as mp

mp.mp.dps = 60


def cheb_fit(f, a, b, deg):
    n = deg + 1
    nodes = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A, y = mp.matrix(n, n), mp.matrix(n, 1)
    for i, x in enumerate(nodes):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[i] for i in range(n)]


def max_err(f, c, a, b, n=4000):
    worst = mp.mpf(0)
    for i in range(n + 1):
        x = a + (b - a) * i / n
        worst = max(worst, abs(sum(ci * x ** k for k, ci in enumerate(c)) - f(x)))
    return worst


def main():
    zmax = (mp.pi / 4) ** 2                                   # z = x^2, |x| <= pi / 4
    f_sin = lambda z: mp.sin(mp.sqrt(z)) / mp.sqrt(z) if z > 0 else mp.mpf(1)
    f_cos = lambda z: mp.cos(mp.sqrt(z))
    zlog = ((mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)) ** 2         # z = f^2, f = (m - 1) / (m + 1), m in [2^-1/2, 2^1/2)

    def f_log(z):                                             # 2 atanh(f) = 2 f + f z P(z)
        if z == 0:
            return mp.mpf(2) / 3
        f = mp.sqrt(z)
        return (mp.atanh(f) / f - 1) * 2 / z

    for name, f, hi, deg, scale in (("kSinTab", f_sin, zmax, 6, 1), ("kCosTab", f_cos, zmax, 7, 1),
                                    ("kLogTab", f_log, zlog, 6, zlog / 2)):
        c = cheb_fit(f, mp.mpf(0), hi, deg)
        print(name, ", ".join(repr(float(x)) for x in c))
        print("   max error", mp.nstr(max_err(f, c, mp.mpf(0), hi) * scale, 3))


if __name__ == "__main__":
    main()
