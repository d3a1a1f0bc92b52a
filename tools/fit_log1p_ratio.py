"""Fit the polynomial of `log1p_ratio` (csrc/bean_math.cuh): log1p(y) = 2 atanh(s) = 2 s + s^3 P(s^2), s = y / (2 + y),
for 1 + y in [lo, hi] (default [0.5, 2] -> |s| <= 1/3).  Chebyshev interpolation in high precision; the printed error is
the float32 evaluation (exact division) against mpmath, relative to log1p(y).

    python tools/fit_log1p_ratio.py [deg] [lo] [hi]
"""
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 40


def Pf(t):  # (2 atanh(s) - 2 s) / s^3 as a function of t = s^2
    if t == 0:
        return mp.mpf(2) / 3
    s = mp.sqrt(t)
    return (2 * mp.atanh(s) - 2 * s) / s ** 3


def cheb_fit(f, S, deg):
    n = deg + 1
    nodes = [(mp.cos(mp.pi * (2 * k + 1) / (2 * n)) + 1) / 2 * S for k in range(n)]
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, x in enumerate(nodes):
        for j in range(n):
            A[i, j] = x ** j
        b[i] = f(x)
    c = mp.lu_solve(A, b)
    return [c[j] for j in range(n)]


def main():
    deg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    lo = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    hi = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
    smax = max(abs((lo - 1) / (lo + 1)), abs((hi - 1) / (hi + 1)))
    c = cheb_fit(Pf, mp.mpf(smax) ** 2, deg)
    c32 = [np.float32(float(v)) for v in c]
    ys = np.concatenate([np.linspace(lo - 1, hi - 1, 20001), np.geomspace(1e-7, 0.3, 500), -np.geomspace(1e-7, 0.3, 500)]).astype(np.float32)
    ys = ys[(ys > lo - 1) & (ys < hi - 1) & (ys != 0)]
    d = (np.float32(2) + ys).astype(np.float32)
    s = (ys / d).astype(np.float32)
    t = (s * s).astype(np.float32)
    p = np.full_like(t, c32[-1])
    for k in range(deg - 1, -1, -1):
        p = (p * t + c32[k]).astype(np.float32)
    r = ((s * t).astype(np.float32) * p + (s + s)).astype(np.float32)
    err = max(abs(float(r[i]) / float(mp.log1p(mp.mpf(float(ys[i])))) - 1) for i in range(0, len(ys), 3))
    print(f"deg={deg} 1+y in [{lo}, {hi}] |s|<={smax:.4f}: max rel err {err:.2e}")
    print("P:", ", ".join(f"{float(v):.9e}f" for v in c))


if __name__ == "__main__":
    main()


def approx_error(deg, lo, hi):
    """error of the fitted polynomial alone (exact arithmetic), relative to log1p"""
    smax = max(abs((lo - 1) / (lo + 1)), abs((hi - 1) / (hi + 1)))
    c = cheb_fit(Pf, mp.mpf(smax) ** 2, deg)
    worst = 0
    for i in range(1, 400):
        s = mp.mpf(smax) * i / 400
        t = s * s
        p = sum(c[k] * t ** k for k in range(deg + 1))
        worst = max(worst, abs((2 * s + s ** 3 * p) / (2 * mp.atanh(s)) - 1))
    return float(worst)
