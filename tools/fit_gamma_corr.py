"""Fit the polynomials behind `gamma_corr` (csrc/bean_math.cuh).

cv(z) = lgamma(z) - [(z - 1/2) ln z - z + ln(2 pi)/2]   (Binet's function: exactly odd in w = 1/z)
dl(z) = digamma(z) - ln z = -w/2 - w^2 R(w^2)
Form:  cv = w Q(s),  dl = -w/2 - s R(s),  s = w^2,  fitted on s in [0, S] (z >= 1/sqrt(S)) by Chebyshev
interpolation in high precision; the printed error is of the FLOAT32 Horner evaluation against mpmath.

    python tools/fit_gamma_corr.py [zmin] [degQ] [degR]
"""
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 40


def cv(z):
    z = mp.mpf(z)
    return mp.loggamma(z) - ((z - mp.mpf(1) / 2) * mp.log(z) - z + mp.log(2 * mp.pi) / 2)


def dl(z):
    z = mp.mpf(z)
    return mp.digamma(z) - mp.log(z)


def Qf(s):  # cv / w as a function of s = w^2
    if s == 0:
        return mp.mpf(1) / 12
    w = mp.sqrt(s)
    return cv(1 / w) / w


def Rf(s):  # -(dl + w/2) / s
    if s == 0:
        return mp.mpf(1) / 12
    w = mp.sqrt(s)
    return -(dl(1 / w) + w / 2) / s


def cheb_fit(f, S, deg):
    """polynomial (monomial coefficients in s, ascending) interpolating f at Chebyshev nodes of [0, S]"""
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


def horner32(c, s):
    acc = np.full_like(s, np.float32(c[-1]))
    for k in range(len(c) - 2, -1, -1):
        acc = acc * s + np.float32(c[k])
    return acc


def main():
    zmin = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    dq = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    dr = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    S = mp.mpf(1) / mp.mpf(zmin) ** 2
    cq = cheb_fit(Qf, S, dq)
    cr = cheb_fit(Rf, S, dr)
    zs = np.concatenate([np.linspace(zmin, 8, 4000), np.geomspace(8, 1e6, 2000)]).astype(np.float32)
    w = (np.float32(1) / zs).astype(np.float32)
    s = (w * w).astype(np.float32)
    cv32 = w * horner32([float(c) for c in cq], s)
    dl32 = np.float32(-0.5) * w - s * horner32([float(c) for c in cr], s)
    ecv = max(abs(float(cv32[i]) - float(cv(float(zs[i])))) for i in range(0, len(zs), 7))
    edl = max(abs(float(dl32[i]) - float(dl(float(zs[i])))) for i in range(0, len(zs), 7))
    print(f"zmin={zmin} degQ={dq} degR={dr}: max abs err cv {ecv:.2e} (cv(zmin)={float(cv(zmin)):.4f}), dl {edl:.2e} (dl(zmin)={float(dl(zmin)):.4f})")
    print("Q:", ", ".join(f"{float(c):.9e}f" for c in cq))
    print("R:", ", ".join(f"{float(c):.9e}f" for c in cr))


if __name__ == "__main__":
    main()
