"""The cancellation-free single-precision form of torch's saddle-point Dirichlet gradient (csrc/bean_rng.cuh:
SaddlePairF, derivation in its comment), restated in numpy float32 and checked against `torch._dirichlet_grad` in double
over the whole regime it is used in.  (CPU test of the FORMULA; the CUDA code is checked by the -m gpu parity tests.)"""
import numpy as np
import torch

f = np.float32


def series(y, first_k, last_k, coef):
    acc = f(coef(last_k))
    for k in range(last_k - 1, first_k - 1, -1):
        acc = f(acc * y + f(coef(k)))
    return f(acc * y)


def l1(y):  # log1p(y) / y - 1  (small |y|: from the g1 series, log1p(y) / y = 1 - y / 2 (1 + g1) identically)
    y = f(y)
    if abs(y) < 0.3:
        return f(f(-0.5) * y * f(f(1) + g1(y)))
    return f(np.log1p(y) / y - f(1))


def pow_m15_minus1(z):  # (1 + z)^-1.5 - 1 = e (3 + e (3 + e)), e = 1 / sqrt(1 + z) - 1 = -z / (q (1 + q))
    z = f(z)
    q = f(np.sqrt(f(f(1) + z)))
    e = f(-z / f(q * f(f(1) + q)))
    return f(e * f(f(3) + f(e * f(f(3) + e))))


def g1(y):  # 2 (y - log1p(y)) / y^2 - 1
    y = f(y)
    if abs(y) < 0.3:
        return series(y, 3, 14, lambda k: (-2.0 if k % 2 else 2.0) / k)
    return f(f(2) * (y - np.log1p(y)) / (y * y) - f(1))


def saddle_pair_f32(x0, a, b):
    x0, a, b = f(x0), f(a), f(b)
    T = a + b
    m, om = a / T, b / T
    # delta = x0 - a / T as the kernel forms it: (x0 b - (1 - x0) a) / T, error-free products (float32 x float32 is exact in
    # float64: that is what FMA gives), exact split 1 - x0 = hi + lo
    hi = f(f(1) - x0)
    lo = f(f(f(1) - hi) - x0)
    pa = f(hi * a)
    ea = f(np.float64(hi) * np.float64(a) - np.float64(pa))
    num = f(f(np.float64(x0) * np.float64(b) - np.float64(pa)) - ea)
    d = f(f(np.float64(-lo) * np.float64(a) + np.float64(num)) / T)
    u, v = d / m, -d / om
    s = np.sqrt(f(2) * a * b / T)
    stir = ((288 * a * a + 24 * a + 1) * (288 * b * b + 24 * b + 1) * T * T) / (288 * a * a * b * b * (288 * T * T + 24 * T + 1))
    k = f(2) / (s * T ** 4)
    C2a, C2b = k * a * b * b, k * b * a * a
    h = f(0.5) * T / (a * b)
    c1a, c1b = h * (2 * a - b), h * (2 * b - a)
    l1u, l1v = l1(u), l1(v)
    G1 = om * g1(u) + m * g1(v)
    wm1 = pow_m15_minus1(G1)  # G^-1.5 - 1
    D0, D1 = l1u + wm1 + l1u * wm1, l1v + wm1 + l1v * wm1  # (1 + l1)(1 + wm1) - 1
    t0 = C2a * (D0 + c1a * d) / (d * d) - s * (1 + l1u) / a
    t1 = C2b * (D1 - c1b * d) / (d * d) - s * (1 + l1v) / b
    return float(stir * (-x0 / s) * t0), float(stir * (-(1 - x0) / s) * t1)


def test_float_saddle_form_matches_torch_double():
    rng = np.random.default_rng(0)
    worst, n = 0.0, 0
    while n < 3000:
        a = float(np.exp(rng.uniform(np.log(6.2), np.log(5000))))
        b = float(np.exp(rng.uniform(np.log(6.2), np.log(5000))))
        T = a + b
        m, sd = a / T, np.sqrt(a * b / (T + 1)) / T
        z = rng.choice([-1.0, 1.0]) * float(np.exp(rng.uniform(np.log(0.101), np.log(8.0))))
        x = float(np.float32(m + z * sd))
        if not (0 < x < 1):
            continue
        bnd = T * x * (1 - x)
        other = lambda y: (y <= 0.5 and bnd < 2.5) or (y >= 0.5 and bnd < 0.75)
        if other(x) or other(1 - x) or abs(x - m) <= 0.1 * sd:
            continue  # torch uses another regime for one of the two components (the kernel queues such draws)
        xs = torch.tensor([x, 1 - x], dtype=torch.float64)
        c = torch.tensor([a, b], dtype=torch.float64)
        ref = torch._dirichlet_grad(xs, c, c.sum().expand(2)).tolist()
        got = saddle_pair_f32(x, a, b)
        worst = max(worst, abs(got[0] - ref[0]) / abs(ref[0]), abs(got[1] - ref[1]) / abs(ref[1]))
        n += 1
    assert worst < 3e-6, worst  # observed 9e-7 (3e-5 before delta was formed without the rounding of the mean)


def near_mean_coeffs(al, be, T):  # csrc/bean_rng.cuh: near_mean_coeffs, in float32
    al, be, T = f(al), f(be), f(T)
    b2 = f(be * be)
    b3 = f(b2 * be)
    a3 = f(f(al * al) * al)
    c0 = f(al * f(f(f(43) * b3) + f(al * f(f(f(f(3) * f(f(59) + f(f(180) * be))) * b2) + f(f(al * f(453)) * be)))))
    cx = f(f(f(f(47) * b2) * b2) + f(al * f(f(f(f(20) * f(f(16) + f(f(27) * be))) * b3) - f(al * f(f(f(270) * b2) + f(f(al * f(455)) * be))))))
    c1 = f(a3 * f(f(f(1620) * b2) + f(f(al * f(8)) * f(f(f(135) * be) - f(11)))))
    K = f(f(f(f(f(1) + f(f(12) * al)) * f(f(1) + f(f(12) * be))) / f(T * T)) / f(f(f(f(12960) * a3) * b2) * f(f(1) + f(f(12) * T))))
    return f(c0 * K), f(cx * K), f(c1 * K)


def test_near_mean_polynomial_from_per_guide_coefficients_matches_torch_double():
    """|x - mean| <= 0.1 std inside the saddle-point regime: torch's polynomial is linear in x, the kernel evaluates it as
    (k0 + kx x + k1 (1 - x)) / (1 - x) with per-guide coefficients."""
    rng = np.random.default_rng(1)
    worst, n = 0.0, 0
    while n < 3000:
        a = float(np.exp(rng.uniform(np.log(6.2), np.log(5000))))
        b = float(np.exp(rng.uniform(np.log(6.2), np.log(5000))))
        T = a + b
        m, sd = a / T, np.sqrt(a * b / (T + 1)) / T
        x = float(np.float32(m + rng.uniform(-0.099, 0.099) * sd))
        if abs(x - m) > 0.1 * sd or not (0 < x < 1):
            continue
        ref = torch._dirichlet_grad(torch.tensor([x], dtype=torch.float64), torch.tensor([a], dtype=torch.float64),
                                    torch.tensor([T], dtype=torch.float64)).item()
        k0, kx, k1 = near_mean_coeffs(a, b, T)
        o = f(f(1) - f(x))
        got = float(f(f(np.float64(k1) * np.float64(o) + np.float64(f(np.float64(kx) * np.float64(f(x)) + np.float64(k0)))) / o))
        worst = max(worst, abs(got - ref) / abs(ref))
        n += 1
    assert worst < 2e-6, worst  # observed 6e-7 (torch's own polynomial order in float32: 6e-7)
