"""The float form of torch's boundary series (`_beta_grad_alpha_small`) that the alpha kernel's per-warp queue evaluates
(csrc/bean_rng.cuh: beta_grad_alpha_small<float> -- MUFU reciprocals, ln x and (1 - x)^-beta through lg2 / ex2, ten unrolled
terms), restated in numpy float32 and compared with `torch._dirichlet_grad` in double over the regime it is used in.  (CPU test
of the FORMULA with exactly rounded float32 operations; the CUDA code is checked by the -m gpu parity tests.)

What it pins: the restructured form is as accurate as the straightforward float32 evaluation of torch's expression.  Both lose
up to ~5e-5 where the ten-term series alternates with growing terms (beta ~ 8, x ~ 0.36); that is float32's price for this
regime, not the restructuring's."""
import numpy as np
import torch
from scipy.special import digamma

f = np.float32


def dg(z):
    return f(digamma(np.float64(z)))


def log1p_series(y):  # csrc/bean_math.cuh: log1p_ratio_series
    y = f(y)
    s = f(y / f(f(2) + y))
    t = f(s * s)
    p = f(2.757020617e-01)
    for c in (2.058080059e-01, 2.868322504e-01, 3.999738930e-01, 6.666667629e-01):
        p = f(p * t + f(c))
    return f(f(s * t) * p + f(s + s))


def alpha_small_kernel_form(x, a, b):
    x, a, b = f(x), f(a), f(b)
    f1 = f(f(dg(f(a + f(1))) - dg(f(a + b))) - f(np.log(x)))
    ia, numer = f(f(1) / a), f(1)
    series = f(ia * f1)
    for i in range(1, 11):
        ci = f(i)
        numer = f(numer * f(f(f(ci - b) * x) * f(f(1) / ci)))
        idn = f(f(1) / f(a + ci))
        series = f(f(numer * idn) * f(f(-ci * ia) * idn + f1) + series)
    return float(f(f(x * f(np.exp(f(-b * log1p_series(-x))))) * series))


def alpha_small_plain_float(x, a, b):
    x, a, b = f(x), f(a), f(b)
    f1 = f(f(dg(f(a + f(1))) - dg(f(a + b))) - f(np.log(x)))
    ia, numer = f(f(1) / a), f(1)
    series = f(f(numer * ia) * f1)
    for i in range(1, 11):
        ci = f(i)
        numer = f(numer * f(f(f(ci - b) * x) / ci))
        idn = f(f(1) / f(a + ci))
        series = f(series + f(f(numer * idn) * f(f1 - f(f(ci * ia) * idn))))
    return float(f(f(x * f(np.power(f(f(1) - x), -b))) * series))


def test_float_boundary_series_is_as_accurate_as_the_plain_float_evaluation():
    rng = np.random.default_rng(0)
    worst_kernel = worst_plain = 0.0
    n = 0
    while n < 2500:
        a = float(np.exp(rng.uniform(np.log(1e-4), np.log(50))))
        b = float(np.exp(rng.uniform(np.log(1e-4), np.log(200))))
        x = float(f(np.exp(rng.uniform(np.log(1e-30), np.log(0.5)))))
        total = float(f(a)) + float(f(b))
        if not (x <= 0.5 and total * x * (1 - x) < 2.5):
            continue  # torch takes another regime
        ref = torch._dirichlet_grad(torch.tensor([x], dtype=torch.float64), torch.tensor([float(f(a))], dtype=torch.float64),
                                    torch.tensor([total], dtype=torch.float64)).item()
        if ref == 0.0:
            continue
        worst_kernel = max(worst_kernel, abs(alpha_small_kernel_form(x, a, b) - ref) / abs(ref))
        worst_plain = max(worst_plain, abs(alpha_small_plain_float(x, a, b) - ref) / abs(ref))
        n += 1
    assert worst_kernel <= 1e-4, worst_kernel               # observed 5.4e-5
    assert worst_kernel <= 1.2 * worst_plain + 1e-6, (worst_kernel, worst_plain)   # observed 5.4e-5 both
