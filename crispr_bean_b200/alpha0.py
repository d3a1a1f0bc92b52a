"""One-off Dirichlet-Multinomial precision (`a0`) fits that produce kernel inputs.

Host-side mirror of bean/preprocessing/get_alpha0.py:70-128 (guide counts) and
bean/preprocessing/get_pi_alpha0.py:78-152 (reporter allele counts); SURVEY App. A.9.
The reference fits the log-log line with `scipy.optimize.curve_fit(linear, x, y)`; for a linear model
that optimum is the ordinary-least-squares solution, which is computed in closed form here.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

FALLBACK_POPT = (-1.510, 0.7861)      # get_alpha0.py:105
FALLBACK_PI_POPT = (-3.214, 0.9873)   # get_pi_alpha0.py:110


def get_size_factor(X: np.ndarray) -> np.ndarray:
    """Per-sample depth / mean depth (data_class.py:237-251)."""
    sf = np.mean(np.asarray(X), axis=0)  # dtype follows the count matrix, as in the reference (float32 counts -> float32)
    return sf / np.mean(sf)


def _ols_line(x: np.ndarray, y: np.ndarray) -> Tuple[float, float]:
    xm, ym = x.mean(), y.mean()
    b1 = ((x - xm) * (y - ym)).sum() / ((x - xm) ** 2).sum()
    return np.float64(ym - b1 * xm), np.float64(b1)  # numpy scalars, like curve_fit's popt: they promote float32 to float64


def _valid(x: torch.Tensor, y: torch.Tensor):
    ok = ~(torch.isnan(x) | torch.isnan(y) | torch.isinf(x) | torch.isinf(y))
    return x[ok].double().numpy(), y[ok].double().numpy()


def _moments(X, size_factor, sample_mask, allele_axis: bool):
    """Masked depth-normalised mean q and population variance w over replicates."""
    sf = size_factor.clone()
    m = sample_mask.to(sf.dtype)
    if not allele_axis:
        sf[(m == 0) & (sf == 0)] = 1.0  # get_alpha0.py:29-31
    ex = (None, None) if allele_axis else (None,)
    sf_b = sf[(slice(None), slice(None)) + ex]
    m_b = m[(slice(None), slice(None)) + ex]
    cnt = m.sum(axis=0)[(slice(None),) + ex]
    xn = X / sf_b
    q = (xn * m_b).sum(axis=0) / cnt
    w = (((xn - q) ** 2) * m_b).sum(axis=0) / cnt
    return q, w


def get_fitted_alpha0(X, size_factor, sample_mask=None, shrink=False, shrink_prior_var=1.0,
                      popt: Optional[Tuple[float, float]] = None):
    """a0[g] = exp(b0 + b1 log n_g) with (b0, b1) fitted on the method-of-moments a0 (get_alpha0.py:70-119).

    X: (R, B, G) counts.  Returns (a0 (G,) float64 tensor, (b0, b1)).
    """
    R, B, G = X.shape
    if sample_mask is None:
        sample_mask = torch.ones((R, B))
    elif (sample_mask.sum(axis=0) == 0).any():
        raise ValueError("Some bins have no data.")
    q, w = _moments(X + 1, size_factor, sample_mask, allele_axis=False)  # (B, G)
    n = torch.nanmean(q, axis=0) * q.shape[0]
    p = q / n[None, :]
    r = (w - q) / (n[None, :] * p * (1 - p))
    a0 = ((n - 1) / (r - 1 + 1 / (1 - p)) - 1).mean(axis=0)
    x, y = _valid(n.log(), a0.log())
    if len(y) < 5:
        if popt is None:
            popt = FALLBACK_POPT
    else:
        popt = _ols_line(x, y)
    log_a0_est = torch.as_tensor(popt[0] + popt[1] * n.log().numpy())  # dtype: numpy promotion, as the reference
    if shrink:
        yy = a0.log().to(log_a0_est.dtype)
        yy = torch.where(torch.isnan(yy), log_a0_est, yy)
        var = ((yy - log_a0_est) ** 2).sum() / (len(yy) - 1)
        pw = var / (var + shrink_prior_var)
        log_a0_est = pw * log_a0_est + (1 - pw) * yy
    return torch.exp(log_a0_est), tuple(popt)


def get_pred_alpha0(X, size_factor, popt, sample_mask=None):
    """a0 from a pre-fitted trend, n = sum_b q (get_alpha0.py:122-128)."""
    R, B, G = X.shape
    if sample_mask is None:
        sample_mask = torch.ones((R, B))
    q, _ = _moments(X + 1, size_factor, sample_mask, allele_axis=False)
    return torch.as_tensor(np.exp(popt[0] + popt[1] * q.sum(axis=0).log().numpy()))


def get_fitted_pi_alpha0(allele_counts, size_factor, shrink=False, shrink_prior_var=1.0):
    """Precision of the reporter allele Dirichlet, control condition 0 only (get_pi_alpha0.py:78-143).

    allele_counts: (R, C, G, A).  Returns (pi_a0 (G,), (b0, b1)).
    """
    R, C, G, A = allele_counts.shape
    q, w = _moments(allele_counts + 1, size_factor, torch.ones((R, C)), allele_axis=True)
    q, w = q[0], w[0]  # (G, A)
    n = q.sum(-1)
    p = q / n[:, None]
    r = (w - q) / (n[:, None] * p * (1 - p))
    a0 = torch.nanmean((n[:, None] - 1) / (r - 1 + 1 / (1 - p)) - 1, axis=-1)
    x, y = _valid(n.log(), a0.log())
    popt = FALLBACK_PI_POPT if len(y) < 10 else _ols_line(x, y)
    log_a0_est = torch.as_tensor(popt[0] + popt[1] * n.log().numpy())  # dtype: numpy promotion, as the reference
    if shrink:
        yy = a0.log().to(log_a0_est.dtype)
        yy = torch.where(torch.isnan(yy), log_a0_est, yy)
        var = (yy - log_a0_est) ** 2 / len(yy)  # get_pi_alpha0.py:66-67 (elementwise, as the reference)
        pw = var / (var + shrink_prior_var)
        log_a0_est = pw * log_a0_est + (1 - pw) * yy
    return torch.exp(log_a0_est), tuple(popt)


def get_pred_pi_alpha0(allele_counts, size_factor, popt):
    R, C, G, A = allele_counts.shape
    q, _ = _moments(allele_counts + 1, size_factor, torch.ones((R, C)), allele_axis=True)
    return torch.as_tensor(np.exp(popt[0] + popt[1] * q[0].sum(axis=-1).log().numpy()))
