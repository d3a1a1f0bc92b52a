"""`write_result_table`: fitted parameters -> `bean_element_result.<model>.csv` / `bean_sgRNA_result.<model>.csv`.

Host-side mirror of bean/model/readwrite.py:49-215 with the same signature, column names, column order and row
order (sorted by |z|), so the output schema of `bean run` is unchanged.  Not on the hot path:
it exists so that end-to-end parity -- identical variant ranking -- can be checked through the reference's own output
format.  Written vectorised: the non-overlap score, which the reference computes row by row with
`statistics.NormalDist.overlap`, is the closed form of that overlap coefficient evaluated on whole columns.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np
import pandas as pd
from scipy.special import expit, logit, ndtr, ndtri


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def non_overlap(mu, sigma):
    """1 - overlap coefficient of N(mu, sigma) with N(0, 1), elementwise (readwrite.py:9-16).

    Two normal densities cross where a quadratic vanishes; the overlap is the total mass under the lower curve
    (Inman & Bradley 1989).  Equal variances leave a single crossing: OVL = 2 Phi(-|mu| / 2 sigma)."""
    mu, sigma = np.asarray(mu, dtype=np.float64), np.asarray(sigma, dtype=np.float64)
    # order the pair as (X, Y) with (sigma_X, mu_X) <= (sigma_Y, mu_Y), like NormalDist.overlap
    swap = (sigma < 1.0) | ((sigma == 1.0) & (mu < 0.0))
    mx, sx = np.where(swap, mu, 0.0), np.where(swap, sigma, 1.0)
    my, sy = np.where(swap, 0.0, mu), np.where(swap, 1.0, sigma)
    vx, vy = sx * sx, sy * sy
    dv, dm = vy - vx, np.abs(my - mx)
    same = dv == 0
    with np.errstate(divide="ignore", invalid="ignore"):
        a = mx * vy - my * vx
        b = sx * sy * np.sqrt(dm * dm + dv * np.log(vy / vx))
        x1, x2 = (a + b) / dv, (a - b) / dv
        two = 1.0 - (np.abs(ndtr((x1 - my) / sy) - ndtr((x1 - mx) / sx)) + np.abs(ndtr((x2 - my) / sy) - ndtr((x2 - mx) / sx)))
        from scipy.special import erf

        one = 1.0 - erf(dm / (2.0 * sx * np.sqrt(2.0)))
    return 1.0 - np.where(same, one, two)


def _credible_interval(df, mu_col, sd_col, alpha=0.05):
    df = df.copy()
    mu, sd = df[mu_col].to_numpy(dtype=np.float64), df[sd_col].to_numpy(dtype=np.float64)
    df[f"CI[{alpha / 2}"] = mu + sd * ndtri(alpha / 2)
    df[f"{1 - alpha / 2}]"] = mu + sd * ndtri(1 - alpha / 2)
    return df


def _adjust_by_control(df, sd0, suffix, mu_col, sd_col, mu0=0.0):
    df[f"mu{suffix}"] = df[mu_col] - mu0
    df[f"mu_sd{suffix}"] = df[sd_col] * sd0
    df[f"mu_z{suffix}"] = df[f"mu{suffix}"] / df[f"mu_sd{suffix}"]
    df[f"novl{suffix}"] = non_overlap(df[f"mu{suffix}"], df[f"mu_sd{suffix}"])
    return df


def scale_pi(pi, guide_acc, fitted_noise_logit=None, a=0.2513, b=-1.9458):
    """Editing rate scaled by accessibility (+ fitted logit noise), readwrite.py:218-247."""
    scaled = pi * np.exp(b) * np.asarray(guide_acc) ** a
    if fitted_noise_logit is None:
        return scaled
    return expit(logit(scaled.clip(min=1e-3, max=1 - 1e-3)) + fitted_noise_logit).clip(min=1e-3, max=1 - 1e-3)


def write_result_table(target_info_df: pd.DataFrame, guide_info_df: pd.DataFrame, param_hist_dict, model_label: str,
                       prefix: str = "", suffix: str = "", negctrl_params=None,
                       adjust_confidence_by_negative_control: bool = True,
                       adjust_confidence_negatives: Optional[np.ndarray] = None, guide_acc: Optional[Sequence] = None,
                       sd_is_fitted: bool = True, sample_covariates: Optional[List[str]] = None,
                       return_result: bool = False, is_survival_screen: bool = False) -> Union[pd.DataFrame, None]:
    """Same arguments, files and return value as the reference's `write_result_table`."""
    mu_loc = param_hist_dict["mu_loc"]
    if mu_loc.dim() not in (1, 2):
        raise ValueError(f"`mu_loc` has invalid shape of {mu_loc.shape}")
    col = (lambda t: _np(t)[:, 0]) if mu_loc.dim() == 2 else _np
    mu, mu_sd = col(param_hist_dict["mu_loc"]), col(param_hist_dict["mu_scale"])
    cols = {"mu": mu, "mu_sd": mu_sd, "mu_z": mu / mu_sd}
    if sd_is_fitted:
        sd = col(param_hist_dict["sd_loc"].detach().exp())
        cols["sd"] = sd
    if sample_covariates is not None:
        cov_loc, cov_scale = _np(param_hist_dict["mu_cov_loc"]), _np(param_hist_dict["mu_cov_scale"])
        for i, name in enumerate(sample_covariates):
            cols[f"mu_{name}"] = mu + cov_loc[i]
            cols[f"mu_sd_{name}"] = np.sqrt(mu_sd ** 2 + cov_scale[i] ** 2)
            cols[f"mu_z_{name}"] = cols[f"mu_{name}"] / cols[f"mu_sd_{name}"]
    fit = pd.DataFrame(cols)
    if negctrl_params is not None:  # centre / scale by the shared negative-control fit (cli/run.py:236-257)
        mu0 = _np(negctrl_params["mu_loc"]).mean()
        sd0 = _np(negctrl_params["sd_loc"].detach().exp()) if sd_is_fitted else 1.0
        fit["mu_scaled"] = (mu - mu0) / sd0
        fit["mu_sd_scaled"] = mu_sd / sd0
        fit["mu_z_scaled"] = fit.mu_scaled / fit.mu_sd_scaled
        if sd_is_fitted:
            fit["sd_scaled"] = sd / sd0
        fit["novl_scaled"] = non_overlap(fit["mu_scaled"], fit["mu_sd_scaled"])
        if sample_covariates is not None:
            for name in sample_covariates:
                fit[f"mu_{name}_scaled"] = (fit[f"mu_{name}"] - mu0) / sd0
                fit[f"mu_sd_{name}_scaled"] = fit[f"mu_sd_{name}"] / sd0
                fit[f"mu_z_{name}_scaled"] = fit[f"mu_{name}_scaled"] / fit["mu_sd_scaled"]
    fit = pd.concat([target_info_df.reset_index(), fit.reset_index(drop=True)], axis=1)

    enough = adjust_confidence_by_negative_control and adjust_confidence_negatives is not None \
        and len(adjust_confidence_negatives) >= 10
    if adjust_confidence_by_negative_control and adjust_confidence_negatives is None:
        raise AssertionError("adjust_confidence_negatives is required")
    if enough:
        # z-scores of the negative-control variants should be N(0, 1): rescale every sd by their zero-centred spread
        nc = fit.iloc[adjust_confidence_negatives]
        z = (nc.mu_z_scaled if "mu_z_scaled" in nc.columns else nc.mu_z).to_numpy(dtype=np.float64)
        std = float(np.sqrt(np.mean(z ** 2)))  # scipy norm.fit(z, floc=0): MLE of the scale with the location fixed at 0
        scaled = "negctrl" in param_hist_dict.keys()
        fit = _adjust_by_control(fit, std, "_adj", "mu_scaled" if scaled else "mu", "mu_sd_scaled" if scaled else "mu_sd", mu0=0.0)
        fit = _credible_interval(fit, "mu_adj", "mu_sd_adj")
        fit = fit.iloc[(-fit.mu_z_adj.abs()).argsort()]
        if sample_covariates is not None:
            for name in sample_covariates:
                fit = _adjust_by_control(fit, std, f"_{name}_adj", f"mu_{name}_scaled" if scaled else f"mu_{name}",
                                         f"mu_sd_{name}_scaled" if scaled else f"mu_sd_{name}")
                fit = _credible_interval(fit, f"mu_{name}_adj", f"mu_sd_{name}_adj")
    else:
        fit = _credible_interval(fit, "mu", "mu_sd")
        fit = fit.iloc[(-fit.mu_z.abs()).argsort()]

    if guide_acc is not None:
        a_fit = _np(param_hist_dict["alpha_pi"]) if "alpha_pi" in param_hist_dict.keys() else None
        pi = 1.0 if a_fit is None else a_fit[..., 1:].sum(axis=1) / a_fit.sum(axis=1)
        guide_info_df.insert(1, "accessibility", guide_acc)
        noise = _np(param_hist_dict["noise_scale"]) if "noise_scale" in param_hist_dict.keys() else None
        guide_info_df.insert(2, "scaled_edit_eff", scale_pi(pi, guide_acc, noise))
    guide_info_df.to_csv(f"{prefix}bean_sgRNA_result.{model_label}{suffix}.csv")
    if return_result:
        return fit
    fit.to_csv(f"{prefix}bean_element_result.{model_label}{suffix}.csv")
    return None
