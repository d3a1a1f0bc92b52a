"""Survival model / guide entry points with the reference's names (bean/model/survival_model.py).

Descriptors like `crispr_bean_b200.model`: `run_inference` lowers a (model, guide) pair onto `survival.SurvivalSviEngine`.
Keyword arguments keep the reference's names, defaults and meaning (survival_model.py:15-21, :133, :215-226, :629-765).
"""
from .model import _Program

NormalModel = _Program("Normal", "model", dict(mask_thres=10, use_bcmatch=True, prior_params=None, mu_negctrl=0.0),
                       selection="survival")
ControlNormalModel = _Program("ControlNormal", "model", dict(mask_thres=10, use_bcmatch=True), selection="survival")
MixtureNormalModel = _Program("MixtureNormal", "model", dict(
    alpha_prior=1, use_bcmatch=True, use_all_timepoints_for_pi=True, sd_scale=0.01, scale_by_accessibility=False,
    fit_noise=False, mask_thres=10, prior_params=None, mu_negctrl=(0.0, 0.1)), selection="survival")
NormalGuide = _Program("Normal", "guide", {}, selection="survival")
ControlNormalGuide = _Program("ControlNormal", "guide", dict(mask_thres=10, use_bcmatch=True), selection="survival")
MixtureNormalGuide = _Program("MixtureNormal", "guide", dict(
    alpha_prior=1, use_bcmatch=True, scale_by_accessibility=False, fit_noise=False), selection="survival")
MultiMixtureNormalModel = _Program("MultiMixtureNormal", "model", dict(
    alpha_prior=1, use_bcmatch=True, use_all_timepoints_for_pi=True, sd_scale=0.01, norm_pi=False, scale_by_accessibility=False,
    fit_noise=False, prior_params=None, epsilon=1e-5, mu_negctrl=(0.0, 0.1)), selection="survival")
MultiMixtureNormalGuide = _Program("MultiMixtureNormal", "guide", dict(
    alpha_prior=1, use_bcmatch=True, epsilon=1e-5, scale_by_accessibility=False, fit_noise=False), selection="survival")
