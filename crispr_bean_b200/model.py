"""Model / guide entry points with the reference's names (bean/model/model.py).

The reference's functions are pyro programs executed op by op; here a model/guide pair is a DESCRIPTOR
that `run_inference` lowers onto the fused CUDA step (`svi.SviEngine`), so the call sites of
bean/cli/run.py:94, :253-257, :272-274 stay the same:

    model_label, model, guide = identify_model_guide(args)
    param_store, history = run_inference(model, guide, ndata, num_steps=args.n_iter)

Keyword arguments keep the reference's names, defaults and meaning.
"""
from __future__ import annotations

from functools import partial
from typing import Optional


class _Program:
    """A named model or guide program plus the keyword arguments bound so far."""

    def __init__(self, name: str, kind: str, defaults: dict, selection: str = "sorting"):
        self.bean_name, self.kind, self.defaults, self.selection = name, kind, defaults, selection
        self.__name__ = f"{name}{'Model' if kind == 'model' else 'Guide'}"

    def __call__(self, data, **kwargs):
        raise RuntimeError(
            f"{self.__name__} is lowered onto the CUDA SVI step by crispr_bean_b200.run.run_inference; "
            "it is not a pyro program and cannot be traced directly.")


def resolve(program):
    """(name, kwargs) of a program or a functools.partial of one."""
    kwargs = {}
    while isinstance(program, partial):
        kwargs = {**program.keywords, **kwargs}
        program = program.func
    if not isinstance(program, _Program):
        raise TypeError(f"not a crispr_bean_b200 model/guide: {program!r}")
    return program.bean_name, {**program.defaults, **kwargs}


def selection_of(program) -> str:
    """"sorting" | "survival": which family (bean/model/model.py vs survival_model.py) a program belongs to."""
    while isinstance(program, partial):
        program = program.func
    return program.selection


# sorting models (bean/model/model.py:19, :168, :378, :550) and guides (:754, :785, :861, :878)
NormalModel = _Program("Normal", "model", dict(mask_thres=10, use_bcmatch=True, sd_scale=0.01, prior_params=None))
ControlNormalModel = _Program("ControlNormal", "model", dict(mask_thres=10, use_bcmatch=True))
MixtureNormalModel = _Program("MixtureNormal", "model", dict(
    alpha_prior=1, use_bcmatch=True, sd_scale=0.01, scale_by_accessibility=False, fit_noise=False, prior_params=None))
MultiMixtureNormalModel = _Program("MultiMixtureNormal", "model", dict(
    alpha_prior=1, use_bcmatch=True, sd_scale=0.01, scale_by_accessibility=False, fit_noise=False, prior_params=None,
    epsilon=1e-5))
NormalGuide = _Program("Normal", "guide", {})
ControlNormalGuide = _Program("ControlNormal", "guide", dict(mask_thres=10, use_bcmatch=True))
MixtureNormalGuide = _Program("MixtureNormal", "guide", dict(
    alpha_prior=1, use_bcmatch=True, scale_by_accessibility=False, fit_noise=False))
MultiMixtureNormalGuide = _Program("MultiMixtureNormal", "guide", dict(
    alpha_prior=1, use_bcmatch=True, epsilon=1e-5, scale_by_accessibility=False, fit_noise=False))
