"""Shared test helpers: fixtures, fixed reparameterisation noise, oracle drivers."""
from __future__ import annotations

import contextlib
import copy
import os

import numpy as np
import pandas as pd
import torch

from crispr_bean_b200.data_class import VariantSortingReporterScreenData, VariantSortingScreenData
from crispr_bean_b200.screen import MiniScreen
from crispr_bean_b200.synth import make_sorting_screen
from oracle import bean_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@contextlib.contextmanager
def default_dtype(dtype):
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        yield
    finally:
        torch.set_default_dtype(old)


def cast_data(data, dtype):
    """Shallow copy of a *ScreenData with every floating tensor cast to `dtype`."""
    nd = copy.copy(data)
    for k, v in vars(data).items():
        if torch.is_tensor(v) and v.is_floating_point():
            setattr(nd, k, v.to(dtype))
    return nd


def var_mini_screen() -> MiniScreen:
    z = np.load(os.path.join(GOLDEN, "var_mini.npz"))
    guides = pd.DataFrame({"target": z["target"], "target_group": z["target_group"]},
                          index=pd.Index(z["guide_names"], name="name"))
    samples = pd.DataFrame({"condition": z["condition"], "replicate": z["replicate"],
                            "lower_quantile": z["lower_quantile"], "upper_quantile": z["upper_quantile"]},
                           index=pd.Index(z["sample_names"]))
    return MiniScreen(z["counts"], guides, samples)


def load_var_mini():
    """c1: the reference's CSV fixture through the tensoriser (Normal model path, no reporter layers)."""
    scr = var_mini_screen()
    scr = scr[np.argsort(scr.guides["target"].to_numpy(), kind="stable"), :]  # prepare_bdata: sort by target
    return VariantSortingScreenData(scr, condition_column="condition", control_condition="bulk",
                                    control_can_be_selected=True)


def make_small_mixture_data(n_variants=12, n_reps=3, seed=3, with_bulk_bin=True, **kw):
    scr = make_sorting_screen(n_variants, 4, n_reps=n_reps, seed=seed, n_negctrl_guides=6, depth=kw.pop("depth", 120.0), **kw)
    return VariantSortingReporterScreenData(scr, control_can_be_selected=with_bulk_bin,
                                            accessibility_col="accessibility" if kw.get("accessibility") else None)


def fixed_noise(model: str, data, seed=0, dtype=torch.float64):
    """Reparameterisation noise injected identically into oracle and kernel."""
    g = torch.Generator().manual_seed(seed)
    G, R = data.n_guides, data.n_reps
    if model == "ControlNormal":
        return {"eps_mu": torch.randn((), generator=g, dtype=dtype), "eps_sd": 0.1 * torch.randn((), generator=g, dtype=dtype)}
    if model == "MultiMixtureNormal":
        E = data.n_edits
        noise = {"eps_mu": torch.randn((E,), generator=g, dtype=dtype), "eps_sd": 0.1 * torch.randn((E,), generator=g, dtype=dtype)}
        A = data.n_max_alleles
    else:
        T = data.n_targets
        noise = {"eps_mu": torch.randn((T, 1), generator=g, dtype=dtype), "eps_sd": 0.1 * torch.randn((T, 1), generator=g, dtype=dtype)}
        A = 2
    if model in ("MixtureNormal", "MultiMixtureNormal"):
        gam = torch._standard_gamma(torch.full((R, 1, G, A), 1.5, dtype=dtype), generator=g)
        if model == "MultiMixtureNormal":
            gam = torch.where(data.allele_mask[None, None], gam, torch.full_like(gam, 1e-30))
        noise["pi"] = (gam / gam.sum(-1, keepdim=True)).clamp(min=1e-30)
        noise["eps_noise"] = torch.randn((G,), generator=g, dtype=dtype)
    return noise


def oracle_loss_and_grads(model: str, data, noise, dtype=torch.float64, **model_kwargs):
    """-ELBO and its gradient w.r.t. the UNCONSTRAINED parameters, from the CPU oracle."""
    with default_dtype(dtype):
        d = cast_data(data, dtype)
        n = {k: v.to(dtype) for k, v in noise.items()} if noise is not None else None
        ps = O.ParamStore()
        loss, aux = O.SORTING_ELBOS[model](d, ps, noise=n, **model_kwargs)
        loss.backward()
        grads = {k: v.grad.detach().clone() for k, v in ps.unconstrained.items()}
    return {"loss": float(loss.detach()), "grads": grads, "aux": {k: (v.detach() if torch.is_tensor(v) else v) for k, v in aux.items()},
            "params": ps}


def oracle_ll_core(data, mu_alleles, sd_alleles, pi, dtype=torch.float64, use_bcmatch=True, mask_thres=10,
                   allele_mask=None):
    """Count log-likelihood + autograd gradients w.r.t. (mu, sd, pi) from the oracle op chain."""
    with default_dtype(dtype):
        d = cast_data(data, dtype)
        mu = mu_alleles.detach().to(dtype).clone().requires_grad_(True)
        sd = sd_alleles.detach().to(dtype).clone().requires_grad_(True)
        A = mu.shape[1]
        if pi is None:
            p = torch.ones((d.n_reps, 1, d.n_guides, A), dtype=dtype)
        else:
            p = pi.detach().to(dtype).clone().requires_grad_(True)
        total, out = O.sorting_ll_core(d, mu, sd, p, use_bcmatch=use_bcmatch, mask_thres=mask_thres, allele_mask=allele_mask)
        total.backward()
    return {"ll": float(total.detach()), "d_mu": mu.grad, "d_sd": sd.grad, "d_pi": p.grad if pi is not None else None, "aux": out}
