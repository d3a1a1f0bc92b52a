"""The kernel-shaped closed form of the tiling (sorting MultiMixtureNormal) step (oracle/tiling_closed_form.py) against the
autograd oracle pinned to the reference's model.py -- synthetic screens, the reference-computed golden cases and the
reference's own tiling_mini_screen with 231 alleles per guide (loss 1e-11, gradients 1e-8 relative, float64)."""
import numpy as np
import pytest
import torch

from crispr_bean_b200 import data_class as dc
from crispr_bean_b200.synth import make_tiling_screen
from oracle import bean_oracle as O
from oracle.tiling_closed_form import tiling_step
from tests.helpers import cast_data, default_dtype
from tests.test_reference_golden import edit_perm, group, load_case, to_ours


def check(data, noise, ref_loss=None, **kw):
    data = cast_data(data, torch.float64)
    with default_dtype(torch.float64):
        ps = O.ParamStore()
        loss, _ = O.elbo_multi_mixture_normal(data, ps, noise=noise, **kw)
        loss.backward()
    theta = {k: v.detach().clone() for k, v in ps.unconstrained.items()}
    got_loss, got = tiling_step(data, theta, noise, **kw)
    loss = float(loss.detach())
    assert abs(got_loss - loss) <= 1e-11 * abs(loss), (got_loss, loss)
    if ref_loss is not None:
        assert abs(got_loss - ref_loss) <= 1e-10 * abs(ref_loss)
    assert set(got) == set(ps.unconstrained)
    for k, v in ps.unconstrained.items():
        g = v.grad.detach().double().numpy()
        err = np.abs(got[k].reshape(g.shape) - g).max() / max(np.abs(g).max(), 1e-300)
        assert err <= 1e-8, (k, err)


@pytest.mark.parametrize("max_alleles,use_bcmatch", [(5, True), (9, False)])
def test_synthetic_screen(max_alleles, use_bcmatch):
    data = dc.TilingSortingReporterScreenData(make_tiling_screen(n_guides=30, n_reps=3, max_alleles=max_alleles, seed=max_alleles),
                                              control_can_be_selected=True, allele_df_key="allele_counts")
    data.repguide_mask[1, ::5] = False
    g = torch.Generator().manual_seed(4)
    E, G, R, A = data.n_edits, data.n_guides, data.n_reps, data.n_max_alleles
    gam = torch._standard_gamma(torch.full((R, 1, G, A), 0.6, dtype=torch.float64), generator=g).clamp(min=1e-300)
    gam = torch.where(data.allele_mask[None, None], gam, torch.full_like(gam, 1e-300))  # draws of non-existent alleles ~ 0
    noise = {"eps_mu": torch.randn(E, generator=g, dtype=torch.float64), "eps_sd": torch.randn(E, generator=g, dtype=torch.float64),
             "pi": gam / gam.sum(-1, keepdim=True)}
    check(data, noise, use_bcmatch=use_bcmatch)


@pytest.mark.parametrize("name", ["tiling_small", "tiling_wide", "tiling_real_mini"])
def test_reference_golden_cases(name):
    z, data = load_case(name)
    perm = edit_perm(z, data)
    noise = {k: torch.as_tensor(to_ours(v, perm, k)) for k, v in group(z, "f64/noise/").items() if "/" not in k}
    check(data, noise, ref_loss=float(z["f64/loss"]))


@pytest.mark.parametrize("name", ["survival_tiling", "survival_tiling_real_mini"])
def test_survival_tiling_closed_form_on_golden_cases(name):
    from oracle.tiling_closed_form import survival_tiling_step

    z, data = load_case(name)
    perm = edit_perm(z, data)
    noise = {k: torch.as_tensor(to_ours(v, perm, k)) for k, v in group(z, "f64/noise/").items() if "/" not in k}
    data = cast_data(data, torch.float64)
    with default_dtype(torch.float64):
        ps = O.ParamStore()
        loss, _ = O.elbo_survival_multi_mixture_normal(data, ps, noise=noise)
        loss.backward()
    theta = {k: v.detach().clone() for k, v in ps.unconstrained.items()}
    got_loss, got = survival_tiling_step(data, theta, noise)
    assert abs(got_loss - float(loss.detach())) <= 1e-11 * abs(float(loss.detach()))
    assert abs(got_loss - float(z["f64/loss"])) <= 1e-10 * abs(float(z["f64/loss"]))
    for k, v in ps.unconstrained.items():
        g = (v.grad if v.grad is not None else torch.zeros_like(v)).detach().double().numpy()
        err = np.abs(got[k].reshape(g.shape) - g).max() / max(np.abs(g).max(), 1e-300)
        assert err <= 1e-8, (k, err)
