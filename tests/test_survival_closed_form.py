"""The kernel-shaped closed form of the survival MixtureNormal step (oracle/survival_closed_form.py) against the autograd
oracle, which is pinned to the reference's survival_model.py -- on a synthetic screen, on the reference-computed golden
case, and on the reference's own survival_var_mini_screen (loss 1e-11, gradients 1e-8 relative, float64)."""
import numpy as np
import pytest
import torch

from crispr_bean_b200 import data_class as dc
from crispr_bean_b200.synth import make_survival_screen
from oracle import bean_oracle as O
from oracle.survival_closed_form import survival_mixture_step
from tests.helpers import cast_data, default_dtype
from tests.test_reference_golden import group, load_case


def oracle_eval(data, noise, **kw):
    with default_dtype(torch.float64):
        ps = O.ParamStore()
        loss, _ = O.elbo_survival_mixture_normal(data, ps, noise=noise, **kw)
        loss.backward()
    grads = {k: v.grad.detach().double().numpy() for k, v in ps.unconstrained.items()}
    theta = {k: v.detach().clone() for k, v in ps.unconstrained.items()}
    return float(loss.detach()), grads, theta


def check(data, noise, **kw):
    data = cast_data(data, torch.float64)
    loss, grads, theta = oracle_eval(data, noise, **kw)
    got_loss, got = survival_mixture_step(data, theta, noise, **kw)
    assert abs(got_loss - loss) <= 1e-11 * abs(loss), (got_loss, loss)
    assert set(got) == set(grads)
    for k, g in grads.items():
        err = np.abs(got[k].reshape(g.shape) - g).max() / max(np.abs(g).max(), 1e-300)
        assert err <= 1e-8, (k, err)


def random_noise(data, seed):
    g = torch.Generator().manual_seed(seed)
    G, R, T = data.n_guides, data.n_reps, data.n_targets
    gam = torch._standard_gamma(torch.full((R, G), 1.3, dtype=torch.float64), generator=g)
    pig = torch._standard_gamma(torch.full((R, 1, G, 2), 0.7, dtype=torch.float64), generator=g)
    return {"eps_mu": torch.randn((T, 1), generator=g, dtype=torch.float64), "eps_negctrl": torch.randn((G,), generator=g, dtype=torch.float64),
            "q0": gam / gam.sum(-1, keepdim=True), "pi": pig / pig.sum(-1, keepdim=True)}


@pytest.mark.parametrize("use_bcmatch", [True, False])
def test_synthetic_screen(use_bcmatch):
    scr = make_survival_screen(14, "lognormal", n_reps=3, seed=12, n_negctrl_guides=5, depth=80.0)
    data = dc.VariantSurvivalReporterScreenData(scr, control_condition="D7")
    data.repguide_mask[0, ::4] = False  # rows outside the replicate x guide mask
    check(data, random_noise(data, 1), use_bcmatch=use_bcmatch, mu_negctrl=(0.02, 0.3))


@pytest.mark.parametrize("name", ["survival_mixture", "survival_mixture_control_d0", "survival_real_var_mixture"])
def test_reference_golden_cases(name):
    """Same draws as the reference run behind the fixture: the closed form lands on the reference's own loss too."""
    z, data = load_case(name)
    noise = {k: torch.as_tensor(v) for k, v in group(z, "f64/noise/").items() if "/" not in k}
    check(data, noise)
    data64 = cast_data(data, torch.float64)
    _, _, theta = oracle_eval(data64, noise)
    loss, _ = survival_mixture_step(data64, theta, noise)
    assert abs(loss - float(z["f64/loss"])) <= 1e-10 * abs(float(z["f64/loss"]))


@pytest.mark.parametrize("name", ["survival_normal", "survival_normal_no_negctrl_idx", "survival_normal_bcmatch", "survival_real_var_normal"])
def test_survival_normal_closed_form_on_golden_cases(name):
    """`--uniform-edit` survival program: the Dirichlet draw over all guides feeds the likelihood (third library-wide sum)."""
    import ast

    from oracle.survival_closed_form import survival_normal_step

    z, data = load_case(name)
    kw = ast.literal_eval(str(z["meta/oracle_kwargs"]))
    noise = {k: torch.as_tensor(v) for k, v in group(z, "f64/noise/").items() if "/" not in k}
    data = cast_data(data, torch.float64)
    with default_dtype(torch.float64):
        ps = O.ParamStore()
        loss, _ = O.elbo_survival_normal(data, ps, noise=noise, **kw)
        loss.backward()
    theta = {k: v.detach().clone() for k, v in ps.unconstrained.items()}
    got_loss, got = survival_normal_step(data, theta, noise, **kw)
    assert abs(got_loss - float(loss.detach())) <= 1e-11 * abs(float(loss.detach()))
    assert abs(got_loss - float(z["f64/loss"])) <= 1e-10 * abs(float(z["f64/loss"]))
    for k, v in ps.unconstrained.items():
        g = v.grad.detach().double().numpy()
        err = np.abs(got[k].reshape(g.shape) - g).max() / max(np.abs(g).max(), 1e-300)
        assert err <= 1e-8, (k, err)
