"""Pin the CPU oracle (the reference's own tests assert no numbers for this path -- SURVEY section 8c).

Independent implementations used as known answers: scipy.stats (Dirichlet-Multinomial pmf, Normal cdf),
torch.distributions / torch.optim (the reference's real dependencies), torch.autograd.gradcheck.
"""
import json
import math
import os

import numpy as np
import pytest
import scipy.stats as st
import torch

from oracle import bean_oracle as O
from tests import helpers as H


def test_dm_log_prob_matches_scipy():
    rng = np.random.default_rng(0)
    for _ in range(50):
        B = int(rng.integers(2, 7))
        alpha = rng.gamma(2.0, 5.0, size=B) + 1e-3
        x = rng.integers(0, 400, size=B)
        want = st.dirichlet_multinomial.logpmf(x, alpha, int(x.sum()))
        got = O.dm_log_prob(torch.tensor(alpha), torch.tensor(x, dtype=torch.float64))
        assert abs(float(got) - want) <= 1e-10 * max(1.0, abs(want))


def test_dm_log_prob_tiny_alpha_and_zero_counts():
    alpha = torch.tensor([1e-5, 3.0, 1e-5, 7.5], dtype=torch.float64)
    x = torch.tensor([0.0, 12.0, 0.0, 30.0], dtype=torch.float64)
    want = st.dirichlet_multinomial.logpmf(x.numpy().astype(int), alpha.numpy(), 42)
    assert abs(float(O.dm_log_prob(alpha, x)) - want) < 1e-9


def test_std_normal_prob_matches_scipy_and_bulk_is_one():
    uq = torch.tensor([0.2, 0.4, 0.8, 1.0, 1.0], dtype=torch.float64)
    lq = torch.tensor([0.0, 0.2, 0.6, 0.0, 0.8], dtype=torch.float64)
    G, A = 7, 2
    g = torch.Generator().manual_seed(1)
    mu = torch.randn((G, A), generator=g, dtype=torch.float64)
    sd = torch.rand((G, A), generator=g, dtype=torch.float64) + 0.5
    P = O.get_std_normal_prob(uq[:, None, None].expand(-1, G, A), lq[:, None, None].expand(-1, G, A),
                              mu[None].expand(5, -1, -1), sd[None].expand(5, -1, -1))
    tu = st.norm.ppf(uq.numpy())
    tl = st.norm.ppf(lq.numpy())
    want = st.norm.cdf((tu[:, None, None] - mu.numpy()[None]) / sd.numpy()[None]) - st.norm.cdf(
        (tl[:, None, None] - mu.numpy()[None]) / sd.numpy()[None])
    np.testing.assert_allclose(P.numpy(), want, rtol=0, atol=1e-14)
    assert (P[3] == 1.0).all()  # the (0, 1) bulk pseudo-bin is exactly 1
    # standard-normal allele: bin mass equals the quantile width
    P0 = O.get_std_normal_prob(uq, lq, torch.zeros(5, dtype=torch.float64), torch.ones(5, dtype=torch.float64))
    np.testing.assert_allclose(P0.numpy(), (uq - lq).numpy(), atol=1e-15)


def test_std_normal_prob_allele_mask():
    uq = torch.tensor([0.5, 1.0], dtype=torch.float64)[:, None, None].expand(-1, 3, 2)
    lq = torch.tensor([0.0, 0.5], dtype=torch.float64)[:, None, None].expand(-1, 3, 2)
    mask = torch.tensor([[True, True], [True, False], [True, True]])[None].expand(2, -1, -1)
    P = O.get_std_normal_prob(uq, lq, torch.zeros((2, 3, 2), dtype=torch.float64), torch.ones((2, 3, 2), dtype=torch.float64), mask=mask)
    assert (P[:, 1, 1] == 0).all() and torch.allclose(P[:, 0, :], torch.full((2, 2), 0.5, dtype=torch.float64))


def test_get_alpha_hand_computed():
    e = torch.tensor([[[0.2], [0.3], [0.5]]], dtype=torch.float64)  # (R=1, B=3, G=1)
    sf = torch.tensor([[1.0, 2.0, 0.5]], dtype=torch.float64)
    sm = torch.tensor([[1.0, 1.0, 0.0]], dtype=torch.float64)
    a0 = torch.tensor([10.0], dtype=torch.float64)
    a = O.get_alpha(e, sf, sm, a0)
    p = np.array([0.2, 0.6, 0.25])
    want = (p + 1e-5 / 3) / (p.sum() + 1e-5) * 10.0
    want[2] = 1e-5  # masked sample -> clamp floor
    np.testing.assert_allclose(a[0, 0].numpy(), want, rtol=1e-14)


def test_injected_dirichlet_backward_equals_rsample_backward():
    torch.manual_seed(5)
    conc = (torch.rand((4, 1, 6, 2), dtype=torch.float64) * 8 + 0.05).requires_grad_(True)
    torch.manual_seed(9)
    x = torch.distributions.Dirichlet(conc).rsample()
    w = torch.randn_like(x)
    (x * w).sum().backward()
    want = conc.grad.clone()
    conc.grad = None
    xi = O.dirichlet_rsample(conc, x.detach())
    (xi * w).sum().backward()
    assert torch.equal(xi.detach(), x.detach())
    torch.testing.assert_close(conc.grad, want, rtol=0, atol=0)


def test_clipped_adam_matches_torch_adam_with_clamp_and_decay():
    torch.manual_seed(0)
    p0 = torch.randn(5, dtype=torch.float64)
    target = torch.randn(5, dtype=torch.float64) * 50
    lr, lrd, steps = 0.01, 0.1 ** (1 / 20), 20
    p = p0.clone().requires_grad_(True)
    opt = O.ClippedAdam(lr=lr, lrd=lrd)
    q = p0.clone().requires_grad_(True)
    ref = torch.optim.Adam([q], lr=lr, betas=(0.9, 0.999), eps=1e-8)
    for t in range(steps):
        for v in (p, q):
            v.grad = None
            (((v - target) ** 2).sum() * 3).backward()
        opt.step({"p": p})
        q.grad.clamp_(-10, 10)
        for grp in ref.param_groups:
            grp["lr"] = lr * lrd ** (t + 1)
        ref.step()
    # torch.optim.Adam puts eps outside the bias correction (sqrt(v)/sqrt(bc2) + eps), ClippedAdam inside
    # (sqrt(v) + eps): the two agree to O(eps = 1e-8) relative, not to rounding.
    torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-6, atol=1e-7)


def test_ll_core_gradcheck():
    data = H.make_small_mixture_data(n_variants=2, n_reps=2)
    d = H.cast_data(data, torch.float64)
    G = d.n_guides
    g = torch.Generator().manual_seed(2)
    mu = torch.randn((G, 2), generator=g, dtype=torch.float64).requires_grad_(True)
    sd = (torch.rand((G, 2), generator=g, dtype=torch.float64) + 0.5).requires_grad_(True)
    pi = torch.rand((d.n_reps, 1, G, 2), generator=g, dtype=torch.float64)
    pi = (pi / pi.sum(-1, keepdim=True)).requires_grad_(True)
    with H.default_dtype(torch.float64):
        assert torch.autograd.gradcheck(lambda m, s, p: O.sorting_ll_core(d, m, s, p)[0], (mu, sd, pi), eps=1e-6, atol=1e-5, rtol=1e-4)


def test_alpha0_ols_equals_curve_fit():
    from scipy.optimize import curve_fit

    from crispr_bean_b200.alpha0 import _ols_line

    rng = np.random.default_rng(4)
    x = rng.normal(5, 1, 200)
    y = -1.5 + 0.78 * x + rng.normal(0, 0.3, 200)
    popt, _ = curve_fit(lambda x, b0, b1: b0 + b1 * x, x, y)
    np.testing.assert_allclose(_ols_line(x, y), popt, rtol=1e-6)


def test_var_mini_fixture_shapes():
    data = H.load_var_mini()
    assert (data.n_guides, data.n_reps, data.n_condits, data.n_targets) == (30, 2, 5, 2)
    assert data.X.shape == (2, 5, 30)
    # bins sorted by (upper, lower): bot(0,.2) low(.2,.4) high(.6,.8) bulk(0,1) top(.8,1)
    assert data.upper_bounds.tolist() == [0.2, 0.4, 0.8, 1.0, 1.0]
    assert data.lower_bounds.tolist() == [0.0, 0.2, 0.6, 0.0, 0.8]
    # first guide of the CSV (CONTROL_8_g1): rep5 bot, low, high, bulk, top = 84, 798, 245, 302, 368
    scr = H.var_mini_screen()
    i = list(data.screen.guides.index).index("CONTROL_8_g1")
    assert data.X[0, :, i].tolist() == [84.0, 798.0, 245.0, 302.0, 368.0]
    np.testing.assert_allclose(data.size_factor.mean().item(), 1.0, atol=0.2)


@pytest.mark.parametrize("model", ["Normal", "ControlNormal"])
def test_var_mini_frozen_elbo(model):
    frozen = json.load(open(os.path.join(H.GOLDEN, "var_mini_oracle.json")))[model]
    data = H.load_var_mini()
    res = H.oracle_loss_and_grads(model, data, H.fixed_noise(model, data, seed=7), dtype=torch.float64, use_bcmatch=False)
    assert math.isclose(res["loss"], frozen["loss"], rel_tol=1e-10)
    for k, v in frozen["grad_abs_sum"].items():
        assert math.isclose(float(res["grads"][k].abs().sum()), v, rel_tol=1e-8), k


def test_synthetic_mixture_frozen_elbo():
    frozen = json.load(open(os.path.join(H.GOLDEN, "synthetic_oracle.json")))["MixtureNormal"]
    data = H.make_small_mixture_data()
    res = H.oracle_loss_and_grads("MixtureNormal", data, H.fixed_noise("MixtureNormal", data, seed=11), dtype=torch.float64)
    assert math.isclose(res["loss"], frozen["loss"], rel_tol=1e-9)
    for k, v in frozen["grad_abs_sum"].items():
        assert math.isclose(float(res["grads"][k].abs().sum()), v, rel_tol=1e-7), k


def test_elbo_site_decomposition_mixture():
    """loss == -(sum of model sites - sum of guide sites), with the asymmetric pi masks."""
    data = H.make_small_mixture_data()
    data.repguide_mask[0, :5] = False  # force some masked rows
    noise = H.fixed_noise("MixtureNormal", data, seed=1)
    res = H.oracle_loss_and_grads("MixtureNormal", data, noise, dtype=torch.float64)
    aux = res["aux"]
    rg = data.repguide_mask.unsqueeze(1)
    # guide pi site is NOT masked, model pi site is
    assert aux["lq_pi_guide"].shape == rg.shape
    with H.default_dtype(torch.float64):
        ps = res["params"]
        c = ps.constrained()
        T = data.n_targets
        mu_t = c["mu_loc"] + c["mu_scale"] * noise["eps_mu"]
        sd_t = torch.exp(c["sd_loc"] + c["sd_scale"] * noise["eps_sd"])
        model_lp = (torch.distributions.Laplace(0., 1.).log_prob(mu_t).sum()
                    + torch.distributions.LogNormal(torch.zeros(T, 1), 0.01 * torch.ones(T, 1)).log_prob(sd_t).sum()
                    + (aux["lp_pi_model"] * rg).sum() + (aux["lp_bulk_allele"] * rg).sum()
                    + (aux["ll_guide_counts"] * aux["w_guide_counts"]).sum()
                    + (aux["ll_guide_bcmatch_counts"] * aux["w_guide_bcmatch_counts"]).sum())
        guide_lp = (torch.distributions.Normal(c["mu_loc"], c["mu_scale"]).log_prob(mu_t).sum()
                    + torch.distributions.LogNormal(c["sd_loc"], c["sd_scale"]).log_prob(sd_t).sum()
                    + aux["lq_pi_guide"].sum())
    assert math.isclose(res["loss"], float(-(model_lp - guide_lp)), rel_tol=1e-12)
