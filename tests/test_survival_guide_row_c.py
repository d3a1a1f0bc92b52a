"""The plain-C per-guide restatement of the survival MixtureNormal step (oracle/survival_guide_row.c, built with gcc) against
the numpy closed form (oracle/survival_closed_form.py), guide by guide: the C loop is the body the fused CUDA step will
carry, checked here on the CPU before any kernel exists."""
import ctypes as C
import shutil

import numpy as np
import pytest
import torch

from crispr_bean_b200 import data_class as dc
from crispr_bean_b200.synth import make_survival_screen
from oracle import bean_oracle as O
from oracle.survival_closed_form import HALF_LOG_2PI, survival_mixture_step
from tests.helpers import cast_data, default_dtype
from tests.test_reference_golden import group, load_case
from tests.test_survival_closed_form import random_noise

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None and shutil.which("cc") is None, reason="no C compiler")
DP, UP = C.POINTER(C.c_double), C.POINTER(C.c_ubyte)


def clib():
    from oracle.build_c import build

    lib = C.CDLL(build())
    fn = lib.survival_mixture_guide
    fn.restype = C.c_double
    fn.argtypes = [C.c_int] * 4 + [DP, DP, DP, DP, UP, DP, DP, DP, C.c_double, DP, DP, DP, DP, C.c_double, C.c_double, DP, DP]
    return fn


def ptr(a):
    return a.ctypes.data_as(UP if a.dtype == np.uint8 else DP)


def run_c(data, theta, noise, mu_negctrl, use_bcmatch=True):
    fn = clib()
    n = lambda t: np.ascontiguousarray(t.detach().double().numpy())
    G, R, B, T = data.n_guides, data.n_reps, data.n_condits, data.n_targets
    layers = [(n(data.X_masked), n(data.a0), n(data.size_factor))]
    if use_bcmatch:
        layers.append((n(data.X_bcmatch_masked), n(data.a0_bcmatch), n(data.size_factor_bcmatch)))
    L = len(layers)
    x = np.ascontiguousarray(np.stack([l[0] for l in layers]).transpose(3, 0, 1, 2))        # (G, L, R, B)
    a0 = np.ascontiguousarray(np.stack([l[1] for l in layers]).T)                              # (G, L)
    sf = np.ascontiguousarray(np.stack([l[2] for l in layers]))                                # (L, R, B)
    smask = n(data.sample_mask)
    rg = np.ascontiguousarray(data.repguide_mask.numpy().astype(np.uint8).T)                   # (G, R)
    tb, tc = n(data.timepoints), n(data.control_timepoint)
    counts = np.ascontiguousarray(n(data.allele_counts_control).transpose(2, 0, 1, 3))         # (G, R, C, 2)
    al = np.exp(n(theta["alpha_pi"]))
    s = np.exp(n(theta["mu_scale"]))
    mu_t = n(theta["mu_loc"]) + s * n(noise["eps_mu"])
    seg = np.repeat(np.arange(T), data.target_lengths.numpy())
    u = mu_negctrl[0] + mu_negctrl[1] * n(noise["eps_negctrl"])
    mu = np.ascontiguousarray(np.stack([u, mu_t[seg, 0] + u], axis=-1))
    pi = np.ascontiguousarray(n(noise["pi"])[:, 0].transpose(1, 0, 2))                         # (G, R, 2)
    pa0 = n(data.pi_a0)
    cm = al / al.sum(-1, keepdims=True) * pa0[:, None]
    cg = np.maximum(cm, 1e-5)
    cgb = np.ascontiguousarray(np.broadcast_to(cg[:, None, :], pi.shape))
    dgrad = torch._dirichlet_grad(torch.as_tensor(pi), torch.as_tensor(cgb), torch.as_tensor(cgb.sum(-1, keepdims=True).repeat(2, -1))).numpy()
    dgrad = np.ascontiguousarray(dgrad)
    elbo = np.zeros(G)
    d_al, d_mu = np.zeros((G, 2)), np.zeros(G)
    for g in range(G):
        da, dm = (C.c_double * 2)(), C.c_double()
        elbo[g] = fn(R, B, len(tc), L, ptr(x[g]), ptr(a0[g]), ptr(sf), ptr(smask), ptr(rg[g]), ptr(tb), ptr(tc), ptr(counts[g]),
                     float(pa0[g]), ptr(np.ascontiguousarray(al[g])), ptr(mu[g]), ptr(pi[g]), ptr(dgrad[g]), 10.0,
                     float(np.finfo(np.float64).eps), da, C.byref(dm))
        d_al[g], d_mu[g] = (da[0], da[1]), dm.value
    return elbo, d_al, d_mu, mu_t, seg, s


def check(data, noise, mu_negctrl=(0.0, 0.1), use_bcmatch=True):
    data = cast_data(data, torch.float64)
    with default_dtype(torch.float64):
        ps = O.ParamStore()
        O.elbo_survival_mixture_normal(data, ps, noise=noise, mu_negctrl=mu_negctrl, use_bcmatch=use_bcmatch)
    theta = {k: v.detach().clone() for k, v in ps.unconstrained.items()}
    loss, grads = survival_mixture_step(data, theta, noise, mu_negctrl=mu_negctrl, use_bcmatch=use_bcmatch)
    elbo_g, d_al, d_mu, mu_t, seg, s = run_c(data, theta, noise, mu_negctrl, use_bcmatch)
    # the terms that are not per guide: variant site, negctrl density, abundance sites
    eps_mu, eps_u = noise["eps_mu"].numpy(), noise["eps_negctrl"].numpy()
    ls = theta["mu_scale"].numpy()
    other = (-np.log(2.0) - np.abs(mu_t) + ls + 0.5 * eps_mu ** 2 + HALF_LOG_2PI).sum()
    other += (-np.log(mu_negctrl[1]) - 0.5 * eps_u ** 2 - HALF_LOG_2PI).sum()
    c = np.exp(theta["q0"].numpy())
    x0 = data.X[:, 0, :].numpy() + 1.0
    other += ((c - 1.0)[None] * (np.log(x0 / x0.sum(-1, keepdims=True)) - np.log(noise["q0"].numpy()))).sum()
    assert abs((elbo_g.sum() + other) - (-loss)) <= 1e-11 * abs(loss)
    assert np.abs(-d_al - grads["alpha_pi"]).max() <= 1e-9 * np.abs(grads["alpha_pi"]).max()
    d_mu_t = -np.sign(mu_t[:, 0]) + np.bincount(seg, weights=d_mu, minlength=len(mu_t))
    assert np.abs(-d_mu_t - grads["mu_loc"][:, 0]).max() <= 1e-9 * np.abs(grads["mu_loc"]).max()


@pytest.mark.parametrize("use_bcmatch", [True, False])
def test_c_row_on_synthetic_screen(use_bcmatch):
    data = dc.VariantSurvivalReporterScreenData(make_survival_screen(14, "lognormal", n_reps=3, seed=12, n_negctrl_guides=5, depth=80.0),
                                                control_condition="D7")
    data.repguide_mask[0, ::4] = False
    check(data, random_noise(data, 1), mu_negctrl=(0.02, 0.3), use_bcmatch=use_bcmatch)


@pytest.mark.parametrize("name", ["survival_mixture", "survival_real_var_mixture"])
def test_c_row_on_reference_golden_cases(name):
    z, data = load_case(name)
    check(data, {k: torch.as_tensor(v) for k, v in group(z, "f64/noise/").items() if "/" not in k})
