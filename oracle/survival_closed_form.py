"""Kernel-shaped restatement of the survival MixtureNormal SVI step: -ELBO and every gradient in CLOSED FORM (numpy, float64,
no autograd), laid out the way a fused CUDA step would compute them -- per guide row maths, two library-wide sums for the
Dirichlet over all guides, one segmented reduction per variant.

TEST INFRASTRUCTURE ONLY (like the rest of oracle/).  It exists to pin the maths of the fused survival step planned in
DESIGN.md before any kernel is written: tests/test_survival_closed_form.py checks it against the autograd oracle
(`bean_oracle.elbo_survival_mixture_normal`, itself pinned to the reference's survival_model.py:215-424 / :651-739).

Unconstrained parameters, as pyro / ClippedAdam see them: q0_u = log q0 (G,), mu_loc (T, 1), mu_scale_u = log mu_scale (T, 1),
alpha_pi_u = log alpha_pi (G, 2).  Draws: eps_mu (T, 1), eps_negctrl (G,), q0 (R, G) on the simplex, pi (R, 1, G, 2).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.special import digamma, gammaln

EPS = 1e-5
HALF_LOG_2PI = 0.9189385332046727


def _np(t):
    return t.detach().double().numpy() if torch.is_tensor(t) else np.asarray(t, dtype=np.float64)


def dm_rows(e, sf, smask, a0, x, w):
    """Dirichlet-Multinomial rows of one count layer.  e (R, B, G) expected fractions -> (sum of masked log-probs,
    d / d e (R, B, G)).  get_alpha (model/utils.py:10-25) + pyro DirichletMultinomial.log_prob + poutine.mask."""
    R, B, G = e.shape
    p = e.transpose(0, 2, 1) * sf[:, None, :]                       # (R, G, B)
    S = p.sum(-1, keepdims=True)
    frac = (p + EPS / B) / (S + EPS)
    raw = frac * a0[None, :, None] * smask[:, None, :]
    live = raw >= EPS                                                # clamp(min=eps) passes the gradient there
    a = np.where(live, raw, EPS)
    xs = x.transpose(0, 2, 1)                                        # (R, G, B)
    A, N = a.sum(-1), xs.sum(-1)
    ll = gammaln(A) + gammaln(1 + N) - gammaln(N + A) - (gammaln(1 + xs) + gammaln(a) - gammaln(xs + a)).sum(-1)
    da = (digamma(A) - digamma(N + A))[..., None] + digamma(xs + a) - digamma(a)   # d ll / d a
    g = np.where(live, da * smask[:, None, :], 0.0)                  # through the clamp and the sample mask
    dot = (g * frac).sum(-1, keepdims=True)
    dp = a0[None, :, None] / (S + EPS) * (g - dot)                   # d ll / d p
    de = (dp * sf[:, None, :]) * w[..., None]
    return float((ll * w).sum()), de.transpose(0, 2, 1)


def dirichlet_grad(x, conc, total):
    return torch._dirichlet_grad(torch.as_tensor(x), torch.as_tensor(conc), torch.as_tensor(total)).numpy()


def survival_mixture_step(data, theta, noise, mu_negctrl=(0.0, 0.1), use_bcmatch=True, mask_thres=10, prob_eps=None):
    """-> (loss, {name: d loss / d unconstrained parameter}).  `data`: a VariantSurvivalReporterScreenData (float64)."""
    G, R, T = data.n_guides, data.n_reps, data.n_targets
    q0_u, mu_loc, ls, al_u = (_np(theta[k]) for k in ("q0", "mu_loc", "mu_scale", "alpha_pi"))
    eps_mu, eps_u, xq, pi = _np(noise["eps_mu"]), _np(noise["eps_negctrl"]), _np(noise["q0"]), _np(noise["pi"])[:, 0]   # pi (R, G, 2)
    tlen = _np(data.target_lengths).astype(np.int64)
    seg = np.repeat(np.arange(T), tlen)                               # variant of every guide
    rg = _np(data.repguide_mask) > 0                                  # (R, G)
    tb, tc = _np(data.timepoints), _np(data.control_timepoint)
    prob_eps = np.finfo(np.float64).eps if prob_eps is None else prob_eps
    elbo = 0.0

    # ---- mu_targets: draw, Laplace prior, Normal guide (closed form: bean_latent_sites) --------------------------------
    s = np.exp(ls)
    mu_t = mu_loc + s * eps_mu                                        # (T, 1)
    elbo += float((-np.log(2.0) - np.abs(mu_t) + ls + 0.5 * eps_mu ** 2 + HALF_LOG_2PI).sum())
    d_mu_t = -np.sign(mu_t)                                           # d ELBO / d mu_t so far; the likelihood adds below
    d_ls_direct = np.ones_like(ls)

    # ---- mu_negctrl: parameter-free prior draw -----------------------------------------------------------------------
    m0, s0 = mu_negctrl
    u = m0 + s0 * eps_u                                               # (G,)
    elbo += float((-np.log(s0) - 0.5 * eps_u ** 2 - HALF_LOG_2PI).sum())
    mu = np.stack([u, mu_t[seg, 0] + u], axis=-1)                     # (G, 2) growth rate of (unedited, edited)

    # ---- abundance sites: Dirichlet over ALL guides, same concentration in model (observed) and guide (drawn) -------------
    c = np.exp(q0_u)
    x0 = _np(data.X)[:, 0, :] + 1.0
    obs = x0 / x0.sum(-1, keepdims=True)
    C = c.sum()
    elbo += float(((c - 1.0)[None, :] * (np.log(obs) - np.log(xq))).sum())
    D = dirichlet_grad(xq, np.broadcast_to(c, (R, G)).copy(), np.full((R, G), C))
    d_c = (np.log(obs) - np.log(xq)).sum(0) + (D * (-(c - 1.0)[None, :] / xq + (C - G))).sum(0)
    grads = {"q0": -(d_c * c)}

    # ---- editing-rate sites (closed form: bean_pi_sites): Dirichlet prior (masked), guide Dirichlet (unmasked), Multinomial --
    al = np.exp(al_u)
    asum = al.sum(-1, keepdims=True)
    pa0 = _np(data.pi_a0)[:, None]
    cm = al / asum * pa0
    cg = np.maximum(cm, 1e-5)
    lp = np.log(pi)
    n_in = rg.sum(0)                                                  # replicates of each guide inside the mask
    norm_m = gammaln(cm.sum(-1)) - gammaln(cm).sum(-1)
    norm_g = gammaln(cg.sum(-1)) - gammaln(cg).sum(-1)
    elbo += float((n_in * norm_m).sum() + (((cm - 1.0)[None] * lp).sum(-1) * rg).sum())
    elbo -= float((R * norm_g).sum() + ((cg - 1.0)[None] * lp).sum())
    d_cm = n_in[:, None] * (digamma(cm.sum(-1))[:, None] - digamma(cm)) + (lp * rg[..., None]).sum(0)
    d_cg = -(R * (digamma(cg.sum(-1))[:, None] - digamma(cg)) + lp.sum(0))
    d_pi = (cm - 1.0)[None] / pi * rg[..., None] - (cg - 1.0)[None] / pi           # (R, G, 2)
    d_mu = np.zeros((G, 2))
    counts = _np(data.allele_counts_control)                          # (R, C, G, 2)
    for ci, t in enumerate(tc):
        w = np.exp(mu * t)                                            # (G, 2)
        q = pi * w[None]
        Sq = q.sum(-1, keepdims=True)
        n = q / Sq
        inside = (n >= prob_eps) & (n <= 1 - prob_eps)
        xc = counts[:, ci]
        elbo += float((xc * np.log(np.clip(n, prob_eps, 1 - prob_eps)) * rg[..., None]).sum())
        elbo += float(((gammaln(xc.sum(-1) + 1) - gammaln(xc + 1).sum(-1)) * rg).sum())      # data-only constant
        h = np.where(inside, xc / n, 0.0)
        dq = (h - (h * n).sum(-1, keepdims=True)) / Sq * rg[..., None]
        d_pi += dq * w[None]
        d_mu += (dq * q * t).sum(0)

    # ---- count likelihood: e[r, b, g] = sum_a pi[r, g, a] exp(mu[g, a] t_b)  (bean_ll, survival mode) -------------------------
    P = np.exp(mu[None] * tb[:, None, None])                          # (B, G, 2)
    e = np.einsum("rga,bga->rbg", pi, P)
    layers = [(_np(data.size_factor), _np(data.a0), _np(data.X_masked))]
    if use_bcmatch:
        layers.append((_np(data.size_factor_bcmatch), _np(data.a0_bcmatch), _np(data.X_bcmatch_masked)))
    smask = _np(data.sample_mask)
    de = np.zeros_like(e)
    for sf, a0, x in layers:
        w = (x.transpose(0, 2, 1).sum(-1) > mask_thres) & rg
        ll, de_l = dm_rows(e, sf, smask, a0, x, w)
        elbo += ll
        de += de_l
    d_pi += np.einsum("rbg,bga->rga", de, P)
    d_mu += np.einsum("rbg,rga,bga,b->ga", de, pi, P, tb)

    # ---- pathwise derivative of the pi draws w.r.t. the guide concentration (torch _Dirichlet_backward) -------------------
    Dpi = dirichlet_grad(pi, np.broadcast_to(cg, (R, G, 2)).copy(), np.broadcast_to(cg.sum(-1, keepdims=True), (R, G, 2)).copy())
    d_cg += (Dpi * (d_pi - (pi * d_pi).sum(-1, keepdims=True))).sum(0)

    # ---- concentrations -> log alpha_pi --------------------------------------------------------------------------------
    dC = d_cm + np.where(cm >= 1e-5, d_cg, 0.0)
    d_al = pa0 / asum ** 2 * (dC * asum - (dC * al).sum(-1, keepdims=True))
    grads["alpha_pi"] = -(d_al * al)

    # ---- edited-allele growth rate -> variant (segmented sum over its guides) -> (mu_loc, log mu_scale) ----------------------
    d_mu_t[:, 0] += np.bincount(seg, weights=d_mu[:, 1], minlength=T)
    grads["mu_loc"] = -d_mu_t
    grads["mu_scale"] = -(d_mu_t * s * eps_mu + d_ls_direct)
    return -elbo, grads


def survival_normal_step(data, theta, noise, use_bcmatch=True, mask_thres=10):
    """survival Normal program (`--uniform-edit`; survival_model.py:15-130, guide :629-650) in closed form.

    Here the Dirichlet draw over all guides DOES reach the likelihood (e[r, b, g] = q_0[r, g] exp(mu_g t_b)), so the pathwise
    derivative needs a third library-wide sum per replicate, sum_g x_rg gout_rg -- the one exchange step of the sharded path
    (crispr_bean_b200/collective.py).  Unconstrained parameters: initial_abundance_u = log c (G,), mu_loc, mu_scale_u (T, 1)."""
    G, R, T = data.n_guides, data.n_reps, data.n_targets
    c_u, mu_loc, ls = (_np(theta[k]) for k in ("initial_abundance", "mu_loc", "mu_scale"))
    eps_mu, x = _np(noise["eps_mu"]), _np(noise["q0"])                 # x (R, G) on the simplex
    seg = np.repeat(np.arange(T), _np(data.target_lengths).astype(np.int64))
    rg = _np(data.repguide_mask) > 0
    tb = _np(data.timepoints)
    s = np.exp(ls)
    mu_t = mu_loc + s * eps_mu
    elbo = float((-np.log(2.0) - np.abs(mu_t) + ls + 0.5 * eps_mu ** 2 + HALF_LOG_2PI).sum())
    d_mu_t = -np.sign(mu_t)
    keep = np.ones(G)
    if hasattr(data, "negctrl_guide_idx"):  # survival_model.py:59-60: None zeroes EVERY guide's growth rate
        keep = np.zeros(G) if data.negctrl_guide_idx is None else keep
        if data.negctrl_guide_idx is not None:
            keep[np.asarray(data.negctrl_guide_idx, dtype=np.int64)] = 0.0
    mu = mu_t[seg, 0] * keep                                           # (G,)
    # Dirichlet over all guides: model prior Dir(1/G), guide Dir(c)
    c = np.exp(c_u)
    prior = np.full(G, 1.0 / G)
    lx = np.log(x)
    C = c.sum()
    norm = lambda a: gammaln(a.sum()) - gammaln(a).sum()
    elbo += R * (norm(prior) - norm(c)) + float((((prior - c)[None]) * lx).sum())
    d_c = -(R * (digamma(C) - digamma(c)) + lx.sum(0))
    gout = (prior - c)[None] / x                                       # d ELBO / d x so far
    # likelihood
    P = np.exp(mu[None] * tb[:, None])                                 # (B, G)
    e = x[:, None, :] * P[None]
    layers = [(_np(data.size_factor), _np(data.a0), _np(data.X_masked))]
    if use_bcmatch:
        layers.append((_np(data.size_factor_bcmatch), _np(data.a0_bcmatch), _np(data.X_bcmatch_masked)))
    smask = _np(data.sample_mask)
    de = np.zeros_like(e)
    for sf, a0, xx in layers:
        w = (xx.transpose(0, 2, 1).sum(-1) > mask_thres) & rg
        ll, de_l = dm_rows(e, sf, smask, a0, xx, w)
        elbo += ll
        de += de_l
    gout += (de * P[None]).sum(1)
    d_mu = (de * e * tb[None, :, None]).sum((0, 1)) * keep
    # pathwise derivative: D (gout - sum_h x_h gout_h): the third sum over all guides
    D = dirichlet_grad(x, np.broadcast_to(c, (R, G)).copy(), np.full((R, G), C))
    d_c += (D * (gout - (x * gout).sum(-1, keepdims=True))).sum(0)
    d_mu_t[:, 0] += np.bincount(seg, weights=d_mu, minlength=T)
    return -elbo, {"initial_abundance": -(d_c * c), "mu_loc": -d_mu_t, "mu_scale": -(d_mu_t * s * eps_mu + 1.0)}
