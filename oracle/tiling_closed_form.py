"""Kernel-shaped restatement of the tiling (sorting MultiMixtureNormal) SVI step: -ELBO and every gradient in CLOSED FORM
(numpy, float64, no autograd), from the CSR allele map -- the shape a fused CUDA step would compute it in.

TEST INFRASTRUCTURE ONLY.  Pins the maths of a fused tiling step (DESIGN.md "what comes next") against the autograd oracle
`bean_oracle.elbo_multi_mixture_normal`, itself pinned to the reference's model.py:550-751 / :878-962
(tests/test_tiling_closed_form.py).  Without `--scale-by-acc`.

Unconstrained parameters: mu_loc, mu_scale_u = log mu_scale, sd_loc, sd_scale_u = log sd_scale (E,), alpha_pi_u = log alpha_pi
(G, A).  Draws: eps_mu, eps_sd (E,), pi (R, 1, G, A).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.special import digamma, gammaln, ndtr, ndtri

from .survival_closed_form import EPS, HALF_LOG_2PI, _np, dirichlet_grad, dm_rows


def bin_probs(uq, lq, mu, sd, exists):
    """P[b, g, a] = Phi((t_u - mu) / sd) - Phi((t_l - mu) / sd) and its derivatives w.r.t. (mu, sd); quantile 1 -> cdf 1,
    quantile 0 -> cdf 0, non-existent alleles -> 0 (model/utils.py:34-76)."""
    B = len(uq)
    P, dPm, dPs = (np.zeros((B,) + mu.shape) for _ in range(3))
    phi = lambda z: np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi)
    sd_safe = np.where(exists, sd, 1.0)
    for b in range(B):
        hi = np.ones_like(mu) if uq[b] == 1.0 else ndtr((ndtri(uq[b]) - mu) / sd_safe)
        lo = np.zeros_like(mu) if lq[b] == 0.0 else ndtr((ndtri(lq[b]) - mu) / sd_safe)
        zu = None if uq[b] == 1.0 else (ndtri(uq[b]) - mu) / sd_safe
        zl = None if lq[b] == 0.0 else (ndtri(lq[b]) - mu) / sd_safe
        fu = 0.0 if zu is None else phi(zu)
        fl = 0.0 if zl is None else phi(zl)
        P[b] = np.where(exists, hi - lo, 0.0)
        dPm[b] = np.where(exists, -(fu - fl) / sd_safe, 0.0)
        dPs[b] = np.where(exists, -((0.0 if zu is None else zu * fu) - (0.0 if zl is None else zl * fl)) / sd_safe, 0.0)
    return P, dPm, dPs


def tiling_step(data, theta, noise, alpha_prior=1.0, sd_scale=0.01, epsilon=EPS, use_bcmatch=True, prob_eps=None):
    """-> (loss, {name: d loss / d unconstrained parameter}).  `data`: a TilingSortingReporterScreenData (float64)."""
    G, R, A, E = data.n_guides, data.n_reps, data.n_max_alleles, data.n_edits
    mu_loc, ls, sd_loc, lt, al_u = (_np(theta[k]) for k in ("mu_loc", "mu_scale", "sd_loc", "sd_scale", "alpha_pi"))
    eps_mu, eps_sd, pi = _np(noise["eps_mu"]), _np(noise["eps_sd"]), _np(noise["pi"])[:, 0]          # pi (R, G, A)
    rg = _np(data.repguide_mask) > 0
    exists = _np(data.allele_mask) > 0                                                                 # (G, A)
    prob_eps = np.finfo(np.float64).eps if prob_eps is None else prob_eps
    elbo = 0.0

    # ---- per-edit latent sites (closed form: bean_latent_sites) ---------------------------------------------------------
    s, t = np.exp(ls), np.exp(lt)
    mu_e = mu_loc + s * eps_mu
    y = sd_loc + t * eps_sd
    sd_e = np.exp(y)
    elbo += float((-np.log(2.0) - np.abs(mu_e) + ls + 0.5 * eps_mu ** 2 + HALF_LOG_2PI).sum())
    elbo += float((-np.log(sd_scale) - 0.5 * (y / sd_scale) ** 2 + lt + 0.5 * eps_sd ** 2).sum())
    d_mu_e = -np.sign(mu_e)
    d_y = -y / sd_scale ** 2

    # ---- allele <- edit gather over the CSR map (bean_allele_gather) -----------------------------------------------------
    ptr = _np(data.allele_ptr).astype(np.int64)
    edits = _np(data.allele_edit).astype(np.int64)
    slot_of = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
    mu_slot = np.bincount(slot_of, weights=mu_e[edits], minlength=len(ptr) - 1)
    sd_slot = np.sqrt(np.bincount(slot_of, weights=sd_e[edits] ** 2, minlength=len(ptr) - 1))
    mu_a = np.concatenate([np.zeros((G, 1)), mu_slot.reshape(G, A - 1)], axis=1)
    sd_a = np.concatenate([np.ones((G, 1)), sd_slot.reshape(G, A - 1)], axis=1)

    # ---- editing-rate sites (closed form: bean_pi_sites; guide site under the mask, not clamped) ------------------------------
    al = np.where(exists, np.exp(al_u), epsilon)
    asum = al.sum(-1, keepdims=True)
    pa0 = _np(data.pi_a0)[:, None]
    cg = al / asum * pa0
    cm_raw = (al + epsilon / A) / (asum + epsilon) * pa0
    cm = np.where(cm_raw < epsilon, epsilon, cm_raw)
    lp = np.log(pi)
    n_in = rg.sum(0)
    norm = lambda c: gammaln(c.sum(-1)) - gammaln(c).sum(-1)
    elbo += float((n_in * (norm(cm) - norm(cg))).sum() + ((((cm - cg)[None]) * lp).sum(-1) * rg).sum())
    d_cm = n_in[:, None] * (digamma(cm.sum(-1))[:, None] - digamma(cm)) + (lp * rg[..., None]).sum(0)
    d_cg = -(n_in[:, None] * (digamma(cg.sum(-1))[:, None] - digamma(cg)) + (lp * rg[..., None]).sum(0))
    d_pi = (cm - cg)[None] / pi * rg[..., None]
    counts = _np(data.allele_counts_control)                                                           # (R, C, G, A)
    Sp = pi.sum(-1, keepdims=True)
    n = pi / Sp
    inside = (n >= prob_eps) & (n <= 1 - prob_eps)
    for ci in range(counts.shape[1]):
        xc = counts[:, ci]
        elbo += float((xc * np.log(np.clip(n, prob_eps, 1 - prob_eps)) * rg[..., None]).sum())
        elbo += float(((gammaln(xc.sum(-1) + 1) - gammaln(xc + 1).sum(-1)) * rg).sum())
        h = np.where(inside, xc / n, 0.0)
        d_pi += (h - (h * n).sum(-1, keepdims=True)) / Sp * rg[..., None]

    # ---- count likelihood: e[r, b, g] = sum_a pi[r, g, a] P[b, g, a]  (bean_ll, sorting mode) ---------------------------------
    P, dPm, dPs = bin_probs(_np(data.upper_bounds), _np(data.lower_bounds), mu_a, sd_a, exists)
    e = np.einsum("rga,bga->rbg", pi, P)
    layers = [(_np(data.size_factor), _np(data.a0), _np(data.X_masked))]
    if use_bcmatch:
        layers.append((_np(data.size_factor_bcmatch), _np(data.a0_bcmatch), _np(data.X_bcmatch_masked)))
    smask = _np(data.sample_mask)
    de = np.zeros_like(e)
    for sf, a0, x in layers:
        w = (x.transpose(0, 2, 1).sum(-1) > 10) & rg
        ll, de_l = dm_rows(e, sf, smask, a0, x, w)
        elbo += ll
        de += de_l
    d_pi += np.einsum("rbg,bga->rga", de, P)
    d_mu_a = np.einsum("rbg,rga,bga->ga", de, pi, dPm)
    d_sd_a = np.einsum("rbg,rga,bga->ga", de, pi, dPs)

    # ---- pathwise derivative of the pi draws w.r.t. the guide concentration -------------------------------------------------
    Dpi = dirichlet_grad(pi, np.broadcast_to(cg, (R, G, A)).copy(), np.broadcast_to(cg.sum(-1, keepdims=True), (R, G, A)).copy())
    d_cg += (Dpi * (d_pi - (pi * d_pi).sum(-1, keepdims=True))).sum(0)

    # ---- concentrations -> log alpha_pi (entries of non-existent alleles are constants) -----------------------------------------
    d_al = pa0 / asum ** 2 * (d_cg * asum - (d_cg * al).sum(-1, keepdims=True))
    dm = np.where(cm_raw < epsilon, 0.0, d_cm)
    S1 = asum + epsilon
    d_al += pa0 / S1 * (dm - (dm * (al + epsilon / A)).sum(-1, keepdims=True) / S1)
    grads = {"alpha_pi": -np.where(exists, d_al * al, 0.0)}

    # ---- allele -> edit scatter (bean_allele_scatter), then the edit parameters ----------------------------------------------------
    g_mu_slot, g_sd_slot = d_mu_a[:, 1:].reshape(-1), d_sd_a[:, 1:].reshape(-1)
    d_mu_e += np.bincount(edits, weights=g_mu_slot[slot_of], minlength=E)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(sd_slot[slot_of] > 0, sd_e[edits] / sd_slot[slot_of], 0.0)
    d_sd_e = np.bincount(edits, weights=g_sd_slot[slot_of] * ratio, minlength=E)
    d_y = d_y + d_sd_e * sd_e
    grads["mu_loc"] = -d_mu_e
    grads["mu_scale"] = -(d_mu_e * s * eps_mu + 1.0)
    grads["sd_loc"] = -d_y
    grads["sd_scale"] = -(d_y * t * eps_sd + 1.0)
    return -elbo, grads


def survival_tiling_step(data, theta, noise, alpha_prior=1.0, epsilon=EPS, mu_negctrl=(0.0, 0.1), use_bcmatch=True, prob_eps=None):
    """Tiling proliferation program (survival MultiMixtureNormal; survival_model.py:427-626, guide :759-833) in closed form:
    growth rates per allele from the CSR map (`mu = u + sum of edit rates`), `exp(mu t)` bin function with non-existent
    alleles multiplied by 0, guide concentration clamped at 1e-5 and scored under the mask, Multinomial on pi exp(mu t_c).
    Unconstrained parameters: mu_loc, mu_scale_u (E,), alpha_pi_u (G, A) (+ the guide's unused initial_abundance)."""
    G, R, A, E = data.n_guides, data.n_reps, data.n_max_alleles, data.n_edits
    mu_loc, ls, al_u = (_np(theta[k]) for k in ("mu_loc", "mu_scale", "alpha_pi"))
    eps_mu, eps_u, pi = _np(noise["eps_mu"]), _np(noise["eps_negctrl"]), _np(noise["pi"])[:, 0]
    rg = _np(data.repguide_mask) > 0
    exists = _np(data.allele_mask) > 0
    tb, tc = _np(data.timepoints), _np(data.control_timepoint)
    prob_eps = np.finfo(np.float64).eps if prob_eps is None else prob_eps
    s = np.exp(ls)
    mu_e = mu_loc + s * eps_mu
    elbo = float((-np.log(2.0) - np.abs(mu_e) + ls + 0.5 * eps_mu ** 2 + HALF_LOG_2PI).sum())
    d_mu_e = -np.sign(mu_e)
    m0, s0 = mu_negctrl
    u = m0 + s0 * eps_u
    elbo += float((-np.log(s0) - 0.5 * eps_u ** 2 - HALF_LOG_2PI).sum())
    ptr = _np(data.allele_ptr).astype(np.int64)
    edits = _np(data.allele_edit).astype(np.int64)
    slot_of = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
    mu_slot = np.bincount(slot_of, weights=mu_e[edits], minlength=len(ptr) - 1)
    mu = u[:, None] + np.concatenate([np.zeros((G, 1)), mu_slot.reshape(G, A - 1)], axis=1)            # (G, A)

    al = np.where(exists, np.exp(al_u), epsilon)
    asum = al.sum(-1, keepdims=True)
    pa0 = _np(data.pi_a0)[:, None]
    cg_raw = al / asum * pa0
    cg = np.maximum(cg_raw, 1e-5)
    cm_raw = (al + epsilon / A) / (asum + epsilon) * pa0
    cm = np.where(cm_raw < epsilon, epsilon, cm_raw)
    lp = np.log(pi)
    n_in = rg.sum(0)
    norm = lambda c: gammaln(c.sum(-1)) - gammaln(c).sum(-1)
    elbo += float((n_in * (norm(cm) - norm(cg))).sum() + (((cm - cg)[None] * lp).sum(-1) * rg).sum())
    d_cm = n_in[:, None] * (digamma(cm.sum(-1))[:, None] - digamma(cm)) + (lp * rg[..., None]).sum(0)
    d_cg = -(n_in[:, None] * (digamma(cg.sum(-1))[:, None] - digamma(cg)) + (lp * rg[..., None]).sum(0))
    d_pi = (cm - cg)[None] / pi * rg[..., None]
    d_mu = np.zeros((G, A))
    counts = _np(data.allele_counts_control)
    for ci, t in enumerate(tc):
        w = np.exp(mu * t)
        q = pi * w[None]
        Sq = q.sum(-1, keepdims=True)
        n = q / Sq
        inside = (n >= prob_eps) & (n <= 1 - prob_eps)
        xc = counts[:, ci]
        elbo += float((xc * np.log(np.clip(n, prob_eps, 1 - prob_eps)) * rg[..., None]).sum())
        elbo += float(((gammaln(xc.sum(-1) + 1) - gammaln(xc + 1).sum(-1)) * rg).sum())
        h = np.where(inside, xc / n, 0.0)
        dq = (h - (h * n).sum(-1, keepdims=True)) / Sq * rg[..., None]
        d_pi += dq * w[None]
        d_mu += (dq * q * t).sum(0)

    P = np.exp(mu[None] * tb[:, None, None]) * exists[None]                                            # (B, G, A)
    e = np.einsum("rga,bga->rbg", pi, P)
    layers = [(_np(data.size_factor), _np(data.a0), _np(data.X_masked))]
    if use_bcmatch:
        layers.append((_np(data.size_factor_bcmatch), _np(data.a0_bcmatch), _np(data.X_bcmatch_masked)))
    smask = _np(data.sample_mask)
    de = np.zeros_like(e)
    for sf, a0, x in layers:
        w = (x.transpose(0, 2, 1).sum(-1) > 10) & rg
        ll, de_l = dm_rows(e, sf, smask, a0, x, w)
        elbo += ll
        de += de_l
    d_pi += np.einsum("rbg,bga->rga", de, P)
    d_mu += np.einsum("rbg,rga,bga,b->ga", de, pi, P, tb)

    Dpi = dirichlet_grad(pi, np.broadcast_to(cg, (R, G, A)).copy(), np.broadcast_to(cg.sum(-1, keepdims=True), (R, G, A)).copy())
    d_cg += (Dpi * (d_pi - (pi * d_pi).sum(-1, keepdims=True))).sum(0)
    dg = np.where(cg_raw >= 1e-5, d_cg, 0.0)
    d_al = pa0 / asum ** 2 * (dg * asum - (dg * al).sum(-1, keepdims=True))
    dm = np.where(cm_raw < epsilon, 0.0, d_cm)
    S1 = asum + epsilon
    d_al += pa0 / S1 * (dm - (dm * (al + epsilon / A)).sum(-1, keepdims=True) / S1)
    d_mu_e += np.bincount(edits, weights=d_mu[:, 1:].reshape(-1)[slot_of], minlength=E)
    return -elbo, {"alpha_pi": -np.where(exists, d_al * al, 0.0), "mu_loc": -d_mu_e, "mu_scale": -(d_mu_e * s * eps_mu + 1.0),
                   "initial_abundance": np.zeros(G)}
