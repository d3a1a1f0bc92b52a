"""SVI engine for the proliferation / survival models (bean/model/survival_model.py): Normal, ControlNormal,
MixtureNormal (+ accessibility scaling) and the tiling MultiMixtureNormal (+ accessibility scaling).

Same split as the tiling engine: the count likelihood -- exp(mu t) allele masses, allele mixture, get_alpha and the
Dirichlet-Multinomial rows of both count layers with their backward -- is the C-ABI kernel (`bean_ll_*` in survival
mode); the Dirichlet-over-guides abundance sites, the editing-rate sites, priors and ClippedAdam are torch CUDA ops
(plumbing).  Sites and quirks follow the reference:
  * survival Normal  (survival_model.py:15-130, guide :629-650): `initial_guide_abundance ~ Dirichlet` over ALL
    guides per replicate, e[r, b, g] = exp(mu_g t_b) q_0[r, g]; negative-control guides have mu := 0 (:58-60);
  * ControlNormal    (:133-213, guide :742-757): one shared growth rate;
  * MixtureNormal    (:215-424, guide :651-739): the guide samples `initial_abundance` although the model observes it,
    `q0` is the guide's (G,) parameter, `mu_negctrl` is a model-only latent drawn from its prior every step;
  * MultiMixtureNormal (:427-626, guide :759-833): tiling screens -- allele growth rate = mu_negctrl + sum of its
    edits' rates (CSR gather kernel), non-existent alleles multiplied by 0; the guide's `initial_abundance`
    parameter is declared but unused.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from ._lib import BeanError
from .collective import ShardedDirichletRsample, global_sum, sharded_dirichlet_log_prob
from .device_pack import DeviceScreen
from .dirichlet import dirichlet_rsample
from .generic import EPS, AutogradSviEngine
from .latent_sites import LatentPrior, latent_sites
from .ll_function import count_log_likelihood
from .pi_sites import PiSiteData, pi_sites
from .tiling import AlleleMap, allele_gather


class SurvivalSviEngine(AutogradSviEngine):
    def __init__(self, data, model: str = "MixtureNormal", device="cuda", dtype=torch.float32, use_bcmatch=True,
                 num_steps=2000, initial_lr=0.01, gamma=0.1, seed=101, alpha_prior=1.0, mask_thres=10,
                 prior_params: Optional[dict] = None, mu_negctrl=(0.0, 0.1), scale_by_accessibility: bool = False,
                 fit_noise: bool = False, epsilon: float = EPS, group=None, guide_offset: int = 0):
        """`group`: torch.distributed process group over which the guides are sharded (variant blocks, `dist.shard_data`);
        the default (None) is the WORLD group when torch.distributed is initialised, otherwise a single rank.  The
        Dirichlet over all guides is then evaluated with one all-reduce of n_reps numbers per sum (collective.py)."""
        if model not in ("Normal", "ControlNormal", "MixtureNormal", "MultiMixtureNormal"):
            raise ValueError(f"SurvivalSviEngine does not implement model {model!r}")
        if not torch.cuda.is_available():
            raise BeanError("SurvivalSviEngine needs a CUDA device: there is no CPU fallback")
        if not getattr(data, "is_survival", False):
            raise ValueError("SurvivalSviEngine needs a *SurvivalScreenData object")
        self.model, self.device, self.dtype = model, torch.device(device), dtype
        self.guide_offset = int(guide_offset)  # global index of this shard's first guide: the `pi` draws are keyed by global ids
        kw = dict(device=self.device, dtype=dtype)
        use_bcmatch = bool(use_bcmatch) and getattr(data, "X_bcmatch_masked", None) is not None
        self.screen = DeviceScreen(data, self.device, dtype=dtype, use_bcmatch=use_bcmatch, mask_thres=mask_thres)
        G, R = data.n_guides, data.n_reps
        self.G, self.R, self.group = G, R, group
        self.G_total = int(global_sum(torch.tensor([float(G)], device=self.device), group).item())  # guides of ALL shards
        self.prior_params = prior_params
        z = lambda *s: torch.zeros(s, **kw)
        self.acc = bool(scale_by_accessibility) and model in ("MixtureNormal", "MultiMixtureNormal")
        if model == "ControlNormal":
            self.T = 1
            theta, positive = {"mu_loc": z(), "mu_scale": z()}, {"mu_scale"}
        elif model == "MultiMixtureNormal":
            self.E, self.A, self.epsilon = int(data.n_edits), int(data.n_max_alleles), float(epsilon)
            self.T = self.E
            self.amap = AlleleMap(data.allele_ptr.numpy(), data.allele_edit.numpy(), G, self.A, self.E, self.device)
            self.allele_mask = data.allele_mask.to(self.device)
            self.allele_mask_u8 = self.allele_mask.to(torch.uint8).contiguous()
            a0 = torch.full((G, self.A), float(alpha_prior), **kw)
            a0[~self.allele_mask] = self.epsilon
            theta = {"initial_abundance": torch.full((G,), 1.0 / self.G_total, **kw).log(), "mu_loc": z(self.E), "mu_scale": z(self.E),
                     "alpha_pi": a0.log()}
            positive = {"initial_abundance", "mu_scale", "alpha_pi"}
        else:
            self.T = int(data.n_targets)
            self.target_lengths = data.target_lengths.to(self.device)
            theta, positive = {"mu_loc": z(self.T, 1), "mu_scale": z(self.T, 1)}, {"mu_scale"}
        uniform = torch.full((G,), 1.0 / self.G_total, **kw)
        if model == "Normal":
            theta["initial_abundance"] = uniform.log()
            positive.add("initial_abundance")
            self.prior_abundance = uniform if not (prior_params and "initial_abundance" in prior_params) else \
                torch.as_tensor(prior_params["initial_abundance"]).to(**kw)
            keep = torch.ones((G, 1), **kw)  # survival_model.py:58-60 (None indexes every guide, as in the reference)
            if hasattr(data, "negctrl_guide_idx"):
                if data.negctrl_guide_idx is None:
                    keep.zero_()
                else:
                    keep[torch.as_tensor(data.negctrl_guide_idx).long().to(self.device)] = 0.0
            self.keep = keep
        if model == "MixtureNormal":
            theta["q0"] = uniform.log()
            theta["alpha_pi"] = torch.full((G, 2), float(alpha_prior), **kw).log()
            positive |= {"q0", "alpha_pi"}
        if model in ("MixtureNormal", "MultiMixtureNormal"):
            self.pi_a0 = torch.as_tensor(data.pi_a0).to(self.device)  # own dtype (float64), see generic.TilingSviEngine
            self.allele_counts_control = data.allele_counts_control.to(**kw)  # (R, C, G, 2)
            # own dtype (float64 in the reference's data class): exp(mu * t) and the control-allele Multinomial are then
            # evaluated in double as in the reference, whose probability clamp [eps, 1 - eps] follows that dtype
            self.control_timepoint = data.control_timepoint.to(self.device, torch.promote_types(data.control_timepoint.dtype, self.dtype))
            self.rg_mask = data.repguide_mask.to(self.device).unsqueeze(1)  # (R, 1, G)
            # MixtureNormalGuide scores its `pi` draws unmasked (survival_model.py:699-712), the tiling guide under repguide_mask
            self.pi_data = PiSiteData(self.allele_counts_control, self.rg_mask, self.control_timepoint,
                                      mask_guide_site=(model == "MultiMixtureNormal"))
            self.pi_dtype = torch.promote_types(torch.promote_types(self.pi_a0.dtype, self.dtype), self.control_timepoint.dtype)
            x0 = data.X[:, 0, :].to(**kw) + 1  # survival_model.py:306-311: observed initial abundance
            self.obs_abundance = x0 / global_sum(x0.sum(-1, keepdim=True), group)
            self.mu_negctrl = (float(mu_negctrl[0]), float(mu_negctrl[1]))
        if self.acc:
            self._acc_init(data, theta, positive, fit_noise)
        prior = {"mu_loc": 0.0, "mu_scale": 1.0} if model == "ControlNormal" else prior_params
        self.latent_prior = LatentPrior(theta["mu_loc"].numel(), prior, 1.0, self.device, dtype)
        self._init_optim(theta, positive, num_steps, initial_lr, gamma, seed)

    # ---------------------------------------------------------------------------------------------
    def elbo_loss(self, noise: Optional[Dict[str, torch.Tensor]] = None):
        """-ELBO of one particle (site lists: SURVEY App. A.8)."""
        kw = dict(device=self.device, dtype=self.dtype)
        G, R, T = self.G, self.R, self.T
        P = self.theta
        shape = () if self.model == "ControlNormal" else ((T,) if self.model == "MultiMixtureNormal" else (T, 1))
        # `mu_targets` site: draw, prior (Laplace | Normal; ControlNormal: Normal(0, 1)) and guide density in one kernel
        mu_t, model_lp = latent_sites(P["mu_loc"], P["mu_scale"], self._draw(noise, "eps_mu", shape), self.latent_prior)
        guide_lp = self._c(0.0)
        injected_q = noise["q0"].to(**kw) if (noise is not None and "q0" in noise) else None

        if self.model == "ControlNormal":
            mu_a = mu_t.reshape(1, 1).expand(G, 1)
            ll = count_log_likelihood(self.screen, mu_a, torch.ones_like(mu_a), None, None)
            return -(model_lp + ll - guide_lp)

        if self.model == "MultiMixtureNormal":
            return self._elbo_tiling(mu_t, model_lp, guide_lp, noise)
        mu_g = torch.repeat_interleave(mu_t, self.target_lengths, dim=0, output_size=self.G)  # (G, 1); output_size: no host sync
        if self.model == "Normal":
            conc = P["initial_abundance"].exp().unsqueeze(0).expand(R, -1)
            # (R, G): one Dirichlet over ALL guides per replicate -> the exchange step when guides are sharded
            q_0 = ShardedDirichletRsample.apply(conc, injected_q, self.gen, self.group)
            guide_lp = guide_lp + sharded_dirichlet_log_prob(conc, q_0, self.group)
            model_lp = model_lp + sharded_dirichlet_log_prob(self.prior_abundance.unsqueeze(0).expand(R, -1), q_0, self.group)
            mu_a = mu_g * self.keep
            ll = count_log_likelihood(self.screen, mu_a, torch.ones_like(mu_a), q_0.t().unsqueeze(-1).contiguous(), None)
            return -(model_lp + ll - guide_lp)

        # MixtureNormal
        alpha_pi = P["alpha_pi"].exp()
        conc_q = P["q0"].exp().unsqueeze(0).expand(R, -1)
        ia = ShardedDirichletRsample.apply(conc_q, injected_q, self.gen, self.group)  # guide-only draw (App. B8)
        # model: Dirichlet(conc_q).log_prob(observed abundance); guide: Dirichlet(conc_q).log_prob(ia).  Same concentration on
        # both sides, so the normalisers lgamma(sum c) - sum lgamma(c) (and their digamma gradients, and the all-reduce of
        # sum c when guides are sharded) cancel in model - guide: only sum (c - 1) (log obs - log ia) remains.
        model_lp = model_lp + (torch.xlogy(conc_q - 1.0, self.obs_abundance.expand_as(conc_q)) - torch.xlogy(conc_q - 1.0, ia)).sum()
        m0, s0 = self.mu_negctrl
        e_u = self._draw(noise, "eps_negctrl", (G,))
        u = m0 + s0 * e_u  # model-only latent: fresh prior noise every step; its Normal(m0, s0) density has no parameters
        model_lp = model_lp - 0.5 * e_u.square().sum() - G * (math.log(s0) + 0.5 * math.log(2 * math.pi))
        mu = torch.cat([u.unsqueeze(-1), mu_g + u.unsqueeze(-1)], dim=-1)  # (G, 2)
        pi_a_scaled = alpha_pi / alpha_pi.sum(-1, keepdim=True) * self.pi_a0[:, None]
        conc_g, conc_m = pi_a_scaled.clamp(min=1e-5), pi_a_scaled  # (G, 2)
        injected = noise["pi"].to(self.device) if (noise is not None and "pi" in noise) else None
        pi = dirichlet_rsample(conc_g, R, self.pi_stream, injected).unsqueeze(1)  # (R, 1, G, 2): bean_dirichlet_rsample_*
        # model `pi` + Multinomial(control counts; pi exp(mu t_c)) under repguide_mask, minus the guide's `pi` density
        model_lp = model_lp + pi_sites(conc_g, conc_m, pi, self.pi_data, growth=mu, work_dtype=self.pi_dtype)
        if self.acc:  # survival_model.py:347-351
            pi, m_lp, g_lp = self._acc_apply(pi, noise)
            model_lp, guide_lp = model_lp + m_lp, guide_lp + g_lp
        pi_g = pi[:, 0].permute(1, 0, 2).contiguous()  # (G, R, 2)
        ll = count_log_likelihood(self.screen, mu, torch.ones_like(mu), pi_g, None)
        return -(model_lp + ll - guide_lp)

    def _elbo_tiling(self, mu_e, model_lp, guide_lp, noise):
        """MultiMixtureNormal: everything after the `mu_targets` site (survival_model.py:469-626, guide :790-833)."""
        kw = dict(device=self.device, dtype=self.dtype)
        G, R, A, eps = self.G, self.R, self.A, self.epsilon
        alpha_pi = self.theta["alpha_pi"].exp()
        alpha_pi = torch.where(self.allele_mask, alpha_pi, torch.full_like(alpha_pi, eps))  # in-place overwrite in the reference
        m0, s0 = self.mu_negctrl
        e_u = self._draw(noise, "eps_negctrl", (G,))
        u = m0 + s0 * e_u
        model_lp = model_lp - 0.5 * e_u.square().sum() - G * (math.log(s0) + 0.5 * math.log(2 * math.pi))
        mu_targets, _ = allele_gather(mu_e, torch.ones_like(mu_e), self.amap)  # (G, A): column 0 = 0, column j = sum of edit rates
        mu = u.unsqueeze(-1) + mu_targets
        conc_g = (alpha_pi / alpha_pi.sum(-1, keepdim=True) * self.pi_a0[:, None]).clamp(min=1e-5)
        conc_m = (alpha_pi + eps / A) / (alpha_pi.sum(-1, keepdim=True) + eps) * self.pi_a0[:, None]
        conc_m = torch.where(conc_m < eps, torch.full_like(conc_m, eps), conc_m)
        injected = noise["pi"].to(self.device) if (noise is not None and "pi" in noise) else None
        pi = dirichlet_rsample(conc_g, R, self.pi_stream, injected).unsqueeze(1)  # (R, 1, G, A)
        model_lp = model_lp + pi_sites(conc_g, conc_m, pi, self.pi_data, growth=mu, work_dtype=self.pi_dtype)
        if self.acc:
            pi, m_lp, g_lp = self._acc_apply(pi, noise)
            model_lp, guide_lp = model_lp + m_lp, guide_lp + g_lp
        pi_g = pi[:, 0].permute(1, 0, 2).contiguous()  # (G, R, A)
        ll = count_log_likelihood(self.screen, mu, torch.ones_like(mu), pi_g, self.allele_mask_u8)
        return -(model_lp + ll - guide_lp)

    def params(self):
        out = super().params()
        if self.model == "ControlNormal":
            out = {k: v.reshape(()) for k, v in out.items()}
        return out
