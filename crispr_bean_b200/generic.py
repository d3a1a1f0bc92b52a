"""SVI engine for the tiling model (`MultiMixtureNormal`): torch autograd around the CUDA kernels.

The hot parts -- Normal-CDF bin masses, allele mixture, get_alpha and the Dirichlet-Multinomial rows with
their backward (`bean_ll_*`), and the allele <- edit contraction (`bean_allele_gather/scatter_*`) -- are the
C-ABI kernels; the O(G A) Dirichlet / Multinomial editing-rate sites, the priors and ClippedAdam are
ordinary torch CUDA ops (plumbing).  Mirrors bean/model/model.py:550-751 (model) and :878-962 (guide);
one `step()` = one `svi.step` of bean/model/run.py:376-380 with the loss kept on the device.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.distributions as tdist

from ._lib import BeanError
from .device_pack import DeviceScreen
from .ll_function import count_log_likelihood
from .tiling import AlleleMap, allele_gather

EPS = 1e-5
PI_NOISE_SD = 0.655


class _DirichletRsample(torch.autograd.Function):
    """Dirichlet.rsample(): value drawn with torch's sampler (or supplied from outside for parity runs);
    backward = torch's pathwise derivative (`_Dirichlet_backward`), with `torch._dirichlet_grad` evaluated in
    DOUBLE also on the float path -- the CUDA float kernel loses the saddle-point cancellation (1e-2 errors),
    the reference's CPU kernel computes in double internally."""

    @staticmethod
    def forward(ctx, concentration, injected, generator):
        x = injected.to(concentration.dtype) if injected is not None else torch._sample_dirichlet(concentration.contiguous(), generator)
        ctx.save_for_backward(x, concentration)
        return x.clone()

    @staticmethod
    def backward(ctx, grad_output):
        x, concentration = ctx.saved_tensors
        c64 = concentration.double().contiguous()
        total = c64.sum(-1, True).expand_as(c64).contiguous()
        grad = torch._dirichlet_grad(x.double().contiguous(), c64, total).to(concentration.dtype)
        return grad * (grad_output - (x * grad_output).sum(-1, True)), None, None


def _masked_sum(mask, lp):
    return torch.where(mask, lp, torch.zeros_like(lp)).sum()


class AutogradSviEngine:
    """Shared SVI plumbing of the models whose ELBO is assembled from torch CUDA ops around the C-ABI kernels:
    unconstrained parameters (positive ones as log), ClippedAdam, device-side loss log."""

    def _init_optim(self, theta, positive, num_steps, initial_lr, gamma, seed):
        self.theta, self.positive = theta, set(positive)
        for p in self.theta.values():
            p.requires_grad_(True)
        self.m = {k: torch.zeros_like(v) for k, v in self.theta.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.theta.items()}
        self.lr0, self.lrd, self.num_steps = float(initial_lr), float(gamma) ** (1.0 / max(num_steps, 1)), int(num_steps)
        self.step = 0
        self.loss = torch.zeros(max(num_steps, 1), dtype=torch.float64, device=self.device)
        self.gen = torch.Generator(device=self.device).manual_seed(int(seed))

    def _draw(self, noise, key, shape):
        if noise is not None and key in noise:
            return noise[key].to(device=self.device, dtype=self.dtype).reshape(shape)
        return torch.randn(shape, generator=self.gen, device=self.device, dtype=self.dtype)

    def _adam(self):
        """pyro.optim.ClippedAdam on the unconstrained tensors."""
        t = self.step + 1
        lr = self.lr0 * self.lrd ** t
        step_size = lr * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        with torch.no_grad():
            for k, p in self.theta.items():
                if p.grad is None:
                    continue
                g = p.grad.clamp(-10.0, 10.0)
                self.m[k].mul_(0.9).add_(g, alpha=0.1)
                self.v[k].mul_(0.999).addcmul_(g, g, value=0.001)
                p.addcdiv_(self.m[k], self.v[k].sqrt().add_(1e-8), value=-step_size)
                p.grad = None

    def run(self, n_steps: int, noise=None):
        for _ in range(n_steps):
            loss = self.elbo_loss(noise)
            loss.backward()
            self.loss[self.step] = loss.detach().double()
            self._adam()
            self.step += 1
        return self.loss[self.step - n_steps:self.step]

    def gradients(self, noise=None):
        loss = self.elbo_loss(noise)
        loss.backward()
        out = {"loss": loss.detach().clone()}
        for k, p in self.theta.items():
            out[k] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().clone()
            p.grad = None
        return out

    def params(self):
        """Constrained values under the reference's parameter names."""
        return {k: (v.detach().exp() if k in self.positive else v.detach().clone()) for k, v in self.theta.items()}

    def losses(self):
        return self.loss[: self.step].cpu()


class TilingSviEngine(AutogradSviEngine):
    """MultiMixtureNormal on one GPU.  Parameter names / shapes follow the pyro guide (model.py:893-937)."""

    def __init__(self, data, device="cuda", dtype=torch.float32, use_bcmatch=True, num_steps=2000, initial_lr=0.01,
                 gamma=0.1, seed=101, alpha_prior=1.0, sd_scale=0.01, epsilon=EPS, prior_params: Optional[dict] = None):
        if not torch.cuda.is_available():
            raise BeanError("TilingSviEngine needs a CUDA device: there is no CPU fallback")
        self.device, self.dtype = torch.device(device), dtype
        kw = dict(device=self.device, dtype=dtype)
        self.screen = DeviceScreen(data, self.device, dtype=dtype, use_bcmatch=use_bcmatch, mask_thres=10)
        self.G, self.R, self.A, self.E = data.n_guides, data.n_reps, data.n_max_alleles, data.n_edits
        self.amap = AlleleMap(data.allele_ptr.numpy(), data.allele_edit.numpy(), self.G, self.A, self.E, self.device)
        self.allele_mask = data.allele_mask.to(self.device)
        self.allele_mask_u8 = self.allele_mask.to(torch.uint8).contiguous()
        self.pi_a0 = torch.as_tensor(data.pi_a0).to(**kw)
        self.allele_counts_control = data.allele_counts_control.to(**kw)  # (R, C, G, A)
        self.rg_mask = data.repguide_mask.to(self.device).unsqueeze(1)  # (R, 1, G)
        self.epsilon, self.sd_scale, self.prior_params = epsilon, sd_scale, prior_params
        a0 = torch.full((self.G, self.A), float(alpha_prior), **kw)
        a0[~self.allele_mask] = epsilon
        z = lambda *s: torch.zeros(s, **kw)
        # unconstrained parameters (positive ones as log), initial values of model.py:893-937
        theta = {"mu_loc": z(self.E), "mu_scale": z(self.E), "sd_loc": z(self.E), "sd_scale": z(self.E), "alpha_pi": a0.log()}
        self._init_optim(theta, {"mu_scale", "sd_scale", "alpha_pi"}, num_steps, initial_lr, gamma, seed)

    # ---------------------------------------------------------------------------------------------
    def elbo_loss(self, noise: Optional[Dict[str, torch.Tensor]] = None):
        """-ELBO of one particle (SURVEY App. A.5 site list for MultiMixtureNormal)."""
        kw = dict(device=self.device, dtype=self.dtype)
        E, G, A, R = self.E, self.G, self.A, self.R
        eps = self.epsilon
        P = self.theta
        mu_loc, sd_loc = P["mu_loc"], P["sd_loc"]
        mu_scale, sd_scale_q, alpha_pi = P["mu_scale"].exp(), P["sd_scale"].exp(), P["alpha_pi"].exp()
        alpha_pi = torch.where(self.allele_mask, alpha_pi, torch.full_like(alpha_pi, eps))  # model.py:645 / :937

        def draw(key, shape):
            if noise is not None and key in noise:
                return noise[key].to(**kw).reshape(shape)
            return torch.randn(shape, generator=self.gen, **kw)

        mu_e = mu_loc + mu_scale * draw("eps_mu", (E,))
        sd_e = torch.exp(sd_loc + sd_scale_q * draw("eps_sd", (E,)))
        guide_lp = tdist.Normal(mu_loc, mu_scale).log_prob(mu_e).sum() + tdist.LogNormal(sd_loc, sd_scale_q).log_prob(sd_e).sum()
        pp = self.prior_params or {}
        mu_prior = (tdist.Normal(torch.as_tensor(pp.get("mu_loc", 0.0), **kw), torch.as_tensor(pp.get("mu_scale", 1.0), **kw))
                    if ("mu_loc" in pp or "mu_scale" in pp) else tdist.Laplace(torch.zeros((), **kw), torch.ones((), **kw)))
        sd_prior = tdist.LogNormal(torch.as_tensor(pp.get("sd_loc", torch.zeros(E, **kw)), **kw),
                                   torch.as_tensor(pp.get("sd_scale", torch.full((E,), self.sd_scale, **kw)), **kw))
        model_lp = mu_prior.log_prob(mu_e).sum() + sd_prior.log_prob(sd_e).sum()

        # allele <- edit contraction (CUDA CSR gather / CSC scatter), WT column (0, 1)
        mu_a, sd_a = allele_gather(mu_e, sd_e, self.amap)

        # editing-rate sites (model.py:632-670 model, :938-950 guide: masked, not clamped)
        conc_g = (alpha_pi / alpha_pi.sum(-1, keepdim=True) * self.pi_a0[:, None]).unsqueeze(0).unsqueeze(0).expand(R, 1, -1, -1)
        conc_m = (alpha_pi + eps / A) / (alpha_pi.sum(-1, keepdim=True) + eps) * self.pi_a0[:, None]
        conc_m = torch.where(conc_m < eps, torch.full_like(conc_m, eps), conc_m).unsqueeze(0).unsqueeze(0).expand(R, 1, -1, -1)
        injected = noise["pi"].to(**kw) if (noise is not None and "pi" in noise) else None
        pi = _DirichletRsample.apply(conc_g, injected, self.gen)
        guide_lp = guide_lp + _masked_sum(self.rg_mask, tdist.Dirichlet(conc_g, validate_args=False).log_prob(pi))
        model_lp = model_lp + _masked_sum(self.rg_mask, tdist.Dirichlet(conc_m, validate_args=False).log_prob(pi))
        lp_mult = tdist.Multinomial(probs=pi, validate_args=False).log_prob(self.allele_counts_control)
        model_lp = model_lp + _masked_sum(self.rg_mask.expand(lp_mult.shape), lp_mult)

        # count likelihood of both layers (CUDA): bin masses, mixture, get_alpha, Dirichlet-Multinomial
        pi_g = pi[:, 0].permute(1, 0, 2).contiguous()  # (G, R, A)
        model_lp = model_lp + count_log_likelihood(self.screen, mu_a, sd_a, pi_g, self.allele_mask_u8)
        return -(model_lp - guide_lp)
