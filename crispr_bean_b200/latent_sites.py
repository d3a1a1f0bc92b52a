"""`mu_targets` / `sd_targets` sites (draws, prior and guide densities) as one autograd node over the C-ABI kernels
`bean_latent_sites_*` / `bean_latent_sites_grad_*` (include/bean_b200.h)."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

_SUF = {torch.float32: "f32", torch.float64: "f64"}


class LatentPrior:
    """Prior hyper-parameters resolved once (python scalars, or per-element device tensors from `--prior-params`).

    mu ~ Laplace(0, 1) unless `prior_params` names mu_loc / mu_scale (then Normal); sd ~ LogNormal(sd_loc, sd_scale)
    (model.py:579-610; run.py:480-542 for the per-variant tensors)."""

    def __init__(self, n: int, prior_params: Optional[dict], sd_scale_default: float, device, dtype):
        pp = prior_params or {}
        self.n, self.mu_normal = int(n), ("mu_loc" in pp or "mu_scale" in pp)
        self.scalars, self.vectors = {}, {}
        for key, default in (("mu_loc", 0.0), ("mu_scale", 1.0), ("sd_loc", 0.0), ("sd_scale", sd_scale_default)):
            v = pp.get(key, default)
            if torch.is_tensor(v) and v.numel() == self.n and self.n > 1:
                self.vectors[key] = v.detach().to(device=device, dtype=dtype).reshape(-1).contiguous()
                self.scalars[key] = 1.0
            else:
                self.scalars[key] = float(v)


class _LatentSites(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu_loc, mu_ls, sd_loc, sd_ls, eps_mu, eps_sd, prior: LatentPrior):
        if not mu_loc.is_cuda:
            raise _lib.BeanError("bean_latent_sites needs CUDA tensors: there is no CPU fallback")
        dtype, dev, n = mu_loc.dtype, mu_loc.device, mu_loc.numel()
        has_sd = sd_loc is not None
        flat = lambda t: t.detach().to(dtype).reshape(-1).contiguous()
        a = _lib.BeanLatentSitesArgs()
        a.n, a.has_sd, a.mu_prior_normal = n, int(has_sd), int(prior.mu_normal)
        keep = [flat(mu_loc), flat(mu_ls), flat(eps_mu)]
        a.mu_loc, a.mu_log_scale, a.eps_mu = (t.data_ptr() for t in keep)
        mu = torch.empty(n, dtype=dtype, device=dev)
        sd = torch.empty(n, dtype=dtype, device=dev) if has_sd else None
        if has_sd:
            keep += [flat(sd_loc), flat(sd_ls), flat(eps_sd)]
            a.sd_loc, a.sd_log_scale, a.eps_sd = (t.data_ptr() for t in keep[3:])
            a.sd = sd.data_ptr()
        a.mu_prior_loc, a.mu_prior_scale = prior.scalars["mu_loc"], prior.scalars["mu_scale"]
        a.sd_prior_loc, a.sd_prior_scale = prior.scalars["sd_loc"], prior.scalars["sd_scale"]
        for key, field in (("mu_loc", "mu_prior_loc_v"), ("mu_scale", "mu_prior_scale_v"), ("sd_loc", "sd_prior_loc_v"), ("sd_scale", "sd_prior_scale_v")):
            if key in prior.vectors:
                assert prior.vectors[key].dtype == dtype
                setattr(a, field, prior.vectors[key].data_ptr())
        lib = _lib.lib()
        partial = torch.empty(lib.bean_latent_sites_num_partials(n), dtype=torch.float64, device=dev)
        dv = torch.empty((4, n), dtype=dtype, device=dev)
        a.mu, a.partial, a.dv = mu.data_ptr(), partial.data_ptr(), dv.data_ptr()
        name = "bean_latent_sites_" + _SUF[dtype]
        _lib.check(getattr(lib, name)(a, torch.cuda.current_stream(dev).cuda_stream), name)
        ctx.save_for_backward(keep[1], keep[2], *(keep[4:6] if has_sd else []), *([sd] if has_sd else []), dv)
        ctx.has_sd, ctx.shape = has_sd, mu_loc.shape
        v = partial.sum().to(dtype)
        if has_sd:
            return mu.reshape(mu_loc.shape), sd.reshape(mu_loc.shape), v
        return mu.reshape(mu_loc.shape), v

    @staticmethod
    def backward(ctx, g_mu, *rest):
        g_sd, g_v = (rest if ctx.has_sd else (None, rest[0]))
        saved = ctx.saved_tensors
        mu_ls, eps_mu, dv = saved[0], saved[1], saved[-1]
        dtype, dev, n = dv.dtype, dv.device, dv.shape[1]
        a = _lib.BeanLatentSitesGradArgs()
        a.n, a.has_sd = n, int(ctx.has_sd)
        a.mu_log_scale, a.eps_mu, a.dv = mu_ls.data_ptr(), eps_mu.data_ptr(), dv.data_ptr()
        keep = []
        if ctx.has_sd:
            sd_ls, eps_sd, sd = saved[2], saved[3], saved[4]
            a.sd_log_scale, a.eps_sd, a.sd = sd_ls.data_ptr(), eps_sd.data_ptr(), sd.data_ptr()
        for g, field in ((g_mu, "g_mu"), (g_sd, "g_sd"), (g_v, "g_v")):
            if g is not None:
                g = g.to(dtype).reshape(-1).contiguous()
                keep.append(g)
                setattr(a, field, g.data_ptr())
        grad = torch.empty((4, n), dtype=dtype, device=dev)
        a.grad = grad.data_ptr()
        name = "bean_latent_sites_grad_" + _SUF[dtype]
        _lib.check(getattr(_lib.lib(), name)(a, torch.cuda.current_stream(dev).cuda_stream), name)
        r = lambda i: grad[i].reshape(ctx.shape)
        return r(0), r(1), (r(2) if ctx.has_sd else None), (r(3) if ctx.has_sd else None), None, None, None


def latent_sites(mu_loc, mu_log_scale, eps_mu, prior: LatentPrior, sd_loc=None, sd_log_scale=None, eps_sd=None):
    """-> (mu, sd, V) with the sd site, (mu, V) without; V = log-prior minus log-guide density of the draws."""
    return _LatentSites.apply(mu_loc, mu_log_scale, sd_loc, sd_log_scale, eps_mu, eps_sd, prior)
