"""`torch.autograd.Function` over the C-ABI count-likelihood kernel (`bean_ll_f32/f64`).

This is the seam BASELINE.json's north_star names: the model's log-likelihood becomes one autograd node
calling the thin C-ABI `.so`.  forward launches the fused forward+local-gradient kernel and stashes
d ll / d (mu, sd, pi); backward only scales them by the incoming scalar gradient (the ELBO is a plain
sum).
"""
from __future__ import annotations

import torch

from . import _lib
from .device_pack import DeviceScreen

_REAL = {torch.float32: "bean_ll_f32", torch.float64: "bean_ll_f64"}


def launch_ll(screen: DeviceScreen, mu_allele, sd_allele, pi=None, allele_mask=None, want_rows=False):
    """Raw launch.  mu/sd `(G, A)`, pi `(G, R, A)` or None (A == 1).

    Returns dict(ll=0-dim tensor, ll_row=(L,R,G)|None, d_mu, d_sd, d_pi)."""
    lib = _lib.lib()
    if not mu_allele.is_cuda:
        raise _lib.BeanError("bean_ll needs CUDA tensors: there is no CPU fallback")
    dtype, dev = screen.dtype, screen.device
    G, R, L = screen.n_guides, screen.n_reps, screen.n_layers
    mu_allele = mu_allele.detach().to(dtype).contiguous()
    sd_allele = sd_allele.detach().to(dtype).contiguous()
    A = mu_allele.shape[1]
    assert mu_allele.shape == sd_allele.shape == (G, A), (mu_allele.shape, sd_allele.shape, G)
    args = _lib.BeanLLArgs()
    args.n_alleles = A
    args.mu_allele, args.sd_allele = mu_allele.data_ptr(), sd_allele.data_ptr()
    d_mu = torch.empty((G, A), dtype=dtype, device=dev)
    d_sd = torch.empty((G, A), dtype=dtype, device=dev)
    d_pi = None
    if pi is not None:
        pi = pi.detach().to(dtype).contiguous()
        assert pi.shape == (G, R, A), (pi.shape, (G, R, A))
        d_pi = torch.empty((G, R, A), dtype=dtype, device=dev)
        args.pi, args.d_pi = pi.data_ptr(), d_pi.data_ptr()
    if allele_mask is not None:
        allele_mask = allele_mask.to(torch.uint8).contiguous()
        args.allele_mask = allele_mask.data_ptr()
    ll_row = torch.empty((L, R, G), dtype=dtype, device=dev) if want_rows else None
    partial = torch.empty((lib.bean_ll_num_partials(G, A),), dtype=torch.float64, device=dev)
    args.ll_row = ll_row.data_ptr() if want_rows else None
    args.ll_partial = partial.data_ptr()
    args.d_mu, args.d_sd = d_mu.data_ptr(), d_sd.data_ptr()
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(getattr(lib, _REAL[dtype])(screen.c, args, stream), _REAL[dtype])
    return {"ll": partial.sum(), "ll_row": ll_row, "d_mu": d_mu, "d_sd": d_sd, "d_pi": d_pi}


class CountLogLikelihood(torch.autograd.Function):
    """sum of masked Dirichlet-Multinomial log-probs of all count layers (model.py:526-547)."""

    @staticmethod
    def forward(ctx, mu_allele, sd_allele, pi, screen, allele_mask):
        out = launch_ll(screen, mu_allele, sd_allele, pi, allele_mask)
        ctx.has_pi = pi is not None
        ctx.save_for_backward(out["d_mu"], out["d_sd"], *([out["d_pi"]] if pi is not None else []))
        ctx.in_dtypes = (mu_allele.dtype, sd_allele.dtype, pi.dtype if pi is not None else None)
        return out["ll"].to(mu_allele.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        d_mu, d_sd = saved[0], saved[1]
        g_pi = (grad_out * saved[2]).to(ctx.in_dtypes[2]) if ctx.has_pi else None
        return (grad_out * d_mu).to(ctx.in_dtypes[0]), (grad_out * d_sd).to(ctx.in_dtypes[1]), g_pi, None, None


def count_log_likelihood(screen: DeviceScreen, mu_allele, sd_allele, pi=None, allele_mask=None):
    return CountLogLikelihood.apply(mu_allele, sd_allele, pi, screen, allele_mask)
