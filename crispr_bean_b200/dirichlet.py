"""`Dirichlet(concentration).rsample()` of the editing-rate site `pi` as an autograd node over the C-ABI kernels
`bean_dirichlet_rsample_*` / `bean_dirichlet_rsample_grad_*` (include/bean_b200.h).

Replaces torch._sample_dirichlet + torch._dirichlet_grad in the guides of the tiling / survival programs
(bean/model/model.py:942-950, bean/model/survival_model.py:699-712, :822-833).  The draws come from the same counter-based
generator as the fused sorting step (Philox keyed by the run seed, indexed by GLOBAL guide id, replicate, allele and the
device-side step counter): reproducible, capturable in a CUDA graph, and independent of how guides are sharded.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

_FWD = {torch.float32: "bean_dirichlet_rsample_f32", torch.float64: "bean_dirichlet_rsample_f64"}
_BWD = {torch.float32: "bean_dirichlet_rsample_grad_f32", torch.float64: "bean_dirichlet_rsample_grad_f64"}


class DirichletStream:
    """Where the draws of one Dirichlet site come from: run seed, site id, global guide offset of this shard and the
    DEVICE step counter (int64 tensor of one element) the kernel reads."""

    def __init__(self, seed: int, step: torch.Tensor, guide_offset: int = 0, site: int = 0):
        assert step.dtype == torch.int64 and step.numel() == 1 and step.is_cuda
        self.seed, self.step, self.guide_offset, self.site = int(seed), step, int(guide_offset), int(site)


def _args(conc, x, n_reps, stream: Optional[DirichletStream]):
    G, A = conc.shape
    a = _lib.BeanDirichletArgs()
    a.n_guides, a.n_reps, a.n_alleles = G, n_reps, A
    a.conc, a.x = conc.data_ptr(), x.data_ptr()
    if stream is not None:
        a.seed, a.guide_offset, a.site, a.step = stream.seed, stream.guide_offset, stream.site, stream.step.data_ptr()
    return a


class _DirichletRsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, conc, n_reps, injected, stream):
        if not conc.is_cuda:
            raise _lib.BeanError("bean_dirichlet_rsample needs CUDA tensors: there is no CPU fallback")
        c = conc.detach().contiguous()
        G, A = c.shape
        if injected is not None:  # parity runs: the reference's recorded draw
            x = injected.to(device=c.device, dtype=c.dtype).reshape(n_reps, G, A).contiguous()
        else:
            x = torch.empty((n_reps, G, A), dtype=c.dtype, device=c.device)
            _lib.check(getattr(_lib.lib(), _FWD[c.dtype])(_args(c, x, n_reps, stream), torch.cuda.current_stream(c.device).cuda_stream),
                       _FWD[c.dtype])
        ctx.save_for_backward(x, c)
        ctx.n_reps = n_reps
        return x.clone()

    @staticmethod
    def backward(ctx, grad_x):
        x, c = ctx.saved_tensors
        g = grad_x.to(c.dtype).contiguous()
        d_conc = torch.empty_like(c)
        a = _args(c, x, ctx.n_reps, None)
        a.grad_x, a.d_conc = g.data_ptr(), d_conc.data_ptr()
        _lib.check(getattr(_lib.lib(), _BWD[c.dtype])(a, torch.cuda.current_stream(c.device).cuda_stream), _BWD[c.dtype])
        return d_conc, None, None, None


def dirichlet_rsample(conc: torch.Tensor, n_reps: int, stream: DirichletStream, injected: Optional[torch.Tensor] = None):
    """conc (G, A) -> draws (R, G, A), one independent Dirichlet(conc[g]) per (replicate, guide); differentiable in conc."""
    return _DirichletRsample.apply(conc, n_reps, injected, stream)
