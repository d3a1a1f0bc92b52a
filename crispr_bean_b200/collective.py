"""The one real exchange step of the path: sums over ALL guides when the guides are sharded across GPUs.

The survival models put a Dirichlet over the whole guide library (per replicate) on the initial guide abundance
(bean/model/survival_model.py:63-67, :306-311, :660-669).  With variants -- and their guides -- sharded over ranks
(SURVEY section 8e), three quantities need the other ranks: the normaliser of the gamma draws (sum_g gamma_g), the total
concentration (sum_g alpha_g, inside lgamma and torch's pathwise derivative) and, in the backward pass, sum_g x_g * grad_g.
Each is one NCCL all-reduce of `n_reps` numbers per step; everything else of the ELBO stays local to the shard.

Pure torch + torch.distributed, device-agnostic (NCCL on the GPUs; the 2-rank gloo tests run the same code on CPU).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def _active(group) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x.  Every rank's loss may depend on y, and the objective is the SUM of the ranks' losses,
    so d objective / d x = sum over ranks of (d loss_rank / d y): the backward is an all-reduce too."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        y = x.detach().clone().contiguous()
        dist.all_reduce(y, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.detach().clone().contiguous()
        dist.all_reduce(g, group=ctx.group)
        return g, None


def global_sum(x: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of `x` over the ranks of `group`, differentiable (identity on a single rank)."""
    return _AllReduceSum.apply(x, group) if _active(group) else x


def is_first_rank(group=None) -> bool:
    return (not _active(group)) or dist.get_rank(group) == 0


class ShardedDirichletRsample(torch.autograd.Function):
    """`Dirichlet(concentration).rsample()` for a Dirichlet whose LAST axis is sharded across ranks.

    forward: local gamma draws normalised by the global sum (or the injected local slice of a draw), clamped like
    torch._sample_dirichlet; backward: torch's `_Dirichlet_backward` with the global total concentration and the
    global sum_g x_g grad_g (`torch._dirichlet_grad` in double, as on the unsharded path)."""

    @staticmethod
    def forward(ctx, concentration, injected, generator, group):
        conc = concentration.detach()
        if injected is not None:
            x = injected.to(device=conc.device, dtype=conc.dtype)
        else:
            gam = torch._standard_gamma(conc.contiguous(), generator)
            tot = gam.sum(-1, keepdim=True)
            if _active(group):
                dist.all_reduce(tot, group=group)
            fi = torch.finfo(conc.dtype)
            x = (gam / tot).clamp(min=fi.tiny, max=1 - fi.eps / 2)
        total = conc.sum(-1, keepdim=True).double()
        if _active(group):
            dist.all_reduce(total, group=group)
        ctx.group = group
        ctx.save_for_backward(x, conc, total)
        return x.clone()

    @staticmethod
    def backward(ctx, grad_output):
        x, conc, total = ctx.saved_tensors
        c64 = conc.double().contiguous()
        grad = torch._dirichlet_grad(x.double().contiguous(), c64, total.expand_as(c64).contiguous()).to(conc.dtype)
        dot = (x * grad_output).sum(-1, keepdim=True)
        if _active(ctx.group):
            dot = dot.contiguous()
            dist.all_reduce(dot, group=ctx.group)
        return grad * (grad_output - dot), None, None, None


def sharded_dirichlet_log_prob(concentration: torch.Tensor, value: torch.Tensor, group=None) -> torch.Tensor:
    """Sum over the batch of `Dirichlet(concentration).log_prob(value)` with the event axis sharded across ranks.

    Returns THIS rank's share: the local part of sum_g [(c_g - 1) log x_g - lgamma(c_g)], plus lgamma(sum_g c_g) on the
    first rank only -- so the shares add up to the unsharded log-prob, and the gradient of lgamma(total) still reaches
    every rank's concentrations through `global_sum`."""
    local = (torch.xlogy(concentration - 1.0, value) - torch.lgamma(concentration)).sum()
    total = global_sum(concentration.sum(-1), group)
    norm = torch.lgamma(total).sum()
    return local + (norm if is_first_rank(group) else norm * 0.0)
