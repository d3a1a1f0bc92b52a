"""pyro.optim.ClippedAdam restated (pyro/optim/clipped_adam.py); PyroOptim keeps ONE optimiser instance per
parameter, so every parameter carries its own `lr` (decayed by `lrd` at each of its own steps)."""
from __future__ import annotations

import math

import torch


class ClippedAdam:
    def __init__(self, optim_args):
        a = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_norm=10.0, lrd=1.0)
        a.update(optim_args)
        self.args = a
        self.state = {}

    def __call__(self, params):
        a = self.args
        for p in params:
            st = self.state.get(p)
            if st is None:
                st = self.state[p] = {"lr": a["lr"], "step": 0, "exp_avg": torch.zeros_like(p.data),
                                      "exp_avg_sq": torch.zeros_like(p.data)}
            st["lr"] *= a["lrd"]
            if p.grad is None:
                continue
            grad = p.grad.data
            grad.clamp_(-a["clip_norm"], a["clip_norm"])
            st["step"] += 1
            if a["weight_decay"] != 0:
                grad = grad.add(p.data, alpha=a["weight_decay"])
            b1, b2 = a["betas"]
            st["exp_avg"].mul_(b1).add_(grad, alpha=1 - b1)
            st["exp_avg_sq"].mul_(b2).addcmul_(grad, grad, value=1 - b2)
            denom = st["exp_avg_sq"].sqrt().add_(a["eps"])
            step_size = st["lr"] * math.sqrt(1 - b2 ** st["step"]) / (1 - b1 ** st["step"])
            p.data.addcdiv_(st["exp_avg"], denom, value=-step_size)
