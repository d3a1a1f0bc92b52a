"""`bean_clipped_adam_f32/f64` (one launch for every parameter tensor of a model) against the oracle's restatement of
pyro.optim.ClippedAdam (bean/model/run.py:368-373: lr 0.01, lrd = gamma^(1/num_steps), clip_norm 10, betas .9/.999).
Tolerance: 1e-12 relative in fp64, 2e-6 in fp32 (the same update in another operation order)."""
import ctypes as C

import pytest
import torch

from crispr_bean_b200 import _lib
from oracle import bean_oracle as O

pytestmark = pytest.mark.gpu

SIZES = [(), (1,), (7, 1), (1000,), (30, 231), (300_001,)]


def launch(theta, grads, m, v, step_sizes, t, dtype):
    args = _lib.BeanAdamArgs()
    args.n_tensors = len(theta)
    for i, (p, g, mi, vi) in enumerate(zip(theta, grads, m, v)):
        s = args.tensors[i]
        s.theta, s.grad, s.m, s.v, s.n = p.data_ptr(), g.data_ptr(), mi.data_ptr(), vi.data_ptr(), p.numel()
    args.step_sizes, args.step, args.n_steps = step_sizes.data_ptr(), t.data_ptr(), step_sizes.numel()
    args.beta1, args.beta2, args.eps, args.clip = 0.9, 0.999, 1e-8, 10.0
    name = "bean_clipped_adam_f64" if dtype == torch.float64 else "bean_clipped_adam_f32"
    return getattr(_lib.lib(), name)(args, torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-6)])
def test_clipped_adam_follows_the_reference_update(cuda_device, dtype, tol):
    n_steps, lr, gamma = 7, 0.01, 0.1
    lrd = gamma ** (1 / n_steps)
    gen = torch.Generator().manual_seed(4)
    host = [torch.randn(s, generator=gen, dtype=torch.float64) for s in SIZES]
    ref = {str(i): p.clone().requires_grad_(True) for i, p in enumerate(host)}
    opt = O.ClippedAdam(lr=lr, lrd=lrd)
    theta = [p.to(cuda_device, dtype).contiguous() for p in host]
    m = [torch.zeros_like(p) for p in theta]
    v = [torch.zeros_like(p) for p in theta]
    k = torch.arange(1, n_steps + 1, dtype=torch.float64)
    step_sizes = (lr * lrd ** k * torch.sqrt(1 - 0.999 ** k) / (1 - 0.9 ** k)).to(cuda_device)
    t = torch.zeros(1, dtype=torch.int64, device=cuda_device)
    for step in range(n_steps):
        # gradients of very different magnitude: some elements beyond the +-10 clamp, some exactly zero
        grads = [torch.randn(s, generator=gen, dtype=torch.float64) * (30.0 if step % 2 else 0.3) for s in SIZES]
        grads[3][::5] = 0.0
        for i, g in enumerate(grads):
            ref[str(i)].grad = g.to(dtype).double()  # the kernel sees the gradient rounded to its dtype
        opt.step(ref)
        assert launch(theta, [g.to(cuda_device, dtype).contiguous() for g in grads], m, v, step_sizes, t, dtype) == 0
        t.add_(1)
    for i, p in enumerate(theta):
        want = ref[str(i)].detach()
        err = ((p.double().cpu() - want).abs().max() / want.abs().max().clamp(min=1e-30)).item()
        assert err <= tol, (SIZES[i], err)
        assert torch.isfinite(m[i]).all() and (v[i] >= 0).all()


def test_step_index_is_read_from_the_device_and_clamped(cuda_device):
    theta = [torch.ones(5, device=cuda_device)]
    g = [torch.full((5,), 2.0, device=cuda_device)]
    m, v = [torch.zeros(5, device=cuda_device)], [torch.zeros(5, device=cuda_device)]
    step_sizes = torch.tensor([0.5, 0.0], dtype=torch.float64, device=cuda_device)
    t = torch.tensor([0], dtype=torch.int64, device=cuda_device)
    assert launch(theta, g, m, v, step_sizes, t, torch.float32) == 0
    first = theta[0].clone()
    assert (first < 1).all()
    t.fill_(99)  # beyond the table: clamped to the last entry (step size 0 -> parameters stay)
    assert launch(theta, g, m, v, step_sizes, t, torch.float32) == 0
    assert torch.equal(theta[0], first)


def test_bad_arguments(cuda_device):
    args = _lib.BeanAdamArgs()
    args.n_tensors = 0
    assert _lib.lib().bean_clipped_adam_f32(C.byref(args), None) == -1
    assert b"n_tensors" in _lib.lib().bean_last_error()
    args.n_tensors = 1
    assert _lib.lib().bean_clipped_adam_f32(C.byref(args), None) == -1  # step_sizes / step NULL
