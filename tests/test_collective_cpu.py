"""2-rank gloo test of the survival models' exchange step (crispr_bean_b200.collective): a Dirichlet over guides sharded
across ranks gives the unsharded log-prob, draw normalisation and pathwise gradients."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _reference(conc, x, w):
    """Unsharded: loss = Dirichlet(conc).log_prob(x).sum() + (w * x).sum() with x a (replayed) reparameterised draw."""
    from oracle.bean_oracle import InjectedDirichlet

    c = conc.clone().requires_grad_(True)
    xs = InjectedDirichlet.apply(c, x)
    loss = torch.distributions.Dirichlet(c, validate_args=False).log_prob(xs).sum() + (w * xs).sum()
    loss.backward()
    return loss.detach(), c.grad


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from crispr_bean_b200.collective import ShardedDirichletRsample, global_sum, sharded_dirichlet_log_prob

    g = torch.Generator().manual_seed(5)
    R, G = 3, 11
    conc = torch.rand((R, G), generator=g, dtype=torch.float64) * 2 + 0.05
    gam = torch._standard_gamma(conc, generator=g)
    x = gam / gam.sum(-1, keepdim=True)
    w = torch.randn((R, G), generator=g, dtype=torch.float64)
    cut = [0, 4, G]  # ragged shards
    sl = slice(cut[rank], cut[rank + 1])
    c_loc = conc[:, sl].clone().requires_grad_(True)
    xs = ShardedDirichletRsample.apply(c_loc, x[:, sl], None, None)
    loss = sharded_dirichlet_log_prob(c_loc, xs) + (w[:, sl] * xs).sum()
    loss.backward()
    total = loss.detach().clone()
    dist.all_reduce(total)
    # sampling path: local gammas normalised by the global sum -> the shards of one draw sum to 1 per replicate
    draw = ShardedDirichletRsample.apply(c_loc.detach(), None, torch.Generator().manual_seed(100 + rank), None)
    s = global_sum(draw.sum(-1))
    torch.save({"loss": total, "grad": c_loc.grad, "draw_sum": s, "slice": (cut[rank], cut[rank + 1])}, f"{out_dir}/r{rank}.pt")
    dist.destroy_process_group()


def test_sharded_dirichlet_equals_unsharded(tmp_path):
    port = 29650 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g = torch.Generator().manual_seed(5)
    R, G = 3, 11
    conc = torch.rand((R, G), generator=g, dtype=torch.float64) * 2 + 0.05
    gam = torch._standard_gamma(conc, generator=g)
    x = gam / gam.sum(-1, keepdim=True)
    w = torch.randn((R, G), generator=g, dtype=torch.float64)
    ref_loss, ref_grad = _reference(conc, x, w)
    outs = [torch.load(f"{tmp_path}/r{r}.pt") for r in range(2)]
    for o in outs:
        assert abs(o["loss"].item() - ref_loss.item()) <= 1e-12 * abs(ref_loss.item())
        a, b = o["slice"]
        assert torch.allclose(o["grad"], ref_grad[:, a:b], rtol=1e-10, atol=1e-12)
        assert torch.allclose(o["draw_sum"], torch.ones(R, dtype=torch.float64), atol=1e-12)


def test_single_process_is_the_plain_dirichlet():
    from crispr_bean_b200.collective import ShardedDirichletRsample, sharded_dirichlet_log_prob

    g = torch.Generator().manual_seed(2)
    conc = torch.rand((2, 6), generator=g, dtype=torch.float64) + 0.1
    x = torch.distributions.Dirichlet(conc).sample()
    w = torch.randn((2, 6), generator=g, dtype=torch.float64)
    ref_loss, ref_grad = _reference(conc, x, w)
    c = conc.clone().requires_grad_(True)
    xs = ShardedDirichletRsample.apply(c, x, None, None)
    loss = sharded_dirichlet_log_prob(c, xs) + (w * xs).sum()
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-12 * abs(ref_loss.item())
    assert torch.allclose(c.grad, ref_grad, rtol=1e-10)
