"""N>1 host logic on CPU: variant sharding + the one collective (ELBO scalar) + parameter gather, run as
2 real processes over `gloo`.  The per-shard engine is the CPU oracle here (test infrastructure); the
same `dist.run_sharded` drives the CUDA engine on the GPU box."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crispr_bean_b200.dist import run_sharded, shard_data, shard_variants
from oracle import bean_oracle as O
from tests import helpers as H


def test_shard_variants_balanced_and_contiguous():
    for tl, world in (([5] * 10, 3), ([101, 5, 5, 5, 3, 7], 4), ([5, 5], 4), ([3], 2), (list(range(1, 40)), 8)):
        sh = shard_variants(tl, world)
        assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == len(tl)
        for (vb, ve, gb, ge), nxt in zip(sh, sh[1:] + [None]):
            assert ge - gb == sum(tl[vb:ve])
            if nxt is not None:
                assert nxt[0] == ve and nxt[2] == ge
        assert sum(ve - vb > 0 for vb, ve, _, _ in sh) == min(world, len(tl))  # nobody starves needlessly
    eq = shard_variants([5] * 200_000, 8)
    assert all(ge - gb == 125_000 for _, _, gb, ge in eq)


class OracleEngine:
    """CPU stand-in for SviEngine with noise indexed by GLOBAL ids (like the Philox counters)."""

    def __init__(self, sub, noise_tables, guide_offset, variant_offset, num_steps):
        self.d = H.cast_data(sub, torch.float64)
        self.tables, self.go, self.vo = noise_tables, guide_offset, variant_offset
        self.ps, self.opt = O.ParamStore(), O.ClippedAdam(lr=0.01, lrd=0.1 ** (1 / num_steps))
        self.loss, self.t = [], 0

    def run(self, n):
        T, G = self.d.n_targets, self.d.n_guides
        for _ in range(n):
            tb = self.tables[self.t]
            noise = {"eps_mu": tb["eps_mu"][self.vo:self.vo + T], "eps_sd": tb["eps_sd"][self.vo:self.vo + T],
                     "pi": tb["pi"][:, :, self.go:self.go + G]}
            with H.default_dtype(torch.float64):
                loss, _ = O.elbo_mixture_normal(self.d, self.ps, noise=noise)
                self.ps.zero_grad()
                loss.backward()
                self.opt.step(self.ps.unconstrained)
            self.loss.append(float(loss.detach()))
            self.t += 1

    def losses(self):
        return torch.tensor(self.loss, dtype=torch.float64)

    def params(self):
        return self.ps.constrained()


def _tables(data, steps):
    return [H.fixed_noise("MixtureNormal", data, seed=50 + t) for t in range(steps)]


def _worker(rank, world, port, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = H.make_small_mixture_data(n_variants=9, n_reps=2, seed=5)
        tables = _tables(data, steps)
        params, loss = run_sharded(
            lambda sub, guide_offset, variant_offset: OracleEngine(sub, tables, guide_offset, variant_offset, steps),
            data, steps, rank, world)
        if rank == 0:
            torch.save({"params": params, "loss": loss}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_run_equals_unsharded(tmp_path):
    steps, port = 3, 29500 + (os.getpid() % 2000)
    out = str(tmp_path / "sharded.pt")
    mp.spawn(_worker, args=(2, port, steps, out), nprocs=2, join=True)
    got = torch.load(out)
    data = H.make_small_mixture_data(n_variants=9, n_reps=2, seed=5)
    ref = OracleEngine(data, _tables(data, steps), 0, 0, steps)
    ref.run(steps)
    torch.testing.assert_close(got["loss"], ref.losses(), rtol=1e-12, atol=0)
    for k, v in ref.params().items():
        torch.testing.assert_close(got["params"][k], v, rtol=1e-12, atol=1e-14)


def test_shard_data_keeps_global_constants():
    data = H.make_small_mixture_data(n_variants=9, n_reps=2, seed=5)
    sub, off = shard_data(data, 1, 2)
    gb = off["guide_offset"]
    assert torch.equal(sub.a0, data.a0[gb:gb + sub.n_guides])  # a0 / pi_a0 are fitted on the whole screen
    assert torch.equal(sub.pi_a0, data.pi_a0[gb:gb + sub.n_guides])
    assert torch.equal(sub.size_factor, data.size_factor)
    assert int(sub.target_lengths.sum()) == sub.n_guides and sub.n_targets == off["n_variants"]
