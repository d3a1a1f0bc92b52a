"""Tiling screens sharded over ranks, host logic on CPU (2 real processes over gloo).

Guides are split into contiguous blocks, every rank keeps ALL edits (an edit's alleles sit in guides of several shards), the
likelihood part of the per-edit gradients is summed over the ranks every step and every rank then applies the same update --
the protocol of `TilingFusedEngine` on the GPUs (tiling_fused.py, bean_svi_tiling_run phases 1 / 2).  The per-shard engine here
is the CPU oracle (test infrastructure); what is under test is `dist.shard_data` / `__getitem__` of the tiling data class, the
all-reduce of the edit gradients with the per-edit terms counted once, and `dist.run_sharded`'s gather."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crispr_bean_b200.data_class import TilingSortingReporterScreenData
from crispr_bean_b200.dist import run_sharded, shard_data, shard_guides
from crispr_bean_b200.synth import make_tiling_screen
from oracle import bean_oracle as O
from tests import helpers as H

EDIT = ("mu_loc", "mu_scale", "sd_loc", "sd_scale")
STEPS = 3


def _data():
    scr = make_tiling_screen(n_guides=40, max_alleles=6, n_reps=2, seed=5)
    return TilingSortingReporterScreenData(scr, control_can_be_selected=True, allele_df_key="allele_counts")


def test_guide_blocks_and_csr_slices():
    assert shard_guides(10, 3) == [(0, 3), (3, 7), (7, 10)]
    data = _data()
    dense = data.allele_to_edit
    seen = 0
    for rank in range(3):
        sub, off = shard_data(data, rank, 3)
        gb, ge = off["guide_offset"], off["guide_offset"] + off["n_guides"]
        assert sub.n_guides == ge - gb and sub.n_edits == data.n_edits and off["n_variants"] == data.n_edits
        assert torch.equal(sub.allele_to_edit, dense[gb:ge])              # the CSR rows of the block, edits numbered globally
        assert torch.equal(sub.allele_mask, data.allele_mask[gb:ge])
        assert torch.equal(sub.allele_counts_control, data.allele_counts_control[:, :, gb:ge])
        assert torch.equal(sub.pi_a0, data.pi_a0[gb:ge]) and torch.equal(sub.X_masked, data.X_masked[:, :, gb:ge])
        seen += sub.n_guides
    assert seen == data.n_guides


def _edit_terms_elbo(ps, noise, sd_scale=0.01):
    """The per-edit part of the ELBO (Laplace / LogNormal priors minus the guide's Normal / LogNormal densities,
    model.py:579-610, :893-921; closed form as in oracle/tiling_closed_form.py)."""
    u = ps.unconstrained
    mu_loc, ls, sd_loc, lt = (u[k].reshape(-1) for k in EDIT)
    mu_e = mu_loc + ls.exp() * noise["eps_mu"].reshape(-1)
    y = sd_loc + lt.exp() * noise["eps_sd"].reshape(-1)
    e_mu, e_sd = noise["eps_mu"].reshape(-1), noise["eps_sd"].reshape(-1)
    half_log_2pi = 0.91893853320467274178
    return (-np.log(2.0) - mu_e.abs() + ls + 0.5 * e_mu ** 2 + half_log_2pi).sum() + \
        (-np.log(sd_scale) - 0.5 * (y / sd_scale) ** 2 + lt + 0.5 * e_sd ** 2).sum()


class ShardedTilingOracle:
    """CPU stand-in for the sharded TilingFusedEngine: same protocol, autograd oracle instead of the kernels."""

    replicated_params = EDIT

    def __init__(self, sub, tables, guide_offset, num_steps, count_edit_terms):
        self.d, self.tables, self.go = H.cast_data(sub, torch.float64), tables, guide_offset
        self.ps, self.opt = O.ParamStore(), O.ClippedAdam(lr=0.01, lrd=0.1 ** (1 / num_steps))
        self.count_edit_terms, self.loss, self.t = count_edit_terms, [], 0

    def run(self, n):
        G = self.d.n_guides
        for _ in range(n):
            tb = self.tables[self.t]
            noise = {"eps_mu": tb["eps_mu"], "eps_sd": tb["eps_sd"], "pi": tb["pi"][:, :, self.go:self.go + G]}
            with H.default_dtype(torch.float64):
                loss, _ = O.elbo_multi_mixture_normal(self.d, self.ps, noise=noise)
                self.ps.zero_grad()
                loss.backward()
                u = self.ps.unconstrained
                # this shard's likelihood part of the edit gradients = total - per-edit terms (loss = -ELBO)
                terms = _edit_terms_elbo(self.ps, noise)
                g_terms = torch.autograd.grad(-terms, [u[k] for k in EDIT])
                for k, gt in zip(EDIT, g_terms):
                    lik = (u[k].grad - gt).contiguous()
                    if dist.is_initialized() and dist.get_world_size() > 1:
                        dist.all_reduce(lik)                      # the exchange step: 4 E numbers here (2 E on the GPU path)
                    u[k].grad = lik + gt                          # every rank adds the per-edit terms once
                self.opt.step(u)
            own = float(loss.detach()) + (0.0 if self.count_edit_terms else float(terms.detach()))  # loss = -ELBO: take the terms out
            self.loss.append(own)
            self.t += 1

    def losses(self):
        return torch.tensor(self.loss, dtype=torch.float64)

    def params(self):
        return self.ps.constrained()


def _tables(data):
    return [{k: v.double() for k, v in H.fixed_noise("MultiMixtureNormal", data, seed=50 + t).items()} for t in range(STEPS)]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    data = _data()
    tables = _tables(data)
    make = lambda sub, guide_offset, variant_offset: ShardedTilingOracle(sub, tables, guide_offset, STEPS, rank == 0)
    params, loss = run_sharded(make, data, STEPS, rank, world)
    torch.save({"params": params, "loss": loss}, f"{out_dir}/r{rank}.pt")
    dist.destroy_process_group()


def test_tiling_fit_sharded_over_two_ranks_equals_the_unsharded_fit(tmp_path):
    port = 29750 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    data = _data()
    full = ShardedTilingOracle(data, _tables(data), 0, STEPS, True)
    full.run(STEPS)
    ref = full.params()
    for rank in range(2):
        out = torch.load(f"{tmp_path}/r{rank}.pt")
        torch.testing.assert_close(out["loss"], full.losses(), rtol=1e-10, atol=0)
        for k, v in ref.items():
            got = out["params"][k]
            assert got.shape == v.shape, (k, got.shape, v.shape)
            torch.testing.assert_close(got, v, rtol=1e-9, atol=1e-12)
