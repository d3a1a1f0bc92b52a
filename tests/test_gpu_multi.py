"""Multi-GPU tests (need >= 2 GPUs; skipped on a single-GPU box): the survival models' exchange step over NCCL.

Guides of a proliferation screen are sharded across 2 ranks in variant blocks (`dist.shard_data`); the Dirichlet over
ALL guides of the initial abundance sites is evaluated with `collective` all-reduces.  The sharded loss (sum over ranks)
and every gradient must equal the single-GPU engine's on the same injected draws.
Run by hand on a 2-GPU box:  gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -q
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _noise(data, seed):
    g = torch.Generator().manual_seed(seed)
    G, R, T = data.n_guides, data.n_reps, data.n_targets
    gam = torch._standard_gamma(torch.full((R, G), 1.3, dtype=torch.float64), generator=g)
    pig = torch._standard_gamma(torch.full((R, 1, G, 2), 1.5, dtype=torch.float64), generator=g)
    return {"eps_mu": torch.randn((T, 1), generator=g, dtype=torch.float64), "q0": gam / gam.sum(-1, keepdim=True),
            "pi": pig / pig.sum(-1, keepdim=True), "eps_negctrl": torch.randn((G,), generator=g, dtype=torch.float64)}


def _data(model):
    from crispr_bean_b200.data_class import VariantSurvivalReporterScreenData, VariantSurvivalScreenData
    from crispr_bean_b200.synth import make_survival_screen

    scr = make_survival_screen(40, "lognormal", n_reps=3, seed=11, n_negctrl_guides=7)
    if model == "Normal":
        return VariantSurvivalScreenData(scr, control_condition="D7", negctrl_guide_idx=list(range(7)))
    return VariantSurvivalReporterScreenData(scr, control_condition="D7")


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from crispr_bean_b200.dist import shard_data
    from crispr_bean_b200.survival import SurvivalSviEngine

    solo = dist.new_group([rank], use_local_synchronization=True)  # a single-rank group = the unsharded engine
    res = {}
    for model in ("Normal", "MixtureNormal"):
        data = _data(model)
        noise = _noise(data, 3)
        sub, off = shard_data(data, rank, world)
        gb, ge = off["guide_offset"], off["guide_offset"] + off["n_guides"]
        vb, ve = off["variant_offset"], off["variant_offset"] + off["n_variants"]
        part = {"eps_mu": noise["eps_mu"][vb:ve], "q0": noise["q0"][:, gb:ge], "pi": noise["pi"][:, :, gb:ge],
                "eps_negctrl": noise["eps_negctrl"][gb:ge]}
        for dtype in (torch.float64, torch.float32):
            eng = SurvivalSviEngine(sub, model, f"cuda:{rank}", dtype=dtype, num_steps=4)  # WORLD group: sharded
            got = eng.gradients(part)
            loss = got["loss"].double().clone()
            dist.all_reduce(loss)
            # 3 sharded steps with fresh (non-injected) draws must run: the gamma normaliser is all-reduced
            eng.run(3)
            assert torch.isfinite(eng.losses()).all()
            full = SurvivalSviEngine(data, model, f"cuda:{rank}", dtype=dtype, num_steps=4, group=solo).gradients(noise)
            res[(model, str(dtype))] = {"loss": loss.cpu(), "full_loss": full["loss"].double().cpu(), "slices": (gb, ge, vb, ve),
                                        "got": {k: v.cpu() for k, v in got.items() if k != "loss"},
                                        "full": {k: v.cpu() for k, v in full.items() if k != "loss"}}
    torch.save(res, f"{out_dir}/r{rank}.pt")
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_survival_models_sharded_over_two_gpus_equal_single_gpu(tmp_path):
    port = 29700 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        res = torch.load(f"{tmp_path}/r{rank}.pt")
        for (model, dtype), r in res.items():
            tol = 1e-10 if "64" in dtype else 2e-5
            assert abs(r["loss"].item() - r["full_loss"].item()) <= tol * abs(r["full_loss"].item()), (model, dtype)
            gb, ge, vb, ve = r["slices"]
            for k, g in r["got"].items():
                f = r["full"][k]
                ref = f[vb:ve] if f.shape[0] == (r["full"]["mu_loc"].shape[0]) and k.startswith("mu_") else f[gb:ge]
                scale = f.abs().mean().item() + 1e-300
                assert (g.double() - ref.double()).abs().max().item() <= tol * 50 * scale, (model, dtype, k)


def _fused_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from crispr_bean_b200.dist import shard_data
    from crispr_bean_b200.survival_fused import SurvivalFusedEngine

    solo = dist.new_group([rank], use_local_synchronization=True)
    data = _data("MixtureNormal")
    sub, off = shard_data(data, rank, world)
    res = {}
    for dtype in (torch.float64, torch.float32):
        # sharded over the two ranks: the library-wide sums of the abundance Dirichlet travel through peer memory inside the kernels
        eng = SurvivalFusedEngine(sub, f"cuda:{rank}", dtype=dtype, num_steps=12, seed=4, guide_offset=off["guide_offset"],
                                  variant_offset=off["variant_offset"])
        eng.run(5)   # two calls: the second starts from sums the host all-reduced after the first
        eng.run(7)
        loss = eng.losses().to(f"cuda:{rank}")
        dist.all_reduce(loss)
        # the same shards with the host all-reducing between the steps (the path taken where peer memory cannot be mapped)
        host = SurvivalFusedEngine(sub, f"cuda:{rank}", dtype=dtype, num_steps=12, seed=4, guide_offset=off["guide_offset"],
                                   variant_offset=off["variant_offset"])
        os.environ["BEAN_NO_PEER_EXCHANGE"] = "1"
        host.run(12)
        del os.environ["BEAN_NO_PEER_EXCHANGE"]
        full = SurvivalFusedEngine(data, f"cuda:{rank}", dtype=dtype, num_steps=12, seed=4, group=solo)  # unsharded, same seed
        full.run(12)
        res[str(dtype)] = {"loss": loss.cpu(), "full_loss": full.losses(), "off": off, "peer_path": eng.peers is not None,
                           "host_path": host.peers is None, "timeouts": eng.peer_timeouts(),
                           "got": {k: v.cpu() for k, v in eng.params().items()}, "full": {k: v.cpu() for k, v in full.params().items()},
                           "host": {k: v.cpu() for k, v in host.params().items()}}
    torch.save(res, f"{out_dir}/f{rank}.pt")
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_survival_run_sharded_over_two_gpus_equals_single_gpu(tmp_path):
    """12 free-running steps (Philox draws keyed by global guide / variant ids): the two shards together reproduce the
    single-GPU run -- same losses, same parameters (the library-wide sums differ only in summation order)."""
    port = 29900 + os.getpid() % 90
    mp.spawn(_fused_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        res = torch.load(f"{tmp_path}/f{rank}.pt")
        for dtype, r in res.items():
            tol = 1e-9 if "64" in dtype else 1e-4
            assert r["peer_path"] and r["host_path"] and r["timeouts"] == 0, (r["peer_path"], r["host_path"], r["timeouts"])
            torch.testing.assert_close(r["loss"], r["full_loss"], rtol=tol, atol=0)
            for k, g in r["got"].items():  # device-side exchange vs NCCL between the steps: the same sums up to their order
                torch.testing.assert_close(g.double(), r["host"][k].double(), rtol=tol * 10, atol=tol * 10 * g.abs().mean().item())
            off = r["off"]
            gb, ge = off["guide_offset"], off["guide_offset"] + off["n_guides"]
            vb, ve = off["variant_offset"], off["variant_offset"] + off["n_variants"]
            for k, g in r["got"].items():
                f = r["full"][k][vb:ve] if k.startswith("mu_") else r["full"][k][gb:ge]
                torch.testing.assert_close(g.double(), f.double(), rtol=tol * 10, atol=tol * 10 * f.abs().mean().item())


def _run_inference_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from crispr_bean_b200 import model as sorting_model
    from crispr_bean_b200 import survival_model
    from crispr_bean_b200.run import run_inference
    from tests import helpers as H

    out = {}
    data = H.make_small_mixture_data(n_variants=40, n_reps=3, seed=5)
    params, hist = run_inference(sorting_model.MixtureNormalModel, sorting_model.MixtureNormalGuide, data, num_steps=20,
                                 device=f"cuda:{rank}", seed=7)
    out["sorting"] = {"loss": torch.tensor(hist["loss"]), "params": hist["params"]}
    sdata = _data("MixtureNormal")
    params, hist = run_inference(survival_model.MixtureNormalModel, survival_model.MixtureNormalGuide, sdata, num_steps=20,
                                 device=f"cuda:{rank}", seed=7)
    out["survival"] = {"loss": torch.tensor(hist["loss"]), "params": hist["params"]}
    torch.save(out, f"{out_dir}/ri{rank}.pt")
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_run_inference_shards_over_the_ranks_of_torch_distributed(tmp_path):
    """`run_inference` under 2 ranks == `run_inference` in one process (sorting: bit for bit in the parameters -- no data-path
    collective; survival: up to the summation order of the two library-wide sums), on every rank."""
    from crispr_bean_b200 import model as sorting_model
    from crispr_bean_b200 import survival_model
    from crispr_bean_b200.run import run_inference
    from tests import helpers as H

    port = 29800 + os.getpid() % 90
    mp.spawn(_run_inference_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    data = H.make_small_mixture_data(n_variants=40, n_reps=3, seed=5)
    _, solo = run_inference(sorting_model.MixtureNormalModel, sorting_model.MixtureNormalGuide, data, num_steps=20, device="cuda:0", seed=7)
    _, ssolo = run_inference(survival_model.MixtureNormalModel, survival_model.MixtureNormalGuide, _data("MixtureNormal"), num_steps=20,
                             device="cuda:0", seed=7)
    for rank in range(2):
        res = torch.load(f"{tmp_path}/ri{rank}.pt")
        torch.testing.assert_close(res["sorting"]["loss"], torch.tensor(solo["loss"]), rtol=1e-6, atol=0)
        for k, v in solo["params"].items():
            assert torch.equal(res["sorting"]["params"][k], v), k
        torch.testing.assert_close(res["survival"]["loss"], torch.tensor(ssolo["loss"]), rtol=1e-4, atol=0)
        for k, v in ssolo["params"].items():
            torch.testing.assert_close(res["survival"]["params"][k], v, rtol=1e-3, atol=1e-3 * v.abs().mean().item())


def _tiling_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from crispr_bean_b200 import model as sm
    from crispr_bean_b200.data_class import TilingSortingReporterScreenData
    from crispr_bean_b200.dist import shard_data
    from crispr_bean_b200.run import run_inference
    from crispr_bean_b200.synth import make_tiling_screen
    from crispr_bean_b200.tiling_fused import TilingFusedEngine

    solo = dist.new_group([rank], use_local_synchronization=True)
    scr = make_tiling_screen(n_guides=90, max_alleles=8, n_reps=3, seed=11)
    data = TilingSortingReporterScreenData(scr, control_can_be_selected=True, allele_df_key="allele_counts")
    sub, off = shard_data(data, rank, world)
    res = {}
    for dtype in (torch.float64, torch.float32):
        eng = TilingFusedEngine(sub, f"cuda:{rank}", dtype=dtype, num_steps=10, seed=4, guide_offset=off["guide_offset"])
        eng.run(10)
        loss = eng.losses().to(f"cuda:{rank}")
        dist.all_reduce(loss)
        full = TilingFusedEngine(data, f"cuda:{rank}", dtype=dtype, num_steps=10, seed=4, group=solo)  # unsharded, same seed
        full.run(10)
        res[str(dtype)] = {"loss": loss.cpu(), "full_loss": full.losses(), "off": off,
                           "got": {k: v.cpu() for k, v in eng.params().items()}, "full": {k: v.cpu() for k, v in full.params().items()}}
    # and through the seam: run_inference shards the tiling design over the ranks and returns the whole screen's result
    params, hist = run_inference(sm.MultiMixtureNormalModel, sm.MultiMixtureNormalGuide, data, num_steps=10, device=f"cuda:{rank}",
                                 dtype=torch.float64, seed=4)
    res["seam"] = {"loss": torch.tensor(hist["loss"], dtype=torch.float64), "params": {k: torch.as_tensor(v).cpu() for k, v in hist["params"].items()}}
    torch.save(res, f"{out_dir}/t{rank}.pt")
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_tiling_run_sharded_over_two_gpus_equals_single_gpu(tmp_path):
    """10 free-running steps of the tiling step with the guides split over two GPUs (all edits on both, their gradient sums
    all-reduced every step; Philox draws keyed by global guide ids): the same losses and parameters as on one GPU."""
    port = 29990 + os.getpid() % 9
    mp.spawn(_tiling_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        res = torch.load(f"{tmp_path}/t{rank}.pt")
        for dtype in ("torch.float64", "torch.float32"):
            r = res[dtype]
            tol = 1e-9 if "64" in dtype else 1e-4
            torch.testing.assert_close(r["loss"], r["full_loss"], rtol=tol, atol=0)
            gb, ge = r["off"]["guide_offset"], r["off"]["guide_offset"] + r["off"]["n_guides"]
            for k, g in r["got"].items():
                f = r["full"][k][gb:ge] if k == "alpha_pi" else r["full"][k]   # per-edit parameters: replicated on every rank
                torch.testing.assert_close(g.double(), f.double(), rtol=tol * 10, atol=tol * 10 * f.abs().mean().item())
        seam, full = res["seam"], res["torch.float64"]
        torch.testing.assert_close(seam["loss"], full["full_loss"], rtol=1e-9, atol=0)
        for k, v in full["full"].items():
            torch.testing.assert_close(seam["params"][k].double().reshape(v.shape), v.double(), rtol=1e-8, atol=1e-8 * v.abs().mean().item())
