"""Tiling (MultiMixtureNormal) path: allele <- edit CSR/CSC kernels and the full ELBO against the oracle."""
import pytest
import torch

from crispr_bean_b200.data_class import TilingSortingReporterScreenData
from crispr_bean_b200.generic import TilingSviEngine
from crispr_bean_b200.synth import make_tiling_screen
from crispr_bean_b200.tiling import AlleleMap, allele_gather
from oracle import bean_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def rel_err(got, ref):
    got, ref = got.detach().double().cpu().reshape(-1), ref.detach().double().cpu().reshape(-1)
    return ((got - ref).abs() / (ref.abs() + ref.abs().mean() + 1e-300)).max().item()


def tiling_data(n_guides=60, n_reps=3, seed=2):
    scr = make_tiling_screen(n_guides, n_reps=n_reps, seed=seed)
    return scr, TilingSortingReporterScreenData(scr, control_can_be_selected=True, allele_df_key="allele_counts")


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 1e-6)])
def test_allele_gather_scatter_match_dense_matmul_and_norm(cuda_device, dtype, tol):
    """bean_allele_gather/scatter == allele_to_edit @ mu, ||allele_to_edit * sd||_2 and their autograd (model.py:618-625)."""
    _, data = tiling_data()
    dense = data.allele_to_edit.double()
    g = torch.Generator().manual_seed(0)
    mu = torch.randn(data.n_edits, generator=g, dtype=torch.float64).requires_grad_(True)
    sd = (torch.rand(data.n_edits, generator=g, dtype=torch.float64) + 0.3).requires_grad_(True)
    mu_ref = torch.cat([torch.zeros(data.n_guides, 1, dtype=torch.float64), dense @ mu], -1)
    sd_ref = torch.cat([torch.ones(data.n_guides, 1, dtype=torch.float64), torch.linalg.norm(dense * sd[None, None], dim=-1)], -1)
    w1, w2 = torch.randn(mu_ref.shape, generator=g, dtype=torch.float64), torch.randn(mu_ref.shape, generator=g, dtype=torch.float64)
    ((mu_ref * w1).sum() + (sd_ref * w2).sum()).backward()
    amap = AlleleMap(data.allele_ptr.numpy(), data.allele_edit.numpy(), data.n_guides, data.n_max_alleles, data.n_edits, cuda_device)
    mu_d = mu.detach().to(cuda_device, dtype).requires_grad_(True)
    sd_d = sd.detach().to(cuda_device, dtype).requires_grad_(True)
    mu_a, sd_a = allele_gather(mu_d, sd_d, amap)
    ((mu_a * w1.to(cuda_device, dtype)).sum() + (sd_a * w2.to(cuda_device, dtype)).sum()).backward()
    assert rel_err(mu_a, mu_ref) <= tol and rel_err(sd_a, sd_ref) <= tol
    assert rel_err(mu_d.grad, mu.grad) <= tol and rel_err(sd_d.grad, sd.grad) <= tol


@pytest.mark.parametrize("dtype,tol,tol_alpha", [(torch.float64, 1e-9, 1e-9), (torch.float32, 1e-5, 1e-5)])
def test_multi_mixture_normal_elbo_and_gradients(cuda_device, dtype, tol, tol_alpha):
    """Full MultiMixtureNormal -ELBO and gradients vs the oracle with injected noise.  The alpha_pi gradient
    goes through `torch._dirichlet_grad`, evaluated in double also on the fp32 path (generic._DirichletRsample)."""
    _, data = tiling_data()
    data.repguide_mask[0, ::5] = False
    noise = H.fixed_noise("MultiMixtureNormal", data, seed=3)
    eng = TilingSviEngine(data, cuda_device, dtype=dtype, num_steps=10)
    g = torch.Generator().manual_seed(1)
    start = {k: 0.3 * torch.randn(v.shape, generator=g, dtype=torch.float64) for k, v in eng.theta.items() if k != "alpha_pi"}
    start["alpha_pi"] = eng.theta["alpha_pi"].detach().double().cpu() + 0.3 * torch.randn(eng.theta["alpha_pi"].shape, generator=g, dtype=torch.float64) * data.allele_mask
    with torch.no_grad():
        for k, v in start.items():
            eng.theta[k].copy_(v)
    got = eng.gradients(noise)
    with H.default_dtype(torch.float64):
        d = H.cast_data(data, torch.float64)
        ps = O.ParamStore()
        n = {k: v.double() for k, v in noise.items()}
        O.elbo_multi_mixture_normal(d, ps, noise=n)
        for k, v in start.items():
            ps.unconstrained[k].data.copy_(v)
        loss, _ = O.elbo_multi_mixture_normal(d, ps, noise=n)
        ps.zero_grad()
        loss.backward()
    assert abs(got["loss"].item() - float(loss)) / abs(float(loss)) <= tol
    for k in ("mu_loc", "mu_scale", "sd_loc", "sd_scale"):
        assert rel_err(got[k], ps.unconstrained[k].grad) <= tol, k
    assert rel_err(got["alpha_pi"], ps.unconstrained["alpha_pi"].grad) <= tol_alpha


def test_tiling_run_improves_elbo(cuda_device):
    from functools import partial

    from crispr_bean_b200 import model as M
    from crispr_bean_b200.run import run_inference

    _, data = tiling_data(n_guides=120, n_reps=3, seed=5)
    params, hist = run_inference(partial(M.MultiMixtureNormalModel, use_bcmatch=(True,)), M.MultiMixtureNormalGuide, data,
                                 num_steps=150, device=cuda_device)
    loss = torch.tensor(hist["loss"])
    assert torch.isfinite(loss).all() and loss[-20:].mean() < loss[:10].mean()
    assert hist["params"]["mu_loc"].shape == (data.n_edits,) and hist["params"]["alpha_pi"].shape == (data.n_guides, data.n_max_alleles)


@pytest.mark.parametrize("kind", ["tiling", "survival"])
def test_cuda_graph_replay_equals_eager_steps(cuda_device, kind):
    """`run()` captures one SVI step (forward, backward, ClippedAdam, loss log, step counter) into a CUDA graph and replays
    it; with the same injected draws the replayed steps must reproduce the eagerly launched ones."""
    if kind == "tiling":
        _, data = tiling_data(n_guides=50, seed=4)
        make = lambda: TilingSviEngine(data, cuda_device, dtype=torch.float64, num_steps=12)
        noise = H.fixed_noise("MultiMixtureNormal", data, seed=8)
    else:
        from crispr_bean_b200.data_class import VariantSurvivalReporterScreenData
        from crispr_bean_b200.survival import SurvivalSviEngine
        from crispr_bean_b200.synth import make_survival_screen

        data = VariantSurvivalReporterScreenData(make_survival_screen(15, 4, n_reps=3, seed=4, n_negctrl_guides=5), control_condition="D7")
        make = lambda: SurvivalSviEngine(data, "MixtureNormal", cuda_device, dtype=torch.float64, num_steps=12)
        g = torch.Generator().manual_seed(1)
        G, R, T = data.n_guides, data.n_reps, data.n_targets
        gam = torch._standard_gamma(torch.full((R, G), 1.3, dtype=torch.float64), generator=g)
        pig = torch._standard_gamma(torch.full((R, 1, G, 2), 1.5, dtype=torch.float64), generator=g)
        noise = {"eps_mu": torch.randn((T, 1), generator=g, dtype=torch.float64), "q0": gam / gam.sum(-1, keepdim=True),
                 "pi": pig / pig.sum(-1, keepdim=True), "eps_negctrl": torch.randn((G,), generator=g, dtype=torch.float64)}
    eager, graph = make(), make()
    eager.run(8, noise=noise, use_graph=False)
    graph.run(8, noise=noise, use_graph=True)
    assert graph._graph is not None and graph.step == eager.step == 8
    torch.testing.assert_close(graph.losses(), eager.losses(), rtol=1e-12, atol=0)
    for k, v in eager.params().items():
        torch.testing.assert_close(graph.params()[k], v, rtol=1e-11, atol=1e-13)
    # a second batch of steps replays the same graph (no re-capture), and un-injected runs draw fresh noise every step
    g0 = graph._graph
    graph.run(4, noise=noise, use_graph=True)
    assert graph._graph is g0
    free = make()
    free.run(6)
    ls = free.losses()
    assert torch.isfinite(ls).all() and len(set(ls.tolist())) == 6
