"""The fused tiling step (`bean_svi_tiling_run_*`, three launches per SVI step) against the CPU oracle and the site-kernel
autograd engine it replaces (bean/model/model.py:550-751, :878-962).  Parity with the reference's own vectors is in
tests/test_gpu_golden.py (tiling_small, tiling_wide: gradients and 6-step trajectories)."""
import pytest
import torch

from crispr_bean_b200.data_class import TilingSortingReporterScreenData
from crispr_bean_b200.generic import TilingSviEngine
from crispr_bean_b200.synth import make_tiling_screen
from crispr_bean_b200.tiling_fused import TilingFusedEngine, supports
from oracle import bean_oracle as O
from tests import helpers as H
from tests.test_gpu_svi import rel_err

pytestmark = pytest.mark.gpu


def _data(n_guides=120, max_alleles=8, n_reps=3, seed=3):
    scr = make_tiling_screen(n_guides=n_guides, max_alleles=max_alleles, n_reps=n_reps, seed=seed)
    return TilingSortingReporterScreenData(scr, control_can_be_selected=True, allele_df_key="allele_counts")


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-5)])
@pytest.mark.parametrize("shape", [(120, 8, 3), (800, 16, 4)])  # the second is BASELINE config c3
def test_fused_step_equals_oracle(cuda_device, dtype, tol, shape):
    data = _data(*shape)
    assert supports(data)
    data.repguide_mask[1, ::6] = False
    noise = H.fixed_noise("MultiMixtureNormal", data, seed=5)
    eng = TilingFusedEngine(data, cuda_device, dtype=dtype, num_steps=8)
    g = torch.Generator().manual_seed(7)  # away from the initial point
    eng.edit_params.copy_(0.3 * torch.randn((4, eng.E), generator=g, dtype=torch.float64))
    eng.alpha_u.copy_(torch.where(data.allele_mask, 0.5 * torch.randn(eng.alpha_u.shape, generator=g, dtype=torch.float64),
                                  eng.alpha_u.double().cpu()))
    got = eng.gradients(noise)
    with H.default_dtype(torch.float64):
        ps = O.ParamStore()
        d = H.cast_data(data, torch.float64)
        O.elbo_multi_mixture_normal(d, ps, noise=noise)
        for i, k in enumerate(("mu_loc", "mu_scale", "sd_loc", "sd_scale")):
            ps.unconstrained[k].data.copy_(eng.edit_params[i].double().cpu().reshape(ps.unconstrained[k].shape))
        ps.unconstrained["alpha_pi"].data.copy_(eng.alpha_u.double().cpu())
        loss, _ = O.elbo_multi_mixture_normal(d, ps, noise=noise)
        ps.zero_grad()
        loss.backward()
        # how well float64 itself defines alpha_pi's gradient here: the same formula evaluated a second time on the CPU in another
        # order (oracle/tiling_closed_form.py, numpy, the same torch._dirichlet_grad).  With 16 alleles most of a guide's
        # alleles do not exist, their concentrations are ~1e-6, and torch's pathwise derivative forms psi(c) + 1/c from two
        # numbers of size 1e6 for each: two float64 evaluations agree to ~6e-10 only (800 x 16; 1e-12 at 120 x 8).
        from oracle import tiling_closed_form as TC

        _, closed = TC.tiling_step(d, {k: v.detach() for k, v in ps.unconstrained.items()}, noise)
        floor = rel_err(torch.as_tensor(closed["alpha_pi"]).reshape(ps.unconstrained["alpha_pi"].shape), ps.unconstrained["alpha_pi"].grad)
    ref = float(loss.detach())
    errs = {"loss": abs(got["loss"].item() - ref) / abs(ref)}
    for k, v in ps.unconstrained.items():
        errs[k] = rel_err(got[k], v.grad)
    print(dtype, shape, errs, "float64 floor of alpha_pi", floor)
    assert errs["loss"] <= tol
    for k, e in errs.items():
        bound = 10 * tol if dtype == torch.float32 else tol
        if k == "alpha_pi":
            bound = max(bound, 5 * floor)
        assert e <= bound, (k, e, bound)


def test_fused_steps_equal_autograd_engine_steps(cuda_device):
    data = _data(seed=13)
    fused = TilingFusedEngine(data, cuda_device, dtype=torch.float64, num_steps=20)
    auto = TilingSviEngine(data, cuda_device, dtype=torch.float64, num_steps=20)
    for t in range(5):
        noise = H.fixed_noise("MultiMixtureNormal", data, seed=100 + t)
        fused.run(1, noise=noise)
        auto.run(1, noise=noise, use_graph=False)
    torch.testing.assert_close(fused.losses(), auto.losses(), rtol=1e-10, atol=0)
    pa = auto.params()
    for k, v in fused.params().items():
        assert rel_err(v, pa[k].reshape(v.shape)) <= 1e-9, k


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_free_running_steps_are_deterministic_and_improve_the_elbo(cuda_device, dtype):
    data = _data(n_guides=300, seed=17)
    runs = []
    for _ in range(2):
        eng = TilingFusedEngine(data, cuda_device, dtype=dtype, num_steps=300, seed=9)
        eng.run(150)
        eng.run(150)
        runs.append((eng.losses(), eng.params()))
    assert torch.equal(runs[0][0], runs[1][0])
    for k, v in runs[0][1].items():
        assert torch.equal(v, runs[1][1][k]), k
    loss = runs[0][0]
    assert torch.isfinite(loss).all()
    assert loss[-50:].mean() < loss[:50].mean()


def test_recorded_draws_replay_into_the_oracle(cuda_device):
    """One free-running fp64 step: the kernel's own Philox draws, recorded, give the same loss and gradients in the oracle."""
    data = _data(seed=21)
    eng = TilingFusedEngine(data, cuda_device, dtype=torch.float64, num_steps=4, seed=3)
    loss = eng.run(1, noise={"record": True}, apply_update=False)
    noise = {"eps_mu": eng.eps_used[0].cpu(), "eps_sd": eng.eps_used[1].cpu(), "pi": eng.pi_used.cpu().unsqueeze(1)}
    assert torch.allclose(noise["pi"].sum(-1), torch.ones_like(noise["pi"].sum(-1)), atol=1e-12)
    with H.default_dtype(torch.float64):
        ps = O.ParamStore()
        ref, _ = O.elbo_multi_mixture_normal(H.cast_data(data, torch.float64), ps, noise=noise)
        ref.backward()
    assert abs(loss[0].item() - float(ref.detach())) <= 1e-9 * abs(float(ref.detach()))
    assert rel_err(eng.alpha_grad, ps.unconstrained["alpha_pi"].grad) <= 1e-9
    assert rel_err(eng.edit_grad[0], ps.unconstrained["mu_loc"].grad) <= 1e-9


def test_run_inference_routes_filtered_tiling_designs_to_the_fused_engine(cuda_device):
    from crispr_bean_b200 import model as sm
    from crispr_bean_b200.run import make_engine

    data = _data(seed=23)
    eng = make_engine(sm.MultiMixtureNormalModel, sm.MultiMixtureNormalGuide, data, num_steps=10, device=cuda_device)
    assert isinstance(eng, TilingFusedEngine)
    eng.run(10)
    assert torch.isfinite(eng.losses()).all()
    assert set(eng.params()) == {"mu_loc", "mu_scale", "sd_loc", "sd_scale", "alpha_pi"}
