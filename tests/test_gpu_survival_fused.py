"""The fused survival MixtureNormal step (`bean_svi_survival_run_*`, three launches per SVI step) against the autograd
engine it replaces and the CPU oracle (bean/model/survival_model.py:215-424, :651-739).  Parity with the reference's own
vectors is in tests/test_gpu_golden.py (survival_mixture*, survival_real_var_mixture: gradients and 6-step trajectories)."""
import numpy as np
import pytest
import torch
from scipy import stats

from crispr_bean_b200.data_class import VariantSurvivalReporterScreenData
from crispr_bean_b200.survival import SurvivalSviEngine
from crispr_bean_b200.survival_fused import SurvivalFusedEngine
from crispr_bean_b200.synth import make_survival_screen
from oracle import bean_oracle as O
from tests import helpers as H
from tests.test_gpu_svi import rel_err

pytestmark = pytest.mark.gpu


def _data(n_variants=60, seed=11, n_reps=3):
    scr = make_survival_screen(n_variants, "lognormal", n_reps=n_reps, seed=seed, n_negctrl_guides=7)
    return VariantSurvivalReporterScreenData(scr, control_condition="D7")


def _noise(data, seed):
    g = torch.Generator().manual_seed(seed)
    G, R, T = data.n_guides, data.n_reps, data.n_targets
    gam = torch._standard_gamma(torch.full((R, G), 1.3, dtype=torch.float64), generator=g)
    pig = torch._standard_gamma(torch.full((R, 1, G, 2), 1.5, dtype=torch.float64), generator=g)
    return {"eps_mu": torch.randn((T, 1), generator=g, dtype=torch.float64), "q0": gam / gam.sum(-1, keepdim=True),
            "pi": pig / pig.sum(-1, keepdim=True), "eps_negctrl": torch.randn((G,), generator=g, dtype=torch.float64)}


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-5)])
def test_fused_step_equals_oracle(cuda_device, dtype, tol):
    data = _data()
    data.repguide_mask[1, ::5] = False  # rows outside the mask
    noise = _noise(data, 3)
    eng = SurvivalFusedEngine(data, cuda_device, dtype=dtype, num_steps=8)
    g = torch.Generator().manual_seed(5)  # away from the initial point
    eng.var_params[:2].copy_(0.3 * torch.randn((2, eng.T), generator=g, dtype=torch.float64))
    eng.alpha_u.copy_(0.5 * torch.randn(eng.alpha_u.shape, generator=g, dtype=torch.float64))
    eng.q0_u.add_((0.3 * torch.randn(eng.G, generator=g, dtype=torch.float64)).to(eng.q0_u))
    got = eng.gradients(noise)
    with H.default_dtype(torch.float64):
        ps = O.ParamStore()
        d = H.cast_data(data, torch.float64)
        O.elbo_survival_mixture_normal(d, ps, noise=noise)  # creates the parameters
        ps.unconstrained["mu_loc"].data.copy_(eng.var_params[0].double().cpu().reshape(-1, 1))
        ps.unconstrained["mu_scale"].data.copy_(eng.var_params[1].double().cpu().reshape(-1, 1))
        ps.unconstrained["alpha_pi"].data.copy_(eng.alpha_u.double().cpu())
        ps.unconstrained["q0"].data.copy_(eng.q0_u.double().cpu())
        loss, _ = O.elbo_survival_mixture_normal(d, ps, noise=noise)
        ps.zero_grad()
        loss.backward()
    ref = float(loss.detach())
    errs = {"loss": abs(got["loss"].item() - ref) / abs(ref)}
    for k, v in ps.unconstrained.items():
        errs[k] = rel_err(got[k], v.grad)
    print(dtype, errs)
    assert errs["loss"] <= tol
    for k, e in errs.items():
        assert e <= (10 * tol if dtype == torch.float32 and k in ("q0", "alpha_pi") else tol), (k, e)


def test_fused_steps_equal_autograd_engine_steps(cuda_device):
    """Same injected draws, 5 ClippedAdam steps: parameters and losses of the fused engine == the autograd engine's (fp64)."""
    data = _data(seed=13)
    fused = SurvivalFusedEngine(data, cuda_device, dtype=torch.float64, num_steps=20)
    auto = SurvivalSviEngine(data, "MixtureNormal", cuda_device, dtype=torch.float64, num_steps=20)
    for t in range(5):
        noise = _noise(data, 100 + t)
        fused.run(1, noise=noise)
        auto.run(1, noise=noise, use_graph=False)
    torch.testing.assert_close(fused.losses(), auto.losses(), rtol=1e-10, atol=0)
    pa = auto.params()
    for k, v in fused.params().items():
        assert rel_err(v, pa[k].reshape(v.shape)) <= 1e-9, k


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_free_running_steps_are_deterministic_and_improve_the_elbo(cuda_device, dtype):
    data = _data(n_variants=200, seed=17)
    runs = []
    for _ in range(2):
        eng = SurvivalFusedEngine(data, cuda_device, dtype=dtype, num_steps=300, seed=9)
        eng.run(150)
        eng.run(150)  # a run continues exactly where the previous call stopped
        runs.append((eng.losses(), eng.params()))
    assert torch.equal(runs[0][0], runs[1][0])
    for k, v in runs[0][1].items():
        assert torch.equal(v, runs[1][1][k]), k
    loss = runs[0][0]
    assert torch.isfinite(loss).all()
    assert loss[-50:].mean() < loss[:50].mean()
    one = SurvivalFusedEngine(data, cuda_device, dtype=dtype, num_steps=300, seed=9)
    one.run(300)  # chunking does not change the noise stream
    assert torch.equal(one.losses(), loss)


def test_abundance_draw_is_a_dirichlet_over_all_guides(cuda_device):
    """x[r][g] = gamma[r][g] / sum_g gamma[r][.] with gamma ~ Gamma(q0[g]): every component is Beta(q0_g, sum q0 - q0_g)."""
    data = _data(n_variants=400, seed=19)
    eng = SurvivalFusedEngine(data, cuda_device, dtype=torch.float64, num_steps=4, seed=2)
    g = torch.Generator().manual_seed(1)
    conc = torch.exp(torch.empty(eng.G, dtype=torch.float64).uniform_(np.log(0.3), np.log(30.0), generator=g))
    eng.q0_u.copy_(conc.log())
    eng.run(1)  # primes gamma[0] / sums[0] from the q0 set above, then steps
    gam, sums = eng.gamma[0].cpu(), eng.sums[0].cpu()
    torch.testing.assert_close(sums[:-1], gam.sum(-1), rtol=1e-12, atol=0)
    torch.testing.assert_close(sums[-1], conc.sum(), rtol=1e-12, atol=0)
    x = (gam / sums[:-1, None]).numpy()
    u = stats.beta.cdf(x, conc.numpy()[None], (conc.sum() - conc).numpy()[None]).reshape(-1)
    assert stats.kstest(u, "uniform").statistic < 1.5 * 1.63 / np.sqrt(u.size)


def test_run_inference_routes_survival_mixture_to_the_fused_engine(cuda_device):
    from crispr_bean_b200 import survival_model as sm
    from crispr_bean_b200.run import make_engine

    data = _data(seed=23)
    eng = make_engine(sm.MixtureNormalModel, sm.MixtureNormalGuide, data, num_steps=10, device=cuda_device)
    assert isinstance(eng, SurvivalFusedEngine)
    eng.run(10)
    assert torch.isfinite(eng.losses()).all()
    assert set(eng.params()) == {"mu_loc", "mu_scale", "alpha_pi", "q0"}
