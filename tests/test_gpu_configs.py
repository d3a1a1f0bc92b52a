"""Parity at the shapes BASELINE.json names (SURVEY section 8 shape table): c2 LDL-C variant sorting, c3 CDS tiling,
c4 survival -- CUDA path vs the CPU oracle on the same seeded inputs and injected draws -- and, at c5's FULL size
(1M guides x 8 replicates x 4 bins), size-independent properties: shard additivity with global Philox ids,
determinism, and a recomputation of a random sample of guides by the oracle.
Tolerances as everywhere: 1e-9 fp64 / 1e-5 fp32 relative (alpha_pi fp32: 2e-4)."""
import numpy as np
import pytest
import torch

from crispr_bean_b200.data_class import (TilingSortingReporterScreenData, VariantSortingReporterScreenData,
                                         VariantSurvivalReporterScreenData)
from crispr_bean_b200.svi import SviEngine, VAR_PARAM_NAMES
from crispr_bean_b200.synth import make_config, make_survival_screen, make_tiling_screen
from oracle import bean_oracle as O
from tests import helpers as H
from tests.test_gpu_svi import check_grads, rel_err

pytestmark = pytest.mark.gpu
CASES = [(torch.float64, 1e-9, 1e-9), (torch.float32, 1e-5, 2e-4)]


@pytest.mark.parametrize("dtype,tol,tol_alpha", CASES)
def test_c2_ldlc_variant_shape(cuda_device, dtype, tol, tol_alpha):
    """~3.5k guides, 690 variants + 101 negative-control guides, ragged guides/variant, 4 replicates, 4 bins + bulk."""
    data = VariantSortingReporterScreenData(make_config("c2_ldlc_variant", seed=101), control_can_be_selected=True)
    assert 3000 < data.n_guides < 6000 and data.n_condits == 5 and data.n_reps == 4
    check_grads("MixtureNormal", data, cuda_device, dtype, tol, tol_alpha, perturb_seed=4)


@pytest.mark.parametrize("dtype,tol,tol_alpha", [(torch.float64, 1e-9, 1e-9), (torch.float32, 1e-5, 1e-5)])
def test_c3_tiling_shape(cuda_device, dtype, tol, tol_alpha):
    """~800 guides x up to 16 alleles per guide (MultiMixtureNormal over filtered alleles)."""
    from crispr_bean_b200.generic import TilingSviEngine

    scr = make_tiling_screen(n_guides=800, max_alleles=16, n_reps=4, seed=3)
    data = TilingSortingReporterScreenData(scr, control_can_be_selected=True, allele_df_key="allele_counts")
    assert data.n_guides == 800 and data.n_max_alleles >= 12
    noise = H.fixed_noise("MultiMixtureNormal", data, seed=5)
    eng = TilingSviEngine(data, cuda_device, dtype=dtype, num_steps=4)
    got = eng.gradients(noise)
    ref = H.oracle_loss_and_grads("MultiMixtureNormal", data, noise)
    assert abs(got["loss"].item() - ref["loss"]) <= tol * abs(ref["loss"])
    for k, g in ref["grads"].items():
        assert rel_err(got[k], g) <= (tol_alpha if k == "alpha_pi" else tol), k


@pytest.mark.parametrize("dtype,tol,tol_alpha", CASES)
def test_c4_survival_shape(cuda_device, dtype, tol, tol_alpha):
    """Proliferation screen, 3 timepoints (D0, D7, D14; control D7 stays selected), 3 replicates, ~3.5k guides."""
    from crispr_bean_b200.survival import SurvivalSviEngine

    scr = make_survival_screen(690, "lognormal", n_reps=3, seed=21, n_negctrl_guides=101)
    data = VariantSurvivalReporterScreenData(scr, control_condition="D7")
    assert data.n_condits == 3 and 3000 < data.n_guides < 6000
    g = torch.Generator().manual_seed(2)
    G, R, T = data.n_guides, data.n_reps, data.n_targets
    gam = torch._standard_gamma(torch.full((R, G), 1.3, dtype=torch.float64), generator=g)
    pig = torch._standard_gamma(torch.full((R, 1, G, 2), 1.5, dtype=torch.float64), generator=g)
    noise = {"eps_mu": torch.randn((T, 1), generator=g, dtype=torch.float64), "q0": gam / gam.sum(-1, keepdim=True),
             "pi": pig / pig.sum(-1, keepdim=True), "eps_negctrl": torch.randn((G,), generator=g, dtype=torch.float64)}
    eng = SurvivalSviEngine(data, "MixtureNormal", cuda_device, dtype=dtype, num_steps=4)
    got = eng.gradients(noise)
    with H.default_dtype(torch.float64):
        ps = O.ParamStore()
        loss, _ = O.elbo_survival_mixture_normal(H.cast_data(data, torch.float64), ps, noise=noise)
        loss.backward()
    assert abs(got["loss"].item() - float(loss.detach())) <= tol * abs(float(loss.detach()))
    for k, v in ps.unconstrained.items():
        assert rel_err(got[k], v.grad) <= (tol_alpha if k == "alpha_pi" else tol), k


@pytest.fixture(scope="module")
def c5_data():
    from bench import build_data

    return build_data("c5_genome_scale", seed=101)


def test_c5_full_size_properties(cuda_device, c5_data):
    """1M guides x 8 x 4 (the bench workload).  (1) determinism: two engines, same seed -> bit-identical loss and
    parameters; (2) shard additivity: the two halves, run with their global Philox ids, reproduce the full run's
    parameters bit for bit and its loss to 1e-12; (3) a random sample of 2,000 variants re-run by the CPU oracle on the
    kernel's own recorded draws agrees on loss and gradients."""
    from crispr_bean_b200.dist import shard_data

    data, steps = c5_data, 3
    assert data.n_guides == 1_000_000 and data.n_reps == 8 and data.n_condits == 4
    full = SviEngine(data, "MixtureNormal", cuda_device, num_steps=steps, seed=9)
    full.run(steps)
    again = SviEngine(data, "MixtureNormal", cuda_device, num_steps=steps, seed=9)
    again.run(steps)
    assert torch.equal(full.losses(), again.losses()) and torch.equal(full.var_params, again.var_params)
    assert torch.equal(full.alpha_u, again.alpha_u)
    del again
    mu, loss = [], 0
    for rank in range(2):
        sub, off = shard_data(data, rank, 2)
        eng = SviEngine(sub, "MixtureNormal", cuda_device, num_steps=steps, seed=9,
                        guide_offset=off["guide_offset"], variant_offset=off["variant_offset"])
        eng.run(steps)
        mu.append(eng.var_params.clone())
        loss = loss + eng.losses()
        del eng
    assert torch.equal(torch.cat(mu, dim=1), full.var_params)
    torch.testing.assert_close(loss, full.losses(), rtol=1e-12, atol=0)
    # (3) oracle on a sample: variants [v0, v0 + 2000) with the draws the kernel used for them
    v0, nv = 123_400, 2_000
    gsel = np.arange(int(data.variant_ptr[v0]), int(data.variant_ptr[v0 + nv]))
    sub = data[gsel]
    eng = SviEngine(sub, "MixtureNormal", cuda_device, num_steps=2, seed=9, guide_offset=int(gsel[0]), variant_offset=v0)
    got = eng.gradients({"record": True})
    eps, pi = eng.eps_used.double().cpu(), eng.pi_used.double().cpu()
    noise = {"eps_mu": eps[0].reshape(-1, 1), "eps_sd": eps[1].reshape(-1, 1), "pi": pi.permute(1, 0, 2).unsqueeze(1)}
    ref = H.oracle_loss_and_grads("MixtureNormal", sub, noise)
    assert abs(got["loss"].item() - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    for k in VAR_PARAM_NAMES:
        assert rel_err(got[k], ref["grads"][k]) <= 1e-5, k
    assert rel_err(got["alpha_pi"], ref["grads"]["alpha_pi"]) <= 2e-4
    # and the sub-screen's draws are the full screen's draws for those guides (global counter ids)
    full2 = SviEngine(data, "MixtureNormal", cuda_device, num_steps=2, seed=9)
    full2.gradients({"record": True})
    assert torch.equal(full2.pi_used[gsel[0]:gsel[-1] + 1], eng.pi_used)
