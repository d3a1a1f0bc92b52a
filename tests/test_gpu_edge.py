"""Edge cases of the fused SVI step (`bean_svi_run_*`) against the CPU oracle: maximum sizes of the C-ABI
(B = 8 bins, R * B = 64 cells per guide), degenerate shapes (one guide, one replicate, one-guide variants, sizes that are
not multiples of the warp / CTA), masks (whole guides, whole replicates, samples), the count threshold, and the
Dirichlet-gradient regimes that go through the deferred per-warp queue (tiny and huge concentrations, draws at the
clamps) -- including a warp in which EVERY draw is deferred, so the queue flushes mid-loop."""
import numpy as np
import pytest
import torch

from crispr_bean_b200.data_class import VariantSortingReporterScreenData
from crispr_bean_b200.svi import SviEngine, VAR_PARAM_NAMES
from crispr_bean_b200.synth import make_sorting_screen
from tests import helpers as H
from tests.test_gpu_svi import check_grads, oracle_at, rel_err

pytestmark = pytest.mark.gpu
CASES = [(torch.float64, 1e-9, 1e-9), (torch.float32, 1e-5, 2e-4)]


def bins(n):
    edges = np.linspace(0.0, 1.0, n + 1)
    return tuple((float(edges[i]), float(edges[i + 1])) for i in range(n))


@pytest.mark.parametrize("dtype,tol,tol_alpha", CASES)
def test_max_bins_and_cells(cuda_device, dtype, tol, tol_alpha):
    """7 sort bins + bulk pseudo-bin = BEAN_MAX_BINS, 8 replicates -> R * B = 64 = BEAN_MAX_RB."""
    scr = make_sorting_screen(25, 3, n_reps=8, bins=bins(7), seed=31, depth=400.0)
    data = VariantSortingReporterScreenData(scr, control_can_be_selected=True)
    assert data.n_condits == 8 and data.n_reps * data.n_condits == 64
    check_grads("MixtureNormal", data, cuda_device, dtype, tol, tol_alpha, perturb_seed=1)


def test_sizes_beyond_the_abi_are_refused(cuda_device):
    from crispr_bean_b200._lib import BeanError

    scr = make_sorting_screen(5, 3, n_reps=9, bins=bins(7), seed=32)  # R * B = 72 > 64
    data = VariantSortingReporterScreenData(scr, control_can_be_selected=True)
    with pytest.raises(BeanError):
        SviEngine(data, "MixtureNormal", cuda_device, num_steps=2).run(1)


@pytest.mark.parametrize("dtype,tol,tol_alpha", CASES)
@pytest.mark.parametrize("shape", ["one_guide", "one_replicate", "single_guide_variants", "odd_sizes"])
def test_degenerate_shapes(cuda_device, dtype, tol, tol_alpha, shape):
    if shape == "one_guide":
        data = H.make_small_mixture_data(n_variants=6, n_reps=3, seed=2)[[9]]
    elif shape == "one_replicate":
        data = VariantSortingReporterScreenData(make_sorting_screen(20, 3, n_reps=1, seed=33, depth=200.0), control_can_be_selected=True)
    elif shape == "single_guide_variants":
        data = VariantSortingReporterScreenData(make_sorting_screen(37, 1, n_reps=3, seed=34, depth=200.0), control_can_be_selected=True)
    else:  # 131 guides: one full CTA + 3 lanes of the next; 33 variants: straddles the variant kernel's 32-variant CTA
        data = VariantSortingReporterScreenData(make_sorting_screen(32, 4, n_reps=2, seed=35, depth=200.0, n_negctrl_guides=3),
                                                control_can_be_selected=True)
        assert data.n_guides == 131 and data.n_targets == 33
    if shape == "one_replicate":
        tol_alpha = max(tol_alpha, 1e-8)
    check_grads("MixtureNormal", data, cuda_device, dtype, tol, tol_alpha, perturb_seed=7)


@pytest.mark.parametrize("dtype,tol,tol_alpha", CASES)
def test_masks_and_threshold(cuda_device, dtype, tol, tol_alpha):
    data = H.make_small_mixture_data(n_variants=30, n_reps=4, seed=5)
    data.repguide_mask[:, 3] = False          # a guide masked in every replicate
    data.repguide_mask[2, :] = False          # a whole replicate masked
    data.repguide_mask[0, ::5] = False
    data.sample_mask = data.sample_mask.clone().bool()
    data.sample_mask[1, 0] = False            # a masked sample: its alpha is clamped to eps (utils.py:24)
    data.X_masked = data.X * data.sample_mask[:, :, None]
    data.X_bcmatch_masked = data.X_bcmatch * data.sample_mask[:, :, None]
    data.X_masked[3, :, 8] = torch.tensor([3.0, 2.0, 3.0, 1.0, 1.0])      # N = 10: at the threshold -> masked
    data.X_masked[3, :, 9] = torch.tensor([3.0, 2.0, 3.0, 1.0, 2.0])      # N = 11: counted
    data.X_bcmatch_masked[3, :, 10] = 0                                   # an all-zero row
    check_grads("MixtureNormal", data, cuda_device, dtype, tol, tol_alpha, perturb_seed=9)


@pytest.mark.parametrize("dtype,tol,tol_alpha", CASES)
def test_every_dirichlet_gradient_regime_and_queue_flush(cuda_device, dtype, tol, tol_alpha):
    """alpha_pi spread over 6 orders of magnitude and draws pushed to the clamps: boundary series, rational correction,
    lone saddle-point components, near-mean polynomial.  With pi_a0 forced small, every draw of every lane is deferred:
    8 replicates x 32 lanes = 256 requests per warp -> the 64-entry queue flushes 8 times inside the replicate loop."""
    data = H.make_small_mixture_data(n_variants=40, n_reps=8, with_bulk_bin=False, seed=6)
    G, R = data.n_guides, data.n_reps
    g = torch.Generator().manual_seed(3)
    noise = H.fixed_noise("MixtureNormal", data, seed=22)
    pi1 = torch.rand((R, 1, G), generator=g, dtype=torch.float64)
    pi1[:, :, ::7] = 1e-9                      # at the lower boundary
    pi1[:, :, 1::7] = 1 - 1e-7                 # at the upper boundary
    pi1[:, :, 2::7] = 0.5                      # near the mean for symmetric concentrations
    noise["pi"] = torch.stack([1 - pi1, pi1], dim=-1)
    if dtype == torch.float32:
        # float32-representable draws on both sides (components rounded separately, as a float32 sampler produces them)
        noise["pi"] = noise["pi"].float().double().clamp(min=1.2e-38, max=1 - 2.0 ** -24)
    for small in (False, True):
        d = data
        if small:
            import copy

            d = copy.copy(data)
            d.pi_a0 = data.pi_a0 * 0.05       # total concentration ~ 1: no draw is in the saddle-point regime
        eng = SviEngine(d, "MixtureNormal", cuda_device, dtype=dtype, num_steps=4)
        au = 3.0 * torch.randn(eng.alpha_u.shape, generator=g, dtype=torch.float64)   # alpha_pi in e^-9 .. e^9
        eng.alpha_u.copy_(au)
        vp = 0.3 * torch.randn(eng.var_params.shape, generator=g, dtype=torch.float64)
        eng.var_params.copy_(vp)
        got = eng.gradients(noise)
        un = {k: vp[i] for i, k in enumerate(VAR_PARAM_NAMES)}
        un["alpha_pi"] = au
        ref_loss, ref = oracle_at("MixtureNormal", d, noise, un)
        assert abs(got["loss"].item() - ref_loss) <= tol * abs(ref_loss)
        for k in VAR_PARAM_NAMES:
            assert rel_err(got[k], ref[k]) <= tol, (k, small)
        assert rel_err(got["alpha_pi"], ref["alpha_pi"]) <= tol_alpha, small
        assert torch.isfinite(got["alpha_pi"]).all()


@pytest.mark.parametrize("dtype,tol,tol_alpha", CASES)
def test_long_and_short_guide_segments_in_one_warp(cuda_device, dtype, tol, tol_alpha):
    """The per-variant kernel sums a variant's guides in one thread up to 32 of them and with the whole warp beyond: variants
    with 45 guides (warp-cooperative), with 33 (just over) and the 6-guide control variant (sequential), all in one warp."""
    a = make_sorting_screen(5, 45, n_reps=2, seed=41, depth=200.0, n_negctrl_guides=6)
    data = VariantSortingReporterScreenData(a, control_can_be_selected=True)
    assert int(data.target_lengths.max()) == 45 and int(data.target_lengths.min()) == 6
    check_grads("MixtureNormal", data, cuda_device, dtype, tol, tol_alpha, perturb_seed=3)
    b = make_sorting_screen(3, 33, n_reps=2, seed=43, depth=200.0)
    check_grads("MixtureNormal", VariantSortingReporterScreenData(b, control_can_be_selected=True), cuda_device, dtype, tol, tol_alpha, perturb_seed=5)
