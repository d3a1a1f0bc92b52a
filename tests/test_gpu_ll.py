"""Parity of the count-likelihood kernel (C-ABI `bean_ll_f32/f64`) against the CPU oracle.

Tolerances (BASELINE.json north_star): per-guide log-likelihood and gradients within 1e-9 relative in
fp64 and 1e-5 relative in fp32.  "Relative" is taken against the magnitude of the quantity's own
vector (|err| <= tol * (|ref| + mean|ref|)), since individual entries can legitimately be ~0.
"""
import numpy as np
import pytest
import torch

from crispr_bean_b200.device_pack import DeviceScreen, pi_to_guide_major
from crispr_bean_b200.ll_function import count_log_likelihood, launch_ll
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-9, torch.float32: 1e-5}


def rel_close(got, ref, tol, what):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    scale = ref.abs() + ref.abs().mean() + 1e-300
    err = ((got - ref).abs() / scale).max().item()
    assert err <= tol, f"{what}: max relative error {err:.3e} > {tol:.1e}"
    return err


def random_inputs(data, A, seed, allele_mask=None):
    g = torch.Generator().manual_seed(seed)
    G, R = data.n_guides, data.n_reps
    mu = torch.randn((G, A), generator=g, dtype=torch.float64)
    sd = torch.rand((G, A), generator=g, dtype=torch.float64) * 1.5 + 0.4
    mu[:, 0], sd[:, 0] = 0.0, 1.0
    gam = torch._standard_gamma(torch.full((R, 1, G, A), 1.2, dtype=torch.float64), generator=g)
    if allele_mask is not None:
        gam = gam * allele_mask[None, None] + 1e-12
    pi = gam / gam.sum(-1, keepdim=True)
    return mu, sd, pi


def run_case(data, cuda_device, dtype, A=2, use_pi=True, seed=0, allele_mask=None, use_bcmatch=True):
    mu, sd, pi = random_inputs(data, A, seed, allele_mask)
    ref = H.oracle_ll_core(data, mu, sd, pi if use_pi else None, dtype=torch.float64, use_bcmatch=use_bcmatch,
                           allele_mask=allele_mask)
    scr = DeviceScreen(data, cuda_device, dtype=dtype, use_bcmatch=use_bcmatch)
    out = launch_ll(scr, mu.to(cuda_device), sd.to(cuda_device),
                    pi_to_guide_major(pi).to(cuda_device) if use_pi else None,
                    allele_mask.to(cuda_device) if allele_mask is not None else None, want_rows=True)
    tol = TOL[dtype]
    # per-guide log-likelihood: sum over replicates and layers of the masked rows
    aux = ref["aux"]
    ref_rows = [aux["ll_guide_counts"] * aux["w_guide_counts"]]
    if use_bcmatch:
        ref_rows.append(aux["ll_guide_bcmatch_counts"] * aux["w_guide_bcmatch_counts"])
    ref_per_guide = torch.stack(ref_rows).sum(0).sum(0)  # (G,)
    got_per_guide = out["ll_row"].sum(dim=(0, 1))
    errs = {"ll_guide": rel_close(got_per_guide, ref_per_guide, tol, "per-guide ll")}
    assert abs(out["ll"].item() - ref["ll"]) <= tol * abs(ref["ll"])
    errs["d_mu"] = rel_close(out["d_mu"], ref["d_mu"], tol, "d_mu")
    errs["d_sd"] = rel_close(out["d_sd"], ref["d_sd"], tol, "d_sd")
    if use_pi:
        errs["d_pi"] = rel_close(out["d_pi"], pi_to_guide_major(ref["d_pi"]), tol, "d_pi")
    return errs


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("with_bulk", [True, False])
def test_mixture_shape(cuda_device, dtype, with_bulk):
    data = H.make_small_mixture_data(n_variants=40, n_reps=3, with_bulk_bin=with_bulk)
    print(run_case(data, cuda_device, dtype))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_c5_row_shape_8x4(cuda_device, dtype):
    data = H.make_small_mixture_data(n_variants=300, n_reps=8, with_bulk_bin=False, seed=8)
    assert data.n_reps * data.n_condits == 32
    print(run_case(data, cuda_device, dtype))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_normal_model_single_allele_no_pi(cuda_device, dtype):
    data = H.load_var_mini()  # the reference's CSV fixture: 30 guides, 2 reps, 4 bins + bulk
    g = torch.Generator().manual_seed(1)
    mu = torch.randn((30, 1), generator=g, dtype=torch.float64)
    sd = torch.rand((30, 1), generator=g, dtype=torch.float64) + 0.5
    ref = H.oracle_ll_core(data, mu, sd, None, use_bcmatch=False)
    scr = DeviceScreen(data, cuda_device, dtype=dtype, use_bcmatch=False)
    out = launch_ll(scr, mu.to(cuda_device), sd.to(cuda_device), None)
    assert abs(out["ll"].item() - ref["ll"]) <= TOL[dtype] * abs(ref["ll"])
    rel_close(out["d_mu"], ref["d_mu"], TOL[dtype], "d_mu")
    rel_close(out["d_sd"], ref["d_sd"], TOL[dtype], "d_sd")


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_many_alleles_with_mask(cuda_device, dtype):
    data = H.make_small_mixture_data(n_variants=25, n_reps=4)
    A = 6
    g = torch.Generator().manual_seed(3)
    amask = torch.rand((data.n_guides, A), generator=g) < 0.6
    amask[:, 0] = True
    print(run_case(data, cuda_device, dtype, A=A, allele_mask=amask, seed=5))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("A,n_variants", [(8, 25), (45, 17), (231, 6)])
def test_warp_per_guide_kernel_many_alleles(cuda_device, dtype, A, n_variants):
    """A >= 8: a warp owns a guide (raw tiling allele tables: the reference's tiling_mini_screen has up to 231
    alleles per guide).  Ragged existence masks, A not a multiple of 32, G not a multiple of the 4 guides per CTA."""
    data = H.make_small_mixture_data(n_variants=n_variants, n_reps=3, seed=A)
    g = torch.Generator().manual_seed(A)
    n_exist = torch.randint(1, A + 1, (data.n_guides,), generator=g)
    amask = torch.arange(A)[None, :] < n_exist[:, None]
    print(run_case(data, cuda_device, dtype, A=A, allele_mask=amask, seed=7))


def test_thread_and_warp_per_guide_kernels_agree(cuda_device):
    """The same guides scored with 6 alleles (thread per guide) and padded to 9 with non-existent alleles of zero
    weight (warp per guide): identical log-likelihood rows and gradients up to summation order."""
    data = H.make_small_mixture_data(n_variants=30, n_reps=4, seed=2)
    G = data.n_guides
    amask = torch.ones((G, 6), dtype=torch.bool)
    mu, sd, pi = random_inputs(data, 6, seed=9, allele_mask=amask)
    scr = DeviceScreen(data, cuda_device, dtype=torch.float64)
    narrow = launch_ll(scr, mu.to(cuda_device), sd.to(cuda_device), pi_to_guide_major(pi).to(cuda_device), amask.to(cuda_device), want_rows=True)
    pad = lambda t, v: torch.cat([t, torch.full(t.shape[:-1] + (3,), v, dtype=t.dtype)], dim=-1)
    wide = launch_ll(scr, pad(mu, 0.3).to(cuda_device), pad(sd, 1.1).to(cuda_device), pi_to_guide_major(pad(pi, 0.0)).to(cuda_device),
                     pad(amask, False).to(cuda_device), want_rows=True)
    assert torch.allclose(wide["ll_row"], narrow["ll_row"], rtol=1e-13, atol=1e-10)
    assert torch.allclose(wide["d_mu"][:, :6], narrow["d_mu"], rtol=1e-11, atol=1e-10)
    assert torch.allclose(wide["d_sd"][:, :6], narrow["d_sd"], rtol=1e-11, atol=1e-10)
    assert torch.allclose(wide["d_pi"][:, :, :6], narrow["d_pi"], rtol=1e-11, atol=1e-10)
    assert (wide["d_mu"][:, 6:] == 0).all() and (wide["d_pi"][:, :, 6:] == 0).all()  # non-existent alleles: P := 0


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_edge_cases_masks_zero_rows_tiny_counts(cuda_device, dtype):
    data = H.make_small_mixture_data(n_variants=20, n_reps=3)
    data.sample_mask = data.sample_mask.clone().bool()
    data.sample_mask[1, 2] = False  # a masked sample
    data.X_masked = data.X * data.sample_mask[:, :, None]
    data.X_bcmatch_masked = data.X_bcmatch * data.sample_mask[:, :, None]
    data.repguide_mask[0, ::3] = False  # masked (rep, guide) rows
    data.X_masked[2, :, 5] = 0  # an all-zero row (N = 0 <= thres -> masked by the count threshold)
    data.X_masked[2, :, 6] = torch.tensor([3.0, 2.0, 3.0, 1.0, 2.0])  # N = 11: just above the threshold
    data.X_masked[2, :, 7] = torch.tensor([3.0, 2.0, 3.0, 1.0, 1.0])  # N = 10: at the threshold -> masked
    print(run_case(data, cuda_device, dtype))


def test_single_guide_and_non_multiple_of_block(cuda_device):
    data = H.make_small_mixture_data(n_variants=33, n_reps=2)  # G = 6 + 132 = 138 (not a multiple of 128)
    print(run_case(data, cuda_device, torch.float64))
    one = data[[7]]
    print(run_case(one, cuda_device, torch.float64))


def test_autograd_function_scales_by_upstream_gradient(cuda_device):
    data = H.make_small_mixture_data(n_variants=10, n_reps=2)
    mu, sd, pi = random_inputs(data, 2, 0)
    scr = DeviceScreen(data, cuda_device, dtype=torch.float64)
    mu_d = mu.to(cuda_device).requires_grad_(True)
    sd_d = sd.to(cuda_device).requires_grad_(True)
    pi_d = pi_to_guide_major(pi).to(cuda_device).requires_grad_(True)
    ll = count_log_likelihood(scr, mu_d, sd_d, pi_d)
    (-2.5 * ll).backward()
    ref = H.oracle_ll_core(data, mu, sd, pi)
    rel_close(mu_d.grad, -2.5 * ref["d_mu"], 1e-9, "autograd d_mu")
    rel_close(pi_d.grad, -2.5 * pi_to_guide_major(ref["d_pi"]), 1e-9, "autograd d_pi")


def test_bad_arguments_return_error_codes(cuda_device):
    from crispr_bean_b200 import _lib

    data = H.make_small_mixture_data(n_variants=4, n_reps=2)
    scr = DeviceScreen(data, cuda_device, dtype=torch.float32)
    args = _lib.BeanLLArgs()
    args.n_alleles = 2
    assert _lib.lib().bean_ll_f32(scr.c, args, None) == -1  # BEAN_EINVAL: null mu/sd
    assert b"mu_allele" in _lib.lib().bean_last_error()
    with pytest.raises(_lib.BeanError):
        launch_ll(scr, torch.zeros((data.n_guides, 2)), torch.ones((data.n_guides, 2)))  # CPU tensors: no fallback


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_survival_mode_exp_growth(cuda_device, dtype):
    """mode = SURVIVAL: e = sum_a pi_a exp(mu_a t_b) (survival_model.py:352-424), timepoints 0, .5, 1."""
    from oracle import bean_oracle as O

    data = H.make_small_mixture_data(n_variants=30, n_reps=3, with_bulk_bin=False, bins=((0.0, 0.3), (0.3, 0.6), (0.6, 1.0)))
    data.is_survival = True
    data.timepoints = torch.tensor([0.0, 0.5, 1.0], dtype=torch.float64)
    G, R = data.n_guides, data.n_reps
    g = torch.Generator().manual_seed(4)
    mu = 0.8 * torch.randn((G, 2), generator=g, dtype=torch.float64)
    gam = torch._standard_gamma(torch.full((R, 1, G, 2), 1.5, dtype=torch.float64), generator=g)
    pi = gam / gam.sum(-1, keepdim=True)
    with H.default_dtype(torch.float64):
        d = H.cast_data(data, torch.float64)
        m, p_ = mu.clone().requires_grad_(True), pi.clone().requires_grad_(True)
        total, aux = O.survival_ll_core(d, m, p_, mask_thres=10)
        total.backward()
    scr = DeviceScreen(data, cuda_device, dtype=dtype)
    assert scr.mode == 1
    out = launch_ll(scr, mu.to(cuda_device), torch.ones_like(mu).to(cuda_device), pi_to_guide_major(pi).to(cuda_device))
    tol = TOL[dtype]
    assert abs(out["ll"].item() - float(total)) <= tol * abs(float(total))
    rel_close(out["d_mu"], m.grad, tol, "survival d_mu")
    rel_close(out["d_pi"], pi_to_guide_major(p_.grad), tol, "survival d_pi")
    assert float(out["d_sd"].abs().max()) == 0.0
