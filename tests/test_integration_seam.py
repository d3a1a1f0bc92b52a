"""INTEGRATION.md section A executed in two stages (the reference sources live only in the build container, the GPU only on
the GPU box):

  stage 1 (CPU, reference mounted): the reference's unmodified MixtureNormalModel / MixtureNormalGuide run on the pyro shim with
      the count-likelihood block replaced by `pyro.factor("guide_counts", count_ll(data, mu, sd, pi))`; loss and every gradient
      equal the golden vectors of the UNPATCHED reference program (ref_mixture_small.npz).  The seam's inputs, value and
      input-gradients are recorded in tests/golden/seam_mixture_small.npz (generator: tests/golden/make_seam_fixture.py).
  stage 2 (GPU): the CUDA seam (`ll_function.count_log_likelihood` = the torch.autograd.Function over bean_ll_*) returns the
      recorded value and, through autograd, the recorded gradients for the recorded inputs.

Chain rule: program(parameters) -> (mu, sd, pi) is the reference's own code in both stages; (mu, sd, pi) -> ll and its gradient
are equal by stage 2; so the patched program on the GPU has the golden loss and gradients.
"""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN
from tests.refharness import available
from tests.test_reference_golden import group, load_case

FIXTURE = os.path.join(GOLDEN, "seam_mixture_small.npz")


@pytest.mark.skipif(not available(), reason="reference sources not mounted")
def test_stage1_reference_program_with_the_seam_reproduces_the_golden_vectors():
    from tests.golden.make_seam_fixture import run

    z, out, rec = run(write=False)
    assert abs(float(out["loss"]) - float(z["f64/loss"])) <= 1e-11 * abs(float(z["f64/loss"]))
    ref = group(z, "f64/grad/")
    assert set(ref) == {k[5:] for k in out if k.startswith("grad/")}
    for k, g in ref.items():
        assert np.abs(out[f"grad/{k}"] - g).max() <= 1e-9 * np.abs(g).max(), k
    # the committed fixture is what this run records
    fx = np.load(FIXTURE)
    for k in ("mu", "sd", "pi", "ll", "d_mu", "d_sd", "d_pi"):
        assert np.allclose(fx[k], rec[k], rtol=1e-12, atol=0), k


def test_fixture_inputs_are_the_programs_draws():
    """Sanity of the fixture without the reference: pi is the golden case's recorded Dirichlet draw, allele 0 is the wild
    type (mu 0, sd 1) and allele 1 carries one (mu, sd) per variant."""
    z, data = load_case("mixture_small")
    fx = np.load(FIXTURE)
    assert np.array_equal(fx["pi"], z["f64/noise/pi"])
    assert (fx["mu"][:, 0] == 0).all() and (fx["sd"][:, 0] == 1).all()
    assert len(np.unique(fx["mu"][:, 1])) == data.n_targets


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-5)])
def test_stage2_cuda_seam_returns_the_recorded_value_and_gradients(cuda_device, dtype, tol):
    from crispr_bean_b200.device_pack import DeviceScreen, pi_to_guide_major
    from crispr_bean_b200.ll_function import count_log_likelihood

    _, data = load_case("mixture_small")
    fx = np.load(FIXTURE)
    scr = DeviceScreen(data, cuda_device, dtype=dtype, use_bcmatch=True, mask_thres=10)
    mu = torch.as_tensor(fx["mu"]).to(cuda_device, dtype).requires_grad_(True)
    sd = torch.as_tensor(fx["sd"]).to(cuda_device, dtype).requires_grad_(True)
    pi = torch.as_tensor(fx["pi"]).to(cuda_device, dtype).requires_grad_(True)
    ll = count_log_likelihood(scr, mu, sd, pi_to_guide_major(pi), None)  # the seam call of INTEGRATION.md section A
    assert abs(ll.item() - float(fx["ll"])) <= tol * abs(float(fx["ll"]))
    (3.0 * ll).backward()  # an upstream factor, as pyro's -ELBO scaling would apply
    for name, t in (("d_mu", mu), ("d_sd", sd), ("d_pi", pi)):
        ref = 3.0 * fx[name]
        err = np.abs(t.grad.double().cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err <= 10 * tol, (name, err)
