"""The CUDA SVI step against vectors computed by the REFERENCE's own model/guide programs
(tests/golden/ref_*.npz; generator: tests/golden/make_reference_golden.py).  No oracle in between:
loss and gradients w.r.t. the unconstrained parameters at the reference's initial parameters and recorded
reparameterisation draws, and a 6-step `run_inference` trajectory (ClippedAdam, lr decay).
Tolerances: 1e-9 fp64 / 1e-5 fp32 relative (north_star); alpha_pi fp32 2e-4 (see test_gpu_svi.py)."""
import ast
import os

import numpy as np
import pytest
import torch

from crispr_bean_b200.svi import SviEngine
from tests.helpers import GOLDEN
from tests.test_reference_golden import PROGRAMS, SORTING, edit_perm, group, load_case, oracle_kwargs, to_ours

pytestmark = pytest.mark.gpu
FUSED = list(PROGRAMS)  # sorting: fused SVI kernels; survival: bean_ll kernel inside the autograd engine


def rel(got, ref):
    got = got.detach().double().cpu().reshape(-1).numpy()
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    return float((np.abs(got - ref) / (np.abs(ref) + np.abs(ref).mean() + 1e-300)).max())


def make_engine(z, data, cuda_device, dtype, num_steps, autograd_engine=False):
    kw = oracle_kwargs(z)
    model = str(z["meta/oracle_model"])
    acc = dict(scale_by_accessibility=kw.get("scale_by_accessibility", False), fit_noise=kw.get("fit_noise", False))
    if getattr(data, "is_survival", False):
        from crispr_bean_b200.survival import SurvivalSviEngine

        extra = acc if model in ("MixtureNormal", "MultiMixtureNormal") else {}
        if model == "MixtureNormal" and not acc["scale_by_accessibility"] and not autograd_engine:
            from crispr_bean_b200.survival_fused import SurvivalFusedEngine  # what run_inference uses: the fused three-kernel step

            return SurvivalFusedEngine(data, cuda_device, dtype=dtype, num_steps=num_steps, use_bcmatch=kw.get("use_bcmatch", True))
        return SurvivalSviEngine(data, model, cuda_device, dtype=dtype, num_steps=num_steps, use_bcmatch=kw.get("use_bcmatch", True), **extra)
    if getattr(data, "sample_covariates", None) is not None:
        from crispr_bean_b200.generic import CovariateNormalEngine

        return CovariateNormalEngine(data, cuda_device, dtype=dtype, num_steps=num_steps, use_bcmatch=kw.get("use_bcmatch", True))
    if getattr(data, "is_tiling", False):
        from crispr_bean_b200.generic import TilingSviEngine

        return TilingSviEngine(data, cuda_device, dtype=dtype, num_steps=num_steps, **acc)
    return SviEngine(data, model, cuda_device, dtype=dtype, num_steps=num_steps, use_bcmatch=kw.get("use_bcmatch", True),
                     scale_by_accessibility=kw.get("scale_by_accessibility", False), fit_noise=kw.get("fit_noise", False),
                     prior_params=kw.get("prior_params"))


# fp64 kernels against the reference run in float64; fp32 kernels against the reference run in its own mixed
# float32/float64 arithmetic on ITS float32 draws (a float64 Dirichlet draw such as 1 - 2.5e-8 is not
# representable in float32 -- it rounds to exactly 1 -- so float64 draws cannot be replayed into fp32 kernels).
@pytest.mark.parametrize("dtype,tag,tol,tol_alpha", [(torch.float64, "f64", 1e-9, 1e-9), (torch.float32, "native", 1e-5, 2e-4)])
@pytest.mark.parametrize("name", FUSED)
def test_fused_step_equals_reference_programs(cuda_device, name, dtype, tag, tol, tol_alpha):
    z, data = load_case(name)
    eng = make_engine(z, data, cuda_device, dtype, 4)
    perm = edit_perm(z, data)
    noise = {k: torch.as_tensor(to_ours(v, perm, k)) for k, v in group(z, f"{tag}/noise/").items() if "/" not in k}
    got = eng.gradients(noise)
    ref_loss = float(z[f"{tag}/loss"])
    if name in ("control_normal_c1", "survival_control_normal") and dtype == torch.float32:
        tol = 2e-4  # one global (mu, sd): its gradient is a 40x-cancelling sum of per-guide terms
    assert abs(got["loss"].item() - ref_loss) <= tol * abs(ref_loss), (got["loss"].item(), ref_loss)
    ref = group(z, f"{tag}/grad/")
    assert set(ref) <= set(got), (sorted(ref), sorted(got))
    ref = {k: to_ours(g, perm, k) for k, g in ref.items()}
    errs = {k: rel(got[k], g) for k, g in ref.items()}
    print(name, tag, "loss", abs(got["loss"].item() - ref_loss) / abs(ref_loss), errs)
    for k, g in ref.items():
        e = rel(got[k], g)
        assert e <= (tol_alpha if k == "alpha_pi" else tol), f"{k}: {e:.3e}"


@pytest.mark.parametrize("name", [c for c in FUSED if "traj/n_steps" in np.load(os.path.join(GOLDEN, f"ref_{c}.npz")).files])
def test_fused_run_follows_reference_run_inference(cuda_device, name):
    z, data = load_case(name)
    n = int(z["traj/n_steps"])
    eng = make_engine(z, data, cuda_device, torch.float64, n)
    perm = edit_perm(z, data)
    tn = {k: to_ours(v, perm, k) for k, v in group(z, "traj/noise/").items()}
    for t in range(n):
        eng.run(1, noise={k: torch.as_tensor(v[t]) for k, v in tn.items()})
    loss = eng.losses().numpy()
    print(name, "traj loss err", np.abs(loss - z["traj/loss"]).max() / np.abs(z["traj/loss"]).max())
    assert np.abs(loss - z["traj/loss"]).max() <= 1e-9 * np.abs(z["traj/loss"]).max()
    params = eng.params()
    for k, v in group(z, "traj/param/").items():
        assert rel(params[k], to_ours(v, perm, k)) <= 1e-8, k
