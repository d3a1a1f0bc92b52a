"""The tensoriser and the CPU oracle against vectors computed by the REFERENCE's own source files
(tests/golden/ref_*.npz, written by tests/golden/make_reference_golden.py through tests/refharness)."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from crispr_bean_b200 import data_class as dc
from oracle import bean_oracle as O
from tests.helpers import GOLDEN, default_dtype, cast_data
from tests.refharness.golden import screen_from_arrays

CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_*.npz")))
SORTING = [c for c in CASES if not c.startswith(("survival", "tiling"))]
SURVIVAL = [c for c in CASES if c.startswith("survival")]
TILING = [c for c in CASES if c.startswith("tiling")]
PROGRAMS = SORTING + SURVIVAL + TILING
EDIT_AXIS_KEYS = ("mu_loc", "mu_scale", "sd_loc", "sd_scale", "eps_mu", "eps_sd")  # (E,) arrays of the tiling model


def elbo_fn(name, z):
    table = O.SURVIVAL_ELBOS if name.startswith("survival") else O.SORTING_ELBOS  # (sorting tiling lives in SORTING_ELBOS)
    return table[str(z["meta/oracle_model"])]


def load_case(name, reference_fits=True):
    """Fixture + the screen pushed through OUR tensoriser.  `reference_fits`: take the three curve_fit products
    (a0, a0_bcmatch, pi_a0) from the fixture, so that program parity is not blurred by scipy's 1.5e-8
    Levenberg-Marquardt stopping tolerance (ours is the closed-form least-squares optimum)."""
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    scr = screen_from_arrays(z)
    cls = getattr(dc, str(z["meta/data_class"]))
    data = cls(scr, **ast.literal_eval(str(z["meta/data_kwargs"])))
    if reference_fits:
        for k in ("a0", "a0_bcmatch", "pi_a0"):
            if f"data/{k}" in z.files and hasattr(data, k):
                setattr(data, k, torch.as_tensor(z[f"data/{k}"]))
    return z, data


def edit_perm(z, data):
    """Tiling: position in OUR edit order of every edit of the reference's order.  The reference numbers edits in the
    iteration order of Python sets of `Edit` objects (preprocessing/utils.py:149-173) -- a labelling that changes
    from process to process -- so (E,) arrays are aligned through the edit keys before comparing."""
    if "meta/edit_index_keys" not in z.files:
        return None
    keys = [str(k) for k in z["meta/edit_index_keys"]]
    assert set(keys) == set(data.edit_index) and len(keys) == len(data.edit_index)
    return np.asarray([data.edit_index[k] for k in keys])


def to_ours(arr, perm, key):
    """(E,)-shaped reference array (or a stack of them, E last) -> our edit order."""
    if perm is None or key.split("/")[-1] not in EDIT_AXIS_KEYS:
        return arr
    out = np.empty_like(arr)
    out[..., perm] = arr
    return out


def group(z, prefix):
    return {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)}


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_golden_cases_exist():
    assert len(SORTING) >= 9 and len(SURVIVAL) >= 6


@pytest.mark.parametrize("name", CASES)
def test_tensoriser_equals_reference_data_class(name):
    """Every tensor the reference's *ScreenData built from the screen (data_class.py) -- same values and dtypes."""
    z, data = load_case(name, reference_fits=False)
    ref = group(z, "data/")
    checked = 0
    for k, v in ref.items():
        if not hasattr(data, k):
            continue
        mine = getattr(data, k)
        if torch.is_tensor(mine):
            assert tuple(mine.shape) == v.shape, k
            assert str(mine.dtype).replace("torch.", "") == str(v.dtype), (k, mine.dtype, v.dtype)
            if k in ("a0", "a0_bcmatch", "pi_a0"):  # closed-form OLS here vs curve_fit (ftol 1.5e-8) there
                assert rel(mine.numpy(), v) < 1e-7, k
            elif k == "allele_to_edit":
                assert np.array_equal(mine.numpy()[:, :, edit_perm(z, data)], v), k
            else:
                assert np.array_equal(mine.numpy(), v, equal_nan=True), k
        else:
            assert mine == v.item(), k
        checked += 1
    assert checked >= 12
    missing = [k for k in ref if not hasattr(data, k) and ref[k].ndim > 0]
    assert not missing, f"tensors of the reference data class absent from the mirror: {missing}"


def oracle_kwargs(z, dtype=torch.float64):
    kw = ast.literal_eval(str(z["meta/oracle_kwargs"]))
    if kw.get("prior_params") == "FROM_FIXTURE":  # per-variant prior tensors stored in the fixture
        kw["prior_params"] = {k: torch.as_tensor(v) for k, v in group(z, "prior/").items()}
    return kw


def oracle_eval(z, data, tag, dtype, name=""):
    perm = edit_perm(z, data)
    # draws keep the dtype the reference produced them in (native run: pi is float64 because pi_a0 is)
    noise = {k: torch.as_tensor(to_ours(v, perm, k)) for k, v in group(z, f"{tag}/noise/").items() if "/" not in k}
    kw = oracle_kwargs(z)
    with default_dtype(dtype):
        d = cast_data(data, dtype) if dtype == torch.float64 else data
        ps = O.ParamStore()
        loss, aux = elbo_fn(name, z)(d, ps, noise=noise, **kw)
        loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)).detach().double().numpy() for k, v in ps.unconstrained.items()}
    return float(loss.detach()), grads, ps


@pytest.mark.parametrize("name", PROGRAMS)
def test_oracle_equals_reference_programs_float64(name):
    """-ELBO and d(-ELBO)/d(unconstrained params) of the reference model/guide programs, float64, same noise."""
    z, data = load_case(name)
    loss, grads, _ = oracle_eval(z, data, "f64", torch.float64, name)
    assert abs(loss - float(z["f64/loss"])) <= 1e-11 * abs(float(z["f64/loss"]))
    ref_grads = group(z, "f64/grad/")
    assert set(grads) == set(ref_grads)
    perm = edit_perm(z, data)
    for k, g in ref_grads.items():
        assert rel(grads[k].reshape(g.shape), to_ours(g, perm, k)) < 1e-9, k


@pytest.mark.parametrize("name", PROGRAMS)
def test_oracle_equals_reference_programs_native_precision(name):
    """Same in the reference's own mixed float32/float64 arithmetic (tolerance: float32 rounding)."""
    z, data = load_case(name)
    loss, grads, _ = oracle_eval(z, data, "native", torch.float32, name)
    assert abs(loss - float(z["native/loss"])) <= 2e-6 * abs(float(z["native/loss"]))
    perm = edit_perm(z, data)
    for k, g in group(z, "native/grad/").items():
        assert rel(grads[k].reshape(g.shape), to_ours(g, perm, k)) < 2e-4, k


@pytest.mark.parametrize("name", [c for c in PROGRAMS if "traj/n_steps" in np.load(os.path.join(GOLDEN, f"ref_{c}.npz")).files])
def test_oracle_run_inference_follows_reference_trajectory(name):
    """bean/model/run.py:run_inference (SVI + ClippedAdam, lr decay) for a few steps with the recorded draws."""
    z, data = load_case(name)
    n = int(z["traj/n_steps"])
    perm = edit_perm(z, data)
    tn = {k: to_ours(v, perm, k) for k, v in group(z, "traj/noise/").items()}
    kw = oracle_kwargs(z)
    with default_dtype(torch.float64):
        d = cast_data(data, torch.float64)
        ps, hist = O.run_inference(elbo_fn(name, z), d, num_steps=n,
                                   noise_fn=lambda t: {k: torch.as_tensor(v[t]) for k, v in tn.items()}, **kw)
    assert rel(hist["loss"], z["traj/loss"]) < 1e-11
    for k, v in group(z, "traj/param/").items():
        assert rel(hist["params"][k].numpy().reshape(v.shape), to_ours(v, perm, k)) < 1e-9, k


def test_live_reference_matches_committed_vectors():
    """Where the reference sources are mounted (build container): re-run one case live -- guards the vectors
    against drifting from the generator."""
    from tests.refharness import available, load_reference

    if not available():
        pytest.skip("reference sources not mounted")
    import copy
    import warnings
    from tests.refharness import golden as G

    warnings.filterwarnings("ignore")
    ns = load_reference()
    z = np.load(os.path.join(GOLDEN, "ref_mixture_small.npz"))
    scr = screen_from_arrays(z)
    data = ns.data_class.VariantSortingReporterScreenData(copy.deepcopy(scr), **ast.literal_eval(str(z["meta/data_kwargs"])))
    out, _ = G.reference_loss_and_grads(ns.pyro, ns.model.MixtureNormalModel, ns.model.MixtureNormalGuide, data, seed=11,
                                        dtype=torch.float64)
    assert float(out["loss"]) == float(z["f64/loss"])
    assert np.array_equal(out["grad/alpha_pi"], z["f64/grad/alpha_pi"])
