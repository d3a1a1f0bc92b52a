"""`write_result_table` mirror vs the reference's own function (bean/model/readwrite.py, imported in place where the
reference sources are mounted) and vs committed golden tables elsewhere."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from crispr_bean_b200.readwrite import non_overlap, write_result_table
from tests.helpers import GOLDEN


def make_inputs(seed, T=60, two_d=True, with_negctrl=True, with_alpha=True):
    g = torch.Generator().manual_seed(seed)
    shape = (T, 1) if two_d else (T,)
    params = {"mu_loc": torch.randn(shape, generator=g), "mu_scale": torch.rand(shape, generator=g) * 0.5 + 0.05,
              "sd_loc": 0.1 * torch.randn(shape, generator=g), "sd_scale": torch.rand(shape, generator=g) * 0.1 + 0.01}
    G = 3 * T
    if with_alpha:
        params["alpha_pi"] = torch.rand((G, 2), generator=g) + 0.1
        params["noise_scale"] = torch.rand((G,), generator=g) + 0.2
    neg = {"mu_loc": torch.tensor(0.03), "mu_scale": torch.tensor(0.2), "sd_loc": torch.tensor(0.1), "sd_scale": torch.tensor(0.05)} if with_negctrl else None
    target_info = pd.DataFrame({"n_guides": np.arange(T) % 5 + 1}, index=pd.Index([f"v{i}" for i in range(T)], name="target"))
    guide_info = pd.DataFrame({"target": [f"v{i // 3}" for i in range(G)]}, index=pd.Index([f"g{i}" for i in range(G)], name="name"))
    acc = (torch.rand(G, generator=g) * 5 + 0.5).numpy()
    return params, neg, target_info, guide_info, acc


CASES = [dict(seed=1), dict(seed=2, two_d=False), dict(seed=3, with_negctrl=False), dict(seed=4, with_alpha=False),
         dict(seed=5, n_neg=5), dict(seed=6, adjust=False), dict(seed=7, survival=True)]


def run(fn, case, tmp):
    case = dict(case)
    n_neg, adjust, survival = case.pop("n_neg", 15), case.pop("adjust", True), case.pop("survival", False)
    params, neg, target_info, guide_info, acc = make_inputs(**case)
    if survival:
        params.pop("sd_loc"), params.pop("sd_scale")
    os.makedirs(tmp, exist_ok=True)
    out = fn(target_info.copy(), guide_info.copy(), params, "MixtureNormal", prefix=f"{tmp}/", negctrl_params=neg,
             adjust_confidence_by_negative_control=adjust, adjust_confidence_negatives=np.arange(n_neg),
             guide_acc=acc if "alpha_pi" in params else None, sd_is_fitted=not survival, return_result=True,
             is_survival_screen=survival)
    guides = pd.read_csv(f"{tmp}/bean_sgRNA_result.MixtureNormal.csv", index_col=0)
    return out, guides


def test_non_overlap_matches_statistics_normaldist():
    from statistics import NormalDist

    rng = np.random.default_rng(0)
    mu = np.concatenate([rng.normal(size=200), [0.0, 0.0, 1.5, -2.0]])
    sd = np.concatenate([rng.uniform(0.05, 3.0, size=200), [1.0, 2.0, 1.0, 1.0]])
    ref = np.array([1 - NormalDist(m, s).overlap(NormalDist(0, 1)) for m, s in zip(mu, sd)])
    assert np.abs(non_overlap(mu, sd) - ref).max() < 1e-12


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_equals_reference_write_result_table(case, tmp_path):
    from tests.refharness import available, load_reference

    if not available():
        pytest.skip("reference sources not mounted")
    ref_fn = load_reference().readwrite.write_result_table
    ref, ref_guides = run(ref_fn, case, str(tmp_path / "ref"))
    got, got_guides = run(write_result_table, case, str(tmp_path / "ours"))
    assert list(got.columns) == list(ref.columns)
    assert list(got.index) == list(ref.index)  # identical ranking (rows are sorted by |z|)
    for c in ref.columns:
        if ref[c].dtype.kind in "fi":
            assert np.allclose(got[c].to_numpy(dtype=float), ref[c].to_numpy(dtype=float), rtol=1e-12, atol=1e-13), c
        else:
            assert (got[c] == ref[c]).all(), c
    assert list(got_guides.columns) == list(ref_guides.columns)
    for c in ref_guides.columns:
        if ref_guides[c].dtype.kind in "fi":
            assert np.allclose(got_guides[c], ref_guides[c], rtol=1e-12), c


def test_committed_golden_table():
    """Portable pin (runs on the GPU box too): the reference's table for case 1, frozen by make_reference_golden.py."""
    path = os.path.join(GOLDEN, "ref_result_table.csv")
    ref = pd.read_csv(path, index_col=0)
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        got, _ = run(write_result_table, CASES[0], tmp)
    assert list(got.columns) == list(ref.columns)
    assert list(got["target"]) == list(ref["target"])
    for c in ref.columns:
        if ref[c].dtype.kind in "fi":
            assert np.allclose(got[c].to_numpy(dtype=float), ref[c].to_numpy(dtype=float), rtol=1e-9, atol=1e-12), c
