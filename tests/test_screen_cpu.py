"""Host containers of the path: `MiniScreen` (the attribute surface `bean run` reads from a ReporterScreen)
and the prior resolution of the latent-site kernel (`--prior-params`, bean/model/run.py:480-542)."""
import numpy as np
import pandas as pd
import pytest
import torch

from crispr_bean_b200.latent_sites import LatentPrior
from crispr_bean_b200.screen import MiniScreen, _take_columns


def make(G=7, S=6, big=False):
    rng = np.random.default_rng(0)
    X = rng.poisson(20, (G, S)).astype(np.float32)
    guides = pd.DataFrame({"target": [f"v{g // 2}" for g in range(G)]}, index=pd.Index([f"g{g}" for g in range(G)], name="name"))
    samples = pd.DataFrame({"replicate": [f"r{s // 3}" for s in range(S)], "condition": ["a", "b", "c"] * (S // 3)},
                           index=[f"r{s // 3}_{'abc'[s % 3]}" for s in range(S)])
    per_guide = pd.DataFrame({"m": np.arange(G)}, index=guides.index)
    return MiniScreen(X, guides, samples, {"X_bcmatch": X // 2}, {"repguide_mask": per_guide, "tiling": False})


def test_sample_subset_of_the_whole_library_gathers_columns_row_major():
    scr = make()
    sub = scr[:, scr.samples["condition"] != "b"]
    assert sub.X.shape == (7, 4) and sub.X.flags.c_contiguous  # float32 column means depend on the layout (a0 fit)
    assert np.array_equal(sub.X, scr.X[:, [0, 2, 3, 5]]) and np.array_equal(sub.layers["X_bcmatch"], scr.layers["X_bcmatch"][:, [0, 2, 3, 5]])
    assert list(sub.samples.index) == ["r0_a", "r0_c", "r1_a", "r1_c"] and len(sub.guides) == 7
    assert sub.uns["repguide_mask"].equals(scr.uns["repguide_mask"])


def test_copy_owns_its_tables_and_shares_the_count_matrices():
    scr = make()
    cp = scr.copy()
    cp.samples["mask"] = 1
    cp.guides["extra"] = 0
    assert "mask" not in scr.samples.columns and "extra" not in scr.guides.columns
    assert cp.X is scr.X or np.shares_memory(cp.X, scr.X)  # never written in place by the tensoriser
    ident = scr[:, np.arange(6)]  # an identity reorder is recognised
    assert np.shares_memory(ident.X, scr.X)


def test_guide_subset_follows_per_guide_tables():
    scr = make()
    sub = scr[[1, 4, 5], :]
    assert np.array_equal(sub.X, scr.X[[1, 4, 5]]) and list(sub.guides.index) == ["g1", "g4", "g5"]
    assert list(sub.uns["repguide_mask"]["m"]) == [1, 4, 5]
    by_name = scr[np.asarray(["g4", "g1"]), np.asarray([True, False, False, True, False, False])]
    assert np.array_equal(by_name.X, scr.X[[4, 1]][:, [0, 3]])


def test_large_matrices_take_the_threaded_gather():
    a = np.arange(3 * (1 << 19), dtype=np.float32).reshape(1 << 19, 3)
    cols = np.asarray([2, 0])
    out = _take_columns(a, cols)
    assert out.flags.c_contiguous and np.array_equal(out, a[:, cols])


def test_shape_mismatch_is_rejected():
    scr = make()
    with pytest.raises(AssertionError):
        MiniScreen(scr.X[:, :5], scr.guides, scr.samples)


def test_latent_prior_defaults_scalars_and_per_variant_tensors():
    p = LatentPrior(5, None, 0.01, "cpu", torch.float32)
    assert not p.mu_normal and p.scalars == {"mu_loc": 0.0, "mu_scale": 1.0, "sd_loc": 0.0, "sd_scale": 0.01} and not p.vectors
    p = LatentPrior(5, {"mu_scale": 2.0}, 0.01, "cpu", torch.float32)  # naming either mu key switches Laplace -> Normal
    assert p.mu_normal and p.scalars["mu_loc"] == 0.0 and p.scalars["mu_scale"] == 2.0
    t = torch.linspace(0.5, 1.5, 5, dtype=torch.float64).reshape(5, 1)  # (T, 1) tensors of `bean build-prior`
    p = LatentPrior(5, {"mu_loc": t, "mu_scale": t, "sd_scale": torch.tensor(0.05)}, 0.01, "cpu", torch.float32)
    assert p.mu_normal and set(p.vectors) == {"mu_loc", "mu_scale"} and p.vectors["mu_loc"].shape == (5,)
    assert p.vectors["mu_loc"].dtype == torch.float32 and p.vectors["mu_loc"].is_contiguous()
    assert p.scalars["sd_scale"] == pytest.approx(0.05)


def test_get_guide_edit_rate_follows_the_reporter_screen_definition():
    """ReporterScreen.get_guide_edit_rate (bean/framework/ReporterScreen.py:448-529), hand-computed."""
    X = np.full((3, 4), 100.0, dtype=np.float32)
    bc = np.asarray([[10, 20, 30, 40], [0, 0, 5, 5], [8, 8, 0, 0]], dtype=np.float32)
    ed = np.asarray([[1, 2, 3, 4], [0, 0, 1, 1], [4, 4, 0, 0]], dtype=np.float32)
    guides = pd.DataFrame({"sequence": ["GGGAACAAGG", "GGGCCCCCGG", "AAAAAAAAAA"]}, index=["g0", "g1", "g2"])
    samples = pd.DataFrame({"condition": ["top", "bulk", "bot", "bulk_2"]}, index=["s0", "s1", "s2", "s3"])
    scr = MiniScreen(X, guides, samples, {"X_bcmatch": bc, "edits": ed}, {"tiling": False, "target_base_changes": "A>G,C>T"})
    assert scr.tiling is False and scr.target_base_changes == {"A": "G", "C": "T"}
    scr.get_guide_edit_rate(unsorted_condition_label="bulk")  # samples s1 and s3 (label contained in the condition)
    assert np.allclose(scr.guides["edit_rate"], [(2 + 4 + 0.5) / (20 + 40 + 0.5), (0 + 1 + 0.5) / (0 + 5 + 0.5), (4 + 0 + 0.5) / (8 + 0 + 0.5)])
    assert "edit_rate_norm" not in scr.guides.columns
    all_samples = scr.get_guide_edit_rate(return_result=True, bcmatch_thres=17)
    assert np.isnan(all_samples[1]) and np.isnan(all_samples[2]) and np.isclose(all_samples[0], 10.5 / 100.5)
    til = MiniScreen(X, guides.copy(), samples, {"X_bcmatch": bc, "edits": ed}, {"tiling": True, "target_base_change": "A>G"})
    til.get_guide_edit_rate(unsorted_condition_label="bulk")
    # editable A's in positions 3..7 of the protospacer: "AACAA" -> 4, "CCCCC" -> 0 (NaN), "AAAAA" -> 5
    assert np.allclose(til.guides["edit_rate_norm"].to_numpy()[[0, 2]], [6.5 / (60 * 4 + 0.5), 4.5 / (8 * 5 + 0.5)]) and np.isnan(til.guides["edit_rate_norm"].iloc[1])
    with pytest.raises(ValueError, match="is not found"):
        scr.get_guide_edit_rate(unsorted_condition_label="plasmid")
    with pytest.raises(ValueError, match="not available"):
        MiniScreen(X, guides, samples, {}, {"tiling": False}).get_guide_edit_rate()


def test_edit_rate_equals_the_column_stored_in_the_reference_fixture():
    """The reference's survival_var_mini_screen.h5ad (travelling inside the golden fixture) carries the `edit_rate` column its
    own pipeline wrote (ReporterScreen.get_guide_edit_rate with the control condition D7, tests/test_run.py:166): the
    restatement reproduces it exactly."""
    import os

    from tests.helpers import GOLDEN
    from tests.refharness.golden import screen_from_arrays

    scr = screen_from_arrays(np.load(os.path.join(GOLDEN, "ref_survival_real_var_mixture.npz")))
    stored = scr.guides["edit_rate"].to_numpy().copy()
    scr.uns["tiling"] = False
    got = scr.get_guide_edit_rate(return_result=True, unsorted_condition_label="D7")
    assert np.allclose(got, stored, rtol=1e-12, atol=0, equal_nan=True) and np.isfinite(stored).sum() >= 20
