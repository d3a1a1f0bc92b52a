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
