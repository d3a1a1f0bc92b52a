"""CSR versions of the per-variant summaries of the tiling result table (tests/support/tiling_stats.py) against the
reference's dense-tensor functions (bean/preprocessing/utils.py:254-310) run on the reference's own data class."""
import ast
import os

import numpy as np
import pytest
import torch

from crispr_bean_b200 import data_class as dc
from tests.support import tiling_stats as ts
from crispr_bean_b200.synth import make_tiling_screen
from tests.helpers import GOLDEN
from tests.refharness import available, load_reference
from tests.refharness.golden import screen_from_arrays

needs_reference = pytest.mark.skipif(not available(), reason="reference sources not mounted")


def cases():
    yield "synthetic", make_tiling_screen(n_guides=40, n_reps=3, max_alleles=6, seed=2), dict(control_can_be_selected=True, allele_df_key="allele_counts")
    z = np.load(os.path.join(GOLDEN, "ref_tiling_real_mini.npz"))  # the reference's tiling_mini_screen.h5ad
    yield "tiling_mini_screen", screen_from_arrays(z), ast.literal_eval(str(z["meta/data_kwargs"]))


@needs_reference
@pytest.mark.parametrize("count_thres", [0, 10])
def test_summaries_equal_reference(count_thres):
    import sys

    ns = load_reference()
    sys.path.insert(0, GOLDEN)
    from make_reference_golden import with_allele_objects

    for name, scr, kw in cases():
        ref_data = ns.data_class.TilingSortingReporterScreenData(with_allele_objects(ns, scr), **kw)
        data = dc.TilingSortingReporterScreenData(scr, **kw)
        keys = sorted(ref_data.edit_index, key=ref_data.edit_index.get)
        perm = np.asarray([data.edit_index[str(k)] for k in keys])  # our position of the reference's edit j
        r_idx, r_rates, r_tot = ns.prep_utils._obtain_effective_edit_rate(ref_data, count_thres=count_thres)
        m_idx, m_rates, m_tot = ts._obtain_effective_edit_rate(data, count_thres=count_thres)
        assert len(m_idx) == len(r_idx) == data.n_edits
        assert np.allclose(m_tot.numpy()[perm], r_tot.numpy(), rtol=1e-5, atol=1e-7), name
        for j, e in enumerate(perm):
            assert torch.equal(m_idx[e].reshape(-1), r_idx[j].reshape(-1)), (name, j)
            assert np.allclose(m_rates[e], r_rates[j], rtol=1e-5, atol=1e-8), (name, j)
        assert np.array_equal(ts._obtain_n_guides_alleles_per_variant(data).numpy()[perm], ns.prep_utils._obtain_n_guides_alleles_per_variant(ref_data).numpy())
        assert np.array_equal(ts._obtain_n_cooccurring_variants(data)[perm], ns.prep_utils._obtain_n_cooccurring_variants(ref_data))


def test_against_the_dense_definition():
    """Portable: the same quantities from the dense (G, A-1, E) tensor the mirror can still materialise."""
    data = dc.TilingSortingReporterScreenData(make_tiling_screen(n_guides=30, n_reps=2, max_alleles=5, seed=8),
                                              control_can_be_selected=True, allele_df_key="allele_counts")
    dense = data.allele_to_edit
    assert torch.equal(ts._obtain_n_guides_alleles_per_variant(data), (dense.sum(axis=1) > 0).sum(axis=0))
    n_co = ts._obtain_n_cooccurring_variants(data)
    for e in range(data.n_edits):
        g, a = torch.where(dense[:, :, e] > 0)
        assert n_co[e] == int((dense[g, a, :].sum(axis=0) > 0).sum()) - 1
    _, per_guide, total = ts._obtain_effective_edit_rate(data, count_thres=0)
    assert (total >= 0).all() and all(abs(sum(p) - float(t)) <= 1e-5 * max(float(t), 1) for p, t in zip(per_guide, total))
