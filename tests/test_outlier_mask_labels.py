"""`uns[repguide_mask]` is applied by guide / replicate LABEL (the reference asserts the order, data_class.py:167-178)."""
import numpy as np
import pandas as pd
import pytest
import torch

from crispr_bean_b200.data_class import VariantSortingReporterScreenData
from crispr_bean_b200.synth import make_sorting_screen


def _screen():
    scr = make_sorting_screen(6, 3, n_reps=3, seed=4, depth=200.0)
    reps = sorted(scr.samples["replicate"].unique())
    rng = np.random.default_rng(0)
    tbl = pd.DataFrame((rng.random((scr.X.shape[0], len(reps))) > 0.3).astype(int), index=scr.guides.index, columns=reps)
    return scr, tbl


def test_permuted_table_gives_the_same_mask():
    scr, tbl = _screen()
    scr.uns["repguide_mask"] = tbl
    ref = VariantSortingReporterScreenData(scr, control_can_be_selected=True, repguide_mask="repguide_mask").repguide_mask
    scr2, _ = _screen()
    scr2.uns["repguide_mask"] = tbl.iloc[::-1, ::-1]  # rows and columns in another order
    got = VariantSortingReporterScreenData(scr2, control_can_be_selected=True, repguide_mask="repguide_mask").repguide_mask
    assert torch.equal(ref, got)


def test_missing_label_raises():
    scr, tbl = _screen()
    scr.uns["repguide_mask"] = tbl.rename(columns={tbl.columns[0]: "other"})
    with pytest.raises(ValueError):
        VariantSortingReporterScreenData(scr, control_can_be_selected=True, repguide_mask="repguide_mask")
