"""`prepare_bdata`: what `bean run` does to the screen between reading it and tensorising it.

Host-side mirror of bean/preprocessing/utils.py:24-67 (+ bean/qc/guide_qc.py:49-75 `filter_no_info_target`), called at
bean/cli/run.py:97.  Same argument object (`args.replicate_col, selection, exclude_control_condition_for_inference,
condition_col, control_condition, library_design, target_col`), same warnings, same side file
`{prefix}/no_support_targets.csv`, same guide order (variant designs are sorted by target with the reference's own
`Series.argsort()` call, so ties fall the same way).

One deliberate difference: the reference tests `bdata.samples.mask == 1`, where `samples.mask` is pandas' DataFrame.mask
METHOD, not the column; the comparison is a scalar False and AnnData then looks at sample 0 only.  Here every sample whose
`mask` column is 1 is checked, which is what the error message says.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def filter_no_info_target(bdata, condit_col: str, control_condition: str, target_col: str = "target",
                          write_no_support_targets: bool = False, no_support_target_write_path: str = None):
    """Drop the guides of targets whose guides have no counts in ANY sample (guide_qc.py:49-75; `condit_col` and
    `control_condition` are accepted and unused there too).  Returns (number of dropped targets, screen)."""
    totals = pd.Series(np.asarray(bdata.X).sum(axis=1), index=bdata.guides.index).groupby(bdata.guides[target_col]).sum()
    zero = totals.index[(totals == 0).to_numpy()]
    if write_no_support_targets:
        zero.to_series().to_csv(no_support_target_write_path, index=False)
    bdata = bdata[(~bdata.guides[target_col].isin(zero)).to_numpy(), :].copy()
    return len(zero), bdata


def prepare_bdata(bdata, args, warn, prefix: str):
    bdata = bdata.copy()
    bdata.samples["replicate"] = bdata.samples[args.replicate_col].astype("category")
    bdata.guides = bdata.guides.loc[:, ~bdata.guides.columns.duplicated()].copy()
    # guides without a single count in the samples that inform the effect size
    if args.selection == "sorting" or args.exclude_control_condition_for_inference:
        tested = bdata[:, (bdata.samples[args.condition_col] != args.control_condition).to_numpy()]
    else:
        tested = bdata
    empty = np.asarray(tested.X).sum(axis=1) == 0
    if empty.any():
        warn(f"Filtering out {int(empty.sum())} gRNAs without any counts over all samples.")
        bdata = bdata[~empty, :]
    if "mask" in bdata.samples.columns:
        used = (bdata.samples["mask"] == 1).to_numpy()
        dead = used & (np.asarray(bdata.X).sum(axis=0) == 0)
        if dead.any():
            raise ValueError(f"Sample {bdata.samples.index[dead]} has 0 counts. Make sure you mask that sample.")
    if args.library_design != "variant":
        return bdata
    if bdata.guides[args.target_col].isnull().any():
        raise ValueError(f"Some target column (bdata.guides[{args.target_col}]) value is null. Check your input file.")
    bdata = bdata[bdata.guides[args.target_col].argsort(), :]
    n_dropped, bdata = filter_no_info_target(bdata, condit_col=args.condition_col, control_condition=args.control_condition,
                                             target_col=args.target_col, write_no_support_targets=True,
                                             no_support_target_write_path=f"{prefix}/no_support_targets.csv")
    if n_dropped > 0:
        warn(f"Ignoring {n_dropped} targets with 0 gRNA counts across all non-control samples. Ignored targets are written in "
             f"{prefix}/no_support_targets.csv.")
    return bdata
