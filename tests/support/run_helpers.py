"""Caller-side helpers of bean/cli/run.py restated for the parity harness (argument checks, per-target table, --prior-params):
`check_args`, `_get_guide_target_info`, `_get_guide_to_variant_df`, `_check_prior_params` (bean/model/run.py:39-344, :479-542).
Out of scope of the hot path: test support only."""
from __future__ import annotations

from logging import info  # noqa: F401

import torch  # noqa: F401

# ---- caller-side helpers of bean/cli/run.py (argument checks, per-target table, --prior-params) ------------------------
def _require(cond, message):
    if not cond:
        raise ValueError(message)


def check_args(args, bdata, warn=None):
    """bean/model/run.py:39-188: validate the `bean run` arguments against the screen and fill in the derived ones
    (`adjust_confidence_by_negative_control`, `allele_df_key`, `popt`, `outdir`, the all-ones replicate x guide mask).
    Same conditions, same exception type (ValueError) and same side effects; messages name the same option."""
    import logging
    import os

    import pandas as pd

    warn = warn or logging.warning
    guides, samples = bdata.guides, bdata.samples
    if args.scale_by_acc:
        _require(args.acc_col is not None or args.acc_bw_path is not None,
                 "--scale-by-acc not accompanied by --acc-col nor --acc-bw-path to use. Pass either one.")
        if args.acc_col is not None and args.acc_bw_path is not None:
            warn("Both --acc-col and --acc-bw-path is specified. --acc-bw-path is ignored.")
            args.acc_bw_path = None
        elif args.acc_bw_path is not None and "genomic_pos" not in guides.columns:
            _require("start_pos" in guides.columns, "Guides' positions not provided in ReporterScreen.guides['start_pos']. Please check the input.")
            guides["genomic_pos"] = guides["start_pos"]
            warn("'genomic_pos' not in ReporterScreen.guides.columns, using 'start_pos' to retrieve accessibility from the bigWig file.")
    if args.outdir is None:
        args.outdir = os.path.dirname(args.bdata_path)
    _require(not args.fit_negctrl or args.negctrl_col in guides.columns,
             f"--negctrl-col argument '{args.negctrl_col}' not in ReporterScreen.guides.columns {guides.columns}.")
    if args.selection == "sorting":
        for opt, col in (("--sorting-bin-upper-quantile-col", args.sorting_bin_upper_quantile_col),
                         ("--sorting-bin-lower-quantile-col", args.sorting_bin_lower_quantile_col)):
            _require(col in samples.columns, f"{opt} argument '{col}' not in ReporterScreen.samples.columns {samples.columns}.")
    elif args.selection == "survival":
        _require(args.time_col in samples.columns, f"--time-col argument '{args.time_col}' not in ReporterScreen.samples.columns {samples.columns}.")
        try:
            pd.to_numeric(samples[args.time_col])
        except ValueError as exc:
            raise ValueError(f"ReporterScreen.samples['{args.time_col}'] provided is not numeric ({samples[args.time_col]}).") from exc
    if args.library_design == "variant":
        args.adjust_confidence_by_negative_control = args.fit_negctrl and (not args.dont_adjust_confidence_by_negative_control)
    elif args.library_design == "tiling":
        args.adjust_confidence_by_negative_control = not args.dont_adjust_confidence_by_negative_control
        if args.allele_df_key is None:
            # the most filtered allele table: fewest rows among the uns tables whose key contains "allele_counts"
            tables = {k: v for k, v in bdata.uns.items() if "allele_counts" in k and isinstance(v, pd.DataFrame)}
            best, n_best = "allele_counts", len(bdata.uns["allele_counts"])
            for key, tbl in tables.items():
                if len(tbl) < n_best:
                    best, n_best = key, len(tbl)
            warn(f"--allele-df-key not provided for tiling screen. Using the most filtered allele counts with {n_best} alleles stored in '{best}'.")
            args.allele_df_key = best
        else:
            _require(args.allele_df_key in bdata.uns, f"--allele-df-key '{args.allele_df_key}' not in ReporterScreen.uns. Check your input.")
    else:
        raise ValueError("Invalid library_design provided. Select either 'variant' or 'tiling'.")
    if args.fit_negctrl:
        n_negctrl = int((guides[args.negctrl_col].map(lambda s: s.lower()) == args.negctrl_col_value.lower()).sum())
        _require(n_negctrl >= 10, f"Not enough negative control guide in the input data: {n_negctrl}. Please check your input arguments.")
    if args.repguide_mask is not None and args.repguide_mask not in bdata.uns.keys():
        bdata.uns[args.repguide_mask] = pd.DataFrame(1, index=guides.index, columns=samples[args.replicate_col].unique())
        warn(f"{args.bdata_path} does not have replicate x guide outlier mask. All guides are included in analysis.")
    if args.sample_mask_col == "":
        args.sample_mask_col = None
    _require(args.sample_mask_col is None or args.sample_mask_col in samples.columns.tolist(),
             f"{args.bdata_path} does not have specified sample mask column `{args.sample_mask_col}` in .samples")
    _require(args.condition_col in samples.columns,
             f"Condition column `{args.condition_col}` set by `--condition-col` not in ReporterScreen.samples.columns:{samples.columns}.")
    _require(not (args.selection == "survival" and args.condition_col == args.time_col),
             f"Invalid to have the same `--condition-col` ({args.condition_col}) and `--time-col` ({args.time_col}).")
    present = samples[args.condition_col].astype(str).tolist()
    for label in args.control_condition.split(","):
        _require(label in present, f"No sample has control label `{args.control_condition}` (set by `--control-condition`) in "
                                   f"ReporterScreen.samples[{args.condition_col}].")
    _require(args.replicate_col in samples.columns,
             f"Condition column set by `--replicate-col` {args.replicate_col} not in ReporterScreen.samples.columns:{samples.columns}.")
    if args.control_guide_tag is not None:
        _require(args.library_design != "variant", "`--control-guide-tag` is not used for the variant mode.")
        _require(guides.index.map(lambda s: args.control_guide_tag in s).any(),
                 f"Negative control guide label `{args.control_guide_tag}` provided by `--control-guide-tag` doesn't appear in any of the guide names.")
    if args.alpha_if_overdispersion_fitting_fails is not None:
        try:
            b0, b1 = args.alpha_if_overdispersion_fitting_fails.split(",")
            args.popt = (float(b0), float(b1))
        except TypeError:
            raise ValueError(f"Input --alpha-if-overdispersion-fitting-fails `{args.alpha_if_overdispersion_fitting_fails}` is malformatted!")
    else:
        args.popt = None
    return args, bdata


def _get_guide_target_info(bdata, args, cols_include=()):
    """bean/model/run.py:191-225: one row per target -- the `target_*` guide columns that are constant within a target
    (plus `cols_include`), the number of guides, and mean / std of the guides' `edit_rate` when present."""
    g = bdata.guides.copy()
    tcol = args.target_col
    n_targets = len(g[tcol].unique())
    keep = [c for c in g.columns if c != tcol and (c in cols_include or (c.startswith("target_") and len(g[[tcol, c]].drop_duplicates()) == n_targets))]
    info = g[[tcol] + keep].drop_duplicates().set_index(tcol, drop=True)
    info["n_guides"] = g.groupby("target").size()  # (the reference groups by the literal "target" here)
    if "edit_rate" in g.columns.tolist():
        rate = g[[tcol, "edit_rate"]].groupby(tcol, sort=False)["edit_rate"].agg(["mean", "std"])
        info = info.join(rate.rename(columns={"mean": "edit_rate_mean", "std": "edit_rate_std"}))
    return info


def _get_guide_to_variant_df(target_info_df):
    """bean/model/run.py:314-344 (tiling): per guide, the variants it generated and its editing rate for each, from the
    comma-separated `editing_guides` / `per_guide_editing_rates` columns of the per-variant table."""
    import numpy as np
    import pandas as pd

    rows = []
    for variant, guides, rates in zip(target_info_df["edit"], target_info_df["editing_guides"], target_info_df["per_guide_editing_rates"]):
        if guides and pd.isnull(guides):
            continue
        names = guides.strip(",").split(",")
        vals = [(float(x) if x else np.nan) for x in rates.strip(",").split(",")]
        rows += [(n, variant, r) for n, r in zip(names, vals)]  # zip: stops at the shorter list, as the reference's zip does
    df = pd.DataFrame(rows, columns=["guide", "variants", "per_variant_edit_rate"])
    return df.groupby("guide").agg(list)


def _check_prior_params(param_path: str, ndata):
    """bean/model/run.py:479-542: load the `--prior-params` pickle (`bean build-prior`) and bring per-variant arrays to the
    (n_targets, 1) shape the models broadcast against.  The reference compares some shapes with the INT `n_targets`
    instead of a tuple, so 1-D `sd_scale` / `mu_scale` (and, for survival screens, `mu_loc`) are rejected rather than
    reshaped; that behaviour is kept."""
    import os
    import pickle

    if not os.path.exists(param_path):
        raise ValueError(f"Specified prior parameter file --prior-params {param_path} is not found.")
    with open(param_path, "rb") as f:
        prior = pickle.load(f)
    T = ndata.n_targets

    def fix(key, reshape_1d, only_arrays):
        if key not in prior or (only_arrays and not hasattr(prior[key], "__len__")):
            return
        if reshape_1d and prior[key].shape == (T,):
            prior[key] = prior[key].reshape(-1, 1)
        elif prior[key].shape != (T, 1):
            raise ValueError(f"Specified prior parameter --prior-params {param_path}: prior_params['{key}'].shape {prior[key].shape} "
                             f"does not match the number of target variants {(T, 1)}.")

    if getattr(ndata, "is_sorting", False):
        fix("sd_loc", True, False)
        fix("sd_scale", False, False)
        fix("mu_loc", True, True)
        fix("mu_scale", False, True)
    else:
        if "initial_abundance" in prior and prior["initial_abundance"].shape != (T,):
            raise ValueError(f"Specified prior parameter --prior-params {param_path}: prior_params['initial_abundance'].shape does not "
                             f"match the number of guides {(ndata.n_guides, 1)}.")
        fix("mu_loc", False, True)
        fix("mu_scale", False, True)
    return prior
