"""Tensorisation of a screen for the SVI hot path (host-side mirror of bean/preprocessing/data_class.py).

Same class names, constructor keywords and tensor attributes as the reference
(`X, X_masked, X_bcmatch(_masked), size_factor(_bcmatch), sample_mask, repguide_mask, a0, a0_bcmatch,
pi_a0, allele_counts_control, upper_bounds, lower_bounds, target_lengths, n_targets, ...`; reference
layout `(R, B, G)`, data_class.py:124-205, 312-397, 493-532, 913-1000), PLUS what the B200 kernels
need (SURVEY section 7 stage 3): the flat CSR guide->variant arrays `variant_ptr (T+1)`,
`guide_variant (G)` and -- via `device_pack.pack_*` -- guide-major `(G, R, B)` device records.

Written vectorised (no per-guide Python loops, unlike data_class.py:511-532) so that the 1M-guide
configuration tensorises in seconds.
"""
from __future__ import annotations

from copy import copy
from typing import Optional, Sequence, Tuple

import numpy as np
import pandas as pd
import torch

from .alpha0 import (get_fitted_alpha0, get_fitted_pi_alpha0, get_pred_alpha0, get_pred_pi_alpha0,
                     get_size_factor)


def _as_list(control_condition):
    return control_condition.split(",") if isinstance(control_condition, str) else list(control_condition)


ACCESSIBILITY_LOOKUP = None  # callable(bw_path, guides) -> tensor: reader of --acc-bw-path tracks, supplied by the caller


class ScreenData:
    """Base tensoriser (reference: `ScreenData`, data_class.py:35-263)."""

    is_reporter = False
    is_sorting = False
    is_survival = False

    def __init__(self, screen, repguide_mask: str = None, sample_mask_column: str = None,
                 shrink_alpha: bool = False, condition_column: str = "condition",
                 control_condition: str = "bulk", accessibility_col: str = None,
                 accessibility_bw_path: str = None, device: str = None,
                 replicate_column: str = "replicate", popt: Optional[Tuple[float, float]] = None,
                 pi_popt: Optional[Tuple[float, float]] = None, control_can_be_selected=False,
                 negctrl_guide_idx: Optional[Sequence[int]] = None, target_col: str = "target",
                 lower_quantile_column: str = "lower_quantile", upper_quantile_column: str = "upper_quantile",
                 time_column: str = "time", use_bcmatch: bool = False, impute_pi_popt: bool = False,
                 **kwargs):
        self.device = device
        self.condition_column = condition_column
        self.replicate_column = replicate_column
        self.control_condition = _as_list(control_condition)
        self.control_can_be_selected = bool(control_can_be_selected)
        self.sample_mask_column = sample_mask_column
        self.repguide_mask_key = repguide_mask
        self.shrink_alpha = shrink_alpha
        self.popt = popt
        self.pi_popt = pi_popt
        self.negctrl_guide_idx = negctrl_guide_idx
        self.accessibility_col = accessibility_col
        self.accessibility_bw_path = accessibility_bw_path
        self.target_col = target_col
        self._lq_col, self._uq_col, self.time_column = lower_quantile_column, upper_quantile_column, time_column

        screen = screen.copy()
        smp = screen.samples
        if not (replicate_column in smp.columns and condition_column in smp.columns):
            # data_class.py:64-70: sample names "<replicate>_<condition>" supply both columns when either is absent
            parts = [str(s).rsplit("_", 1) for s in smp.index]
            if any(len(p) != 2 for p in parts):
                raise ValueError(f"screen.samples lacks '{replicate_column}' / '{condition_column}' and the sample names "
                                 "are not of the form <replicate>_<condition>.")
            smp[replicate_column] = [p[0] for p in parts]
            smp[condition_column] = [p[1] for p in parts]
        covs = screen.uns.get("sample_covariates")
        if covs is not None:
            # sample covariates (data_class.py:75-92): a "replicate" becomes a (replicate, covariates...) combination
            self.sample_covariates = list(covs)
            self.n_sample_covariates = len(self.sample_covariates)
            smp["_rc"] = [".".join(map(str, row)) for row in smp[[replicate_column] + self.sample_covariates].values.tolist()]
            replicate_column = self.replicate_column = "_rc"
        smp["size_factor"] = get_size_factor(screen.X)  # all samples incl. control (data_class.py:63)
        if "X_bcmatch" in screen.layers:
            smp["size_factor_bcmatch"] = get_size_factor(screen.layers["X_bcmatch"])
        cond = smp[condition_column]
        is_control = cond.astype(str).isin(self.control_condition).to_numpy()
        # the CLI always passes a truthy `control_can_be_selected` (~bool), so the control
        # sample joins the selected set as a pseudo-bin; False gives the documented intent
        # (--exclude-control-condition-for-inference) instead of the reference's crash.
        selected = ~cond.isnull().to_numpy() if self.control_can_be_selected else ~is_control
        reps = sorted(smp[replicate_column].unique())
        smp[f"{replicate_column}_id"] = smp[replicate_column].map({r: i for i, r in enumerate(reps)})
        self.n_reps = len(reps)
        self._assign_condition_ids(smp, selected)
        order = np.lexsort((smp[f"{condition_column}_id"].to_numpy(), smp[f"{replicate_column}_id"].to_numpy()))
        screen = screen[:, order]
        selected, is_control = selected[order], is_control[order]
        self.screen = screen
        if covs is not None:  # 0/1 design of the covariates per (sorted) replicate combination (data_class.py:973-980)
            self.rep_by_cov = torch.as_tensor(screen.samples[["_rc"] + self.sample_covariates].drop_duplicates()
                                              .set_index("_rc").values.astype(int))
        self.screen_selected = screen[:, selected]
        self.screen_control = screen[:, is_control]
        self.n_condits = len(self.screen_selected.samples[condition_column].unique())
        self.n_samples = len(screen.samples)
        self.n_guides = len(screen.guides)
        self._post_init()
        if self.target_col is not None and self.target_col in screen.guides.columns:
            self._variant_init()
        if self.is_reporter or (use_bcmatch and "X_bcmatch" in screen.layers):
            self._bcmatch_init()
        if self.is_reporter:
            self._reporter_init(impute_pi_popt)

    # -- condition ids ------------------------------------------------------------------------
    def _assign_condition_ids(self, smp: pd.DataFrame, selected: np.ndarray):
        """Sorting: bins ordered by (upper, lower) quantile (data_class.py:948-964)."""
        raise NotImplementedError

    # -- shared tensors ------------------------------------------------------------------------
    def transform_data(self, X, n_bins=None):
        n_bins = self.n_condits if n_bins is None else n_bins
        return torch.as_tensor(np.array(X)).T.reshape((self.n_reps, n_bins, self.n_guides)).float()

    def _post_init(self):
        R, B, C = self.n_reps, self.n_condits, len(self.control_condition)
        sel, ctl = self.screen_selected, self.screen_control
        if self.accessibility_col is not None:
            self.guide_accessibility = torch.as_tensor(self.screen.guides[self.accessibility_col].to_numpy().copy())
        elif self.accessibility_bw_path is not None:  # data_class.py:130-133
            # the bigWig lookup of --acc-bw-path is pre-processing, out of scope of the hot path (pyBigWig in the reference):
            # the caller plugs a reader in (tests: tests/support/accessibility.py)
            if ACCESSIBILITY_LOOKUP is None:
                raise NotImplementedError("accessibility_bw_path needs a bigWig reader: set data_class.ACCESSIBILITY_LOOKUP = "
                                          "callable(bw_path, guides_dataframe) -> per-guide accessibility, or pass accessibility_col")
            self.guide_accessibility = ACCESSIBILITY_LOOKUP(self.accessibility_bw_path, self.screen.guides)
        else:
            self.guide_accessibility = None
        if self.sample_mask_column is not None and self.sample_mask_column in sel.samples.columns:
            self.sample_mask = torch.as_tensor(sel.samples[self.sample_mask_column].to_numpy().copy()).reshape(R, B)
            self.control_sample_mask = torch.as_tensor(ctl.samples[self.sample_mask_column].to_numpy().copy()).reshape(R, C)
        else:
            self.sample_mask = torch.ones((R, B), dtype=torch.bool)
            self.control_sample_mask = torch.ones((R, C), dtype=torch.bool)
        self.X = self.transform_data(sel.X)
        self.X_masked = self.X * self.sample_mask[:, :, None]
        self.X_control = self.transform_data(ctl.X, C)
        self.X_control_masked = self.X_control * self.control_sample_mask[:, :, None]
        self.repguide_mask = ~(self.X == 0).any(axis=1)
        if self.repguide_mask_key is not None:
            # guides x replicates table, applied by LABEL: the reference asserts that its index equals guides.index and its
            # columns equal the (sorted) replicate order (data_class.py:167-178) and applies it by position
            tbl = self.screen.uns[self.repguide_mask_key]
            assert tbl.shape == (self.n_guides, R), tbl.shape
            reps = list(pd.unique(sel.samples[self.replicate_column]))
            if not (tbl.index.equals(self.screen.guides.index) and list(tbl.columns) == reps):
                missing = [g for g in self.screen.guides.index if g not in tbl.index][:3] + [r for r in reps if r not in tbl.columns][:3]
                if missing:
                    raise ValueError(f"screen.uns[{self.repguide_mask_key!r}] lacks guides / replicates {missing}")
                tbl = tbl.reindex(index=self.screen.guides.index, columns=reps)
            self.repguide_mask = torch.logical_and(torch.as_tensor(tbl.to_numpy().T) > 0, self.repguide_mask)
        self.size_factor = torch.as_tensor(sel.samples["size_factor"].to_numpy().copy()).reshape(R, B)
        self.size_factor_control = torch.as_tensor(ctl.samples["size_factor"].to_numpy().copy()).reshape(R, C)
        self.a0, self.popt = get_fitted_alpha0(self.X.clone(), self.size_factor.clone(), self.sample_mask,
                                               shrink=self.shrink_alpha, popt=self.popt)

    def _variant_init(self):
        """n_targets / target_lengths (data_class.py:493-532) + CSR guide->variant arrays."""
        codes = pd.Categorical(self.screen.guides[self.target_col]).codes
        change = np.flatnonzero(np.diff(codes) != 0) + 1
        starts = np.concatenate([[0], change])
        n_unique = len(np.unique(codes))
        if len(starts) != n_unique:
            raise ValueError(
                "Input Screen object not sorted for target identity. Sort the screen object so that guides "
                f"targeting the same object would occur as consecutive block by screen[screen.guides[{self.target_col}].argsort(),:]")
        self.n_targets = n_unique
        ptr = np.concatenate([starts, [len(codes)]]).astype(np.int64)
        self.target_lengths = torch.as_tensor(np.diff(ptr))
        self.n_sgRNAs_per_target = int(self.target_lengths.max())
        self.variant_ptr = torch.as_tensor(ptr.astype(np.int32))
        self.guide_variant = torch.repeat_interleave(
            torch.arange(self.n_targets, dtype=torch.int32), self.target_lengths)

    def _bcmatch_init(self):
        """Barcode-matched count layer (data_class.py:320-345, :1190-1222)."""
        R, B, C = self.n_reps, self.n_condits, len(self.control_condition)
        sel, ctl = self.screen_selected, self.screen_control
        self.X_bcmatch = self.transform_data(sel.layers["X_bcmatch"])
        self.X_bcmatch_masked = self.X_bcmatch * self.sample_mask[:, :, None]
        self.X_bcmatch_control = self.transform_data(ctl.layers["X_bcmatch"], C)
        self.X_bcmatch_control_masked = self.X_bcmatch_control * self.control_sample_mask[:, :, None]
        self.size_factor_bcmatch = torch.as_tensor(sel.samples["size_factor_bcmatch"].to_numpy().copy()).reshape(R, B)
        self.size_factor_bcmatch_control = torch.as_tensor(ctl.samples["size_factor_bcmatch"].to_numpy().copy()).reshape(R, C)
        self.a0_bcmatch = get_pred_alpha0(self.X_bcmatch.clone(), self.size_factor_bcmatch.clone(), self.popt,
                                          self.sample_mask)

    def _reporter_init(self, impute_pi_popt=False):
        """Reporter allele counts of the control condition and `pi_a0` (data_class.py:365-397, :416-453)."""
        C = len(self.control_condition)
        edited = self.transform_data(self.screen_control.layers["edits"], C)
        nonedited = (self.X_bcmatch_control - edited).clamp(min=0)
        self.allele_counts_control = torch.stack([nonedited, edited], axis=-1)  # (R, C, G, 2)
        pi_popt = self.popt if impute_pi_popt else self.pi_popt
        if pi_popt is not None:
            self.pi_a0 = get_pred_pi_alpha0(self.allele_counts_control.clone(), self.size_factor_control.clone(), pi_popt)
        else:
            self.pi_a0, self._pi_popt = get_fitted_pi_alpha0(self.allele_counts_control.clone(),
                                                             self.size_factor_control.clone(), shrink=self.shrink_alpha)

    # -- guide subsetting (negative-control fit: cli/run.py:236-257) --------------------------------
    _GUIDE_AXIS = {"X": 2, "X_masked": 2, "X_control": 2, "X_control_masked": 2, "repguide_mask": 1, "a0": 0,
                   "X_bcmatch": 2, "X_bcmatch_masked": 2, "X_bcmatch_control": 2, "X_bcmatch_control_masked": 2,
                   "a0_bcmatch": 0, "pi_a0": 0, "allele_counts_control": 2, "guide_accessibility": 0,
                   "allele_counts": 2}

    def pin_memory(self):
        """Page-lock every tensor of the tensorised screen (in place) so the upload to the GPU runs at full PCIe
        rate and asynchronously; a no-op without CUDA."""
        if torch.cuda.is_available():
            for k, v in vars(self).items():
                if torch.is_tensor(v) and not v.is_cuda and not v.is_pinned() and v.numel() > 0:
                    setattr(self, k, v.contiguous().pin_memory())
        return self

    def __getitem__(self, guide_idx):
        idx = torch.as_tensor(np.asarray(guide_idx)).long()
        nd = copy(self)
        nd.screen = self.screen[idx.numpy(), :]
        nd.screen_selected = self.screen_selected[idx.numpy(), :]
        nd.screen_control = self.screen_control[idx.numpy(), :]
        nd.n_guides = len(idx)
        for name, axis in self._GUIDE_AXIS.items():
            v = getattr(self, name, None)
            if v is not None:
                setattr(nd, name, v.index_select(axis, idx))
        if hasattr(self, "target_lengths"):
            nd._variant_init()
        return nd


class SortingScreenData(ScreenData):
    """Sorting screens: quantile bins (reference: data_class.py:876-1000)."""

    is_sorting = True

    def _assign_condition_ids(self, smp, selected):
        lq, uq = smp[self._lq_col].to_numpy(dtype=np.float64), smp[self._uq_col].to_numpy(dtype=np.float64)
        if ((lq < 0) | (lq > 1)).any() or ((uq < 0) | (uq > 1)).any():
            raise ValueError("Invalid quantile value in screen.samples: check input.")
        if (uq - lq < 0).any():
            raise ValueError(f"Not all screen.samples[{self._uq_col}] larger than screen.samples[{self._lq_col}]: check input.")
        sizes = smp.groupby([self._uq_col, self._lq_col]).size()
        if not (sizes == self.n_reps).all():
            raise ValueError(
                "Not all replicate share same quantile bin definition. If you have missing bin data, add the sample "
                "and add 'mask' column in 'screen.samples' or run `bean-qc` that automatically handles this.")
        bins = np.unique(np.stack([uq[selected], lq[selected]], axis=1), axis=0)  # sorted by (uq, lq)
        ids = np.full(len(smp), -1, dtype=np.int64)
        for j, (u, l) in enumerate(bins):
            ids[selected & (uq == u) & (lq == l)] = j
        smp[f"{self.condition_column}_id"] = ids
        self.upper_bounds = torch.as_tensor(bins[:, 0].copy())
        self.lower_bounds = torch.as_tensor(bins[:, 1].copy())


class VariantSortingScreenData(SortingScreenData):
    """data_class.py:1165-1247: guide counts only (Normal / ControlNormal models)."""

    def __init__(self, screen, *args, condition_column="bin", sample_mask_column="mask", **kwargs):
        super().__init__(screen, *args, condition_column=condition_column,
                         sample_mask_column=sample_mask_column, **kwargs)


class VariantSortingReporterScreenData(SortingScreenData):
    """data_class.py:1251-1295: + barcode-matched counts and reporter edits (MixtureNormal models)."""

    is_reporter = True

    def __init__(self, screen, *args, condition_column="bin", sample_mask_column="mask", **kwargs):
        super().__init__(screen, *args, condition_column=condition_column,
                         sample_mask_column=sample_mask_column, **kwargs)


class SurvivalScreenData(ScreenData):
    """Proliferation screens: conditions are timepoints (reference: data_class.py:1003-1122).

    `time` is divided by its maximum (`timepoints` in [0, 1]); the control condition stays a selected condition
    (`control_can_be_selected=True`, the reference default for survival) and `control_timepoint` holds its time."""

    is_survival = True

    def _assign_condition_ids(self, smp, selected):
        tc = self.time_column
        try:
            t = smp[tc].astype(float)
        except ValueError as exc:
            raise ValueError(f"Invalid timepoint value({smp[tc]}) in screen.samples[{tc}]: check input.") from exc
        smp[tc] = t / t.max()
        if smp[tc].isnull().any():
            raise ValueError(f"NaN values in time points provided in input: {smp[tc]}")
        if not (smp.groupby(self.condition_column).size() == self.n_reps).all():
            raise ValueError(
                "Not all replicate share same timepoint definition. If you have missing bin data, add the sample and add "
                "'mask' column in 'screen.samples', or run `bean-qc` that automatically handles this.")
        times = np.sort(smp[tc].unique())
        smp[f"{tc}_id"] = smp[tc].map({v: j for j, v in enumerate(times)}).astype(np.int64)
        smp[f"{self.condition_column}_id"] = smp[f"{tc}_id"]  # samples are ordered by (replicate, time)
        is_control = smp[self.condition_column].astype(str).isin(self.control_condition).to_numpy()
        control_timepoint = smp.loc[is_control, tc].unique()
        if len(control_timepoint) != len(self.control_condition):
            raise ValueError(
                "All samples with --control-condition should have the same --time-col column in "
                "ReporterScreen.samples[time_col]. Check your input ReporterScreen object.")
        self.control_timepoint = torch.tensor(control_timepoint)
        self.timepoints = torch.as_tensor(np.sort(smp.loc[selected, tc].unique()))
        self.n_timepoints = len(self.timepoints)

    def _reporter_init(self, impute_pi_popt=False):
        """+ `allele_counts (R, n_timepoints, G, 2)`: reporter alleles at every timepoint (data_class.py:347-362)."""
        scr, tc = self.screen, self.time_column
        per_time = []
        for t in self.timepoints:
            st = scr[:, (scr.samples[tc] == t.item()).to_numpy()]
            edited = self.transform_data(st.layers["edits"], 1)
            nonedited = (self.transform_data(st.layers["X_bcmatch"], 1) - edited).clamp(min=0)
            per_time.append(torch.stack([nonedited, edited], axis=-1))
        self.allele_counts = torch.cat(per_time, axis=1)
        super()._reporter_init(impute_pi_popt)


class VariantSurvivalScreenData(SurvivalScreenData):
    """data_class.py:1358-1427: guide counts only (survival Normal / ControlNormal models)."""

    def __init__(self, screen, *args, condition_column="condition", time_column="time", control_can_be_selected=True,
                 sample_mask_column="mask", **kwargs):
        super().__init__(screen, *args, condition_column=condition_column, time_column=time_column,
                         control_can_be_selected=control_can_be_selected, sample_mask_column=sample_mask_column, **kwargs)


class VariantSurvivalReporterScreenData(SurvivalScreenData):
    """data_class.py:1430-1475: + barcode-matched counts and reporter edits (survival MixtureNormal)."""

    is_reporter = True

    def __init__(self, screen, *args, condition_column="condition", time_column="time", control_can_be_selected=True,
                 sample_mask_column="mask", **kwargs):
        super().__init__(screen, *args, condition_column=condition_column, time_column=time_column,
                         control_can_be_selected=control_can_be_selected, sample_mask_column=sample_mask_column, **kwargs)


_EDIT_RE = None
_COMPLEMENT = {"A": "T", "C": "G", "T": "A", "G": "C", "-": "-"}


def abs_edit_key(token: str, uid: Optional[str] = None) -> str:
    """Identity of an edit across guides: the reference's `Edit.get_abs_edit()` (framework/Edit.py:75-87) computed
    from the edit's string form `[chrom:]pos:rel_pos:strand:ref>alt` (optionally `uid!` in front): the sense-strand
    base change at the absolute position -- the same edit seen from two overlapping guides has different `rel_pos`
    but one key.  With a `uid` (`control_guide_tag`: edits of control guides are made unique to their guide,
    data_class.py:604-614) the key is guide-relative instead.  Tokens in any other format (e.g. amino-acid edits) are
    their own key."""
    global _EDIT_RE
    if _EDIT_RE is None:
        import re

        _EDIT_RE = re.compile(r"(?:(?P<uid>[\w*]+)!)?(?:(?P<chrom>(?:chr)?\w+|nan):)?(?P<pos>-?\d+):(?P<rel>-?\d+):(?P<strand>[+-]):"
                              r"(?P<ref>[A-Z*-])>(?P<alt>[A-Z*-])")
    m = _EDIT_RE.fullmatch(token)
    if m is None:
        return token
    ref, alt = m["ref"], m["alt"]
    if m["strand"] == "-":
        ref, alt = _COMPLEMENT.get(ref, ref), _COMPLEMENT.get(alt, alt)
    chrom = f"{m['chrom']}:" if m["chrom"] else ""
    uid = uid if uid is not None else m["uid"]
    if uid is not None:
        return f"{uid}!{chrom}{int(m['rel'])}:{ref}>{alt}"
    return f"{chrom}{int(m['pos'])}:{ref}>{alt}"


class _TilingReporterMixin:
    """data_class.py:536-872 + :1298-1355 / :1487-1537: tiling screens -- every guide has up to `n_max_alleles - 1` edited
    alleles, each a set of edits shared across guides (MultiMixtureNormal models).

    `screen.uns[allele_df_key]` is the filtered allele-count table `bean filter` writes: columns `guide`,
    `allele` | `aa_allele`, then one count column per sample.  An allele's edits are the comma-separated
    tokens of `str(allele)` (the reference's `Allele.__repr__`), identified across guides by `abs_edit_key`.  Emits the reference attributes
    (`n_edits, n_max_alleles, edit_index, allele_mask, allele_counts_control`, dense `allele_to_edit` on
    demand) plus the flat CSR `allele_ptr / allele_edit` the kernels use.  Vectorised: the reference's
    per-guide / per-allele Python loops (:696-698, :756-788, :868-871) are gone.
    """

    is_reporter = True
    is_tiling = True

    def _tiling_args(self, allele_df_key, allele_col, control_guide_tag, kwargs):
        if allele_df_key is None:
            raise ValueError("tiling screens need allele_df_key (a table in screen.uns)")
        self._allele_df_key, self._allele_col, self._control_guide_tag = allele_df_key, allele_col, control_guide_tag
        kwargs["target_col"] = None

    def _reporter_init(self, impute_pi_popt=False):
        self._tiling_init()
        pi_popt = self.popt if impute_pi_popt else self.pi_popt
        if pi_popt is not None:
            self.pi_a0 = get_pred_pi_alpha0(self.allele_counts_control.clone(), self.size_factor_control.clone(), pi_popt)
        else:
            self.pi_a0, self._pi_popt = get_fitted_pi_alpha0(self.allele_counts_control.clone(),
                                                             self.size_factor_control.clone(), shrink=self.shrink_alpha)

    def _tiling_init(self):
        df = self.screen.uns[self._allele_df_key].reset_index(drop=True)
        col = self._allele_col or ("aa_allele" if "aa_allele" in df.columns else "allele")
        G = self.n_guides
        gi = self.screen.guides.index.get_indexer(df["guide"])
        df = df.loc[gi >= 0].reset_index(drop=True)
        gi = gi[gi >= 0]
        aid = df.groupby("guide", sort=False).cumcount().to_numpy() + 1  # allele_id_for_guide, table order
        self.n_max_alleles = int(aid.max()) + 1 if len(aid) else 1
        A = self.n_max_alleles
        tag = self._control_guide_tag
        uids = [g if (tag is not None and tag in str(g)) else None for g in df["guide"]]
        if tag is not None and not any(u is not None for u in uids):
            raise AssertionError("uid not assinged.")  # the reference's check (data_class.py:611-613)
        tokens = [[abs_edit_key(t.strip(), u) for t in str(a).split(",") if t.strip()] for a, u in zip(df[col], uids)]
        self.edit_index = {}
        for ts in tokens:  # unique edits in order of first appearance (preprocessing/utils.py:149-173)
            for t in ts:
                self.edit_index.setdefault(t, len(self.edit_index))
        self.n_edits = len(self.edit_index)
        slot = gi.astype(np.int64) * (A - 1) + (aid - 1)
        order = np.argsort(slot, kind="stable")
        counts = np.zeros(G * (A - 1), dtype=np.int64)
        counts[slot] = [len(ts) for ts in tokens]
        self.allele_ptr = torch.as_tensor(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32))
        self.allele_edit = torch.as_tensor(np.array([self.edit_index[t] for r in order for t in tokens[r]], dtype=np.int32))
        n_valid = np.zeros(G, dtype=np.int64)
        np.maximum.at(n_valid, gi, aid)
        self.allele_mask = torch.as_tensor(np.arange(A)[None, :] <= n_valid[:, None])  # column 0 = WT always exists
        self.allele_counts_control = self._allele_tensor(self.screen_control, df, gi, aid, len(self.control_condition))
        self.allele_counts = self._allele_tensor(self.screen_selected, df, gi, aid, self.n_condits)

    def _allele_tensor(self, scr, df, gi, aid, n_cond):
        """(R, n_cond, G, A) allele counts; WT = barcode-matched total minus the edited alleles, floored at 0."""
        G, A = self.n_guides, self.n_max_alleles
        out = np.zeros((len(scr.samples), G, A), dtype=np.float32)
        for j, name in enumerate(scr.samples.index):
            out[j, gi, aid] = df[name].to_numpy(dtype=np.float32)
        bc = np.asarray(scr.layers["X_bcmatch"]).T.astype(np.float32)  # (S, G)
        out[:, :, 0] = np.clip(bc - out[:, :, 1:].sum(-1), 0, None)
        return torch.as_tensor(out).reshape(self.n_reps, n_cond, G, A)

    @property
    def allele_to_edit(self):
        """Dense (G, A-1, E) 0/1 tensor of the reference (data_class.py:656-699); tests / oracle only."""
        dense = torch.zeros((self.n_guides * (self.n_max_alleles - 1), self.n_edits))
        rows = torch.repeat_interleave(torch.arange(len(self.allele_ptr) - 1), self.allele_ptr[1:].long() - self.allele_ptr[:-1].long())
        dense[rows, self.allele_edit.long()] = 1
        return dense.reshape(self.n_guides, self.n_max_alleles - 1, self.n_edits)

    def __getitem__(self, guide_idx):
        """Guide subset (sharding over GPUs): the guide-axis tensors and the CSR rows of the chosen guides; the EDITS stay those
        of the whole screen (`n_edits`, `edit_index`): an edit's alleles sit in guides of several shards."""
        idx = np.asarray(torch.as_tensor(np.asarray(guide_idx)).long().numpy(), dtype=np.int64)
        nd = super().__getitem__(idx)
        a1 = self.n_max_alleles - 1
        ptr = self.allele_ptr.numpy().astype(np.int64)
        slots = (idx[:, None] * a1 + np.arange(a1)[None, :]).reshape(-1)
        lens = ptr[slots + 1] - ptr[slots]
        new_ptr = np.concatenate([[0], np.cumsum(lens)])
        take = np.repeat(ptr[slots] - new_ptr[:-1], lens) + np.arange(new_ptr[-1])
        nd.allele_ptr = torch.as_tensor(new_ptr.astype(np.int32))
        nd.allele_edit = self.allele_edit[torch.as_tensor(take)] if len(take) else self.allele_edit[:0]
        nd.allele_mask = self.allele_mask[torch.as_tensor(idx)]
        return nd


class TilingSortingReporterScreenData(_TilingReporterMixin, SortingScreenData):
    """Tiling sorting screens (data_class.py:1298-1355)."""

    def __init__(self, screen, *args, condition_column="bin", sample_mask_column="mask", allele_df_key=None,
                 allele_col=None, control_guide_tag=None, **kwargs):
        self._tiling_args(allele_df_key, allele_col, control_guide_tag, kwargs)
        super().__init__(screen, *args, condition_column=condition_column, sample_mask_column=sample_mask_column, **kwargs)


class TilingSurvivalReporterScreenData(_TilingReporterMixin, SurvivalScreenData):
    """Tiling proliferation screens (data_class.py:1487-1537)."""

    def __init__(self, screen, *args, condition_column="condition", time_column="time", control_can_be_selected=True,
                 sample_mask_column="mask", allele_df_key=None, allele_col=None, control_guide_tag=None, **kwargs):
        self._tiling_args(allele_df_key, allele_col, control_guide_tag, kwargs)
        super().__init__(screen, *args, condition_column=condition_column, time_column=time_column,
                         control_can_be_selected=control_can_be_selected, sample_mask_column=sample_mask_column, **kwargs)


DATACLASS_DICT = {
    "sorting": {
        "Normal": VariantSortingScreenData,
        "MixtureNormal": VariantSortingReporterScreenData,
        "_MixtureNormal": VariantSortingReporterScreenData,
        "MixtureNormal+Acc": VariantSortingReporterScreenData,
        "_MixtureNormal+Acc": VariantSortingReporterScreenData,
        "MixtureNormalConstPi": VariantSortingScreenData,
        "MultiMixtureNormal": TilingSortingReporterScreenData,
        "MultiMixtureNormal+Acc": TilingSortingReporterScreenData,
    },
    "survival": {
        "Normal": VariantSurvivalScreenData,
        "MixtureNormal": VariantSurvivalReporterScreenData,
        "_MixtureNormal": VariantSurvivalReporterScreenData,
        "MixtureNormal+Acc": VariantSurvivalReporterScreenData,
        "_MixtureNormal+Acc": VariantSurvivalReporterScreenData,
        "MultiMixtureNormal": TilingSurvivalReporterScreenData,
        "MultiMixtureNormal+Acc": TilingSurvivalReporterScreenData,
    },
}
