"""The `bean run` flow through the reference-facing API, for every (selection, library design) the CLI offers
(bean/cli/run.py:94-310): identify_model_guide(args) -> DATACLASS_DICT[selection][label](screen=...) ->
run_inference(model, guide, ndata) -> write_result_table(...).  Checks the contract the callers rely on: parameter names,
shapes and positivity (readwrite.py:66-75, build_prior.py:40-46), a finite decreasing loss, and a well-formed
bean_element_result table."""
from types import SimpleNamespace

import numpy as np
import pandas as pd
import pytest
import torch

from crispr_bean_b200.data_class import DATACLASS_DICT
from crispr_bean_b200.readwrite import write_result_table
from crispr_bean_b200.run import identify_model_guide, identify_negctrl_model_guide, run_inference
from crispr_bean_b200.synth import make_sorting_screen, make_survival_screen, make_tiling_screen

pytestmark = pytest.mark.gpu


def golden_screen(case):
    """One of the reference's own test screens (tests/data/*.h5ad), as stored in the reference-computed fixtures."""
    import os

    from tests.helpers import GOLDEN
    from tests.refharness.golden import screen_from_arrays

    return screen_from_arrays(np.load(os.path.join(GOLDEN, f"ref_{case}.npz")))


def cli_args(**kw):
    base = dict(selection="sorting", library_design="variant", uniform_edit=False, scale_by_acc=False, ignore_bcmatch=False,
                dont_fit_noise=False, guide_activity_col=None)
    base.update(kw)
    return SimpleNamespace(**base)


CASES = [
    ("sorting-variant", cli_args(), lambda: make_sorting_screen(30, 4, n_reps=3, seed=1, n_negctrl_guides=12), dict(control_can_be_selected=True)),
    ("sorting-variant-acc", cli_args(scale_by_acc=True), lambda: make_sorting_screen(30, 4, n_reps=3, seed=2, n_negctrl_guides=12, accessibility=True),
     dict(control_can_be_selected=True, accessibility_col="accessibility")),
    ("sorting-uniform-edit", cli_args(uniform_edit=True), lambda: make_sorting_screen(30, 4, n_reps=3, seed=3, n_negctrl_guides=12),
     dict(control_can_be_selected=True, use_bcmatch=True)),
    ("sorting-tiling", cli_args(library_design="tiling"), lambda: make_tiling_screen(n_guides=40, n_reps=3, seed=4),
     dict(control_can_be_selected=True, allele_df_key="allele_counts")),
    ("survival-variant", cli_args(selection="survival"), lambda: make_survival_screen(30, 4, n_reps=3, seed=5, n_negctrl_guides=12),
     dict(control_condition="D7")),
    ("survival-uniform-edit", cli_args(selection="survival", uniform_edit=True),
     lambda: make_survival_screen(30, 4, n_reps=3, seed=6, n_negctrl_guides=12), dict(control_condition="D7", negctrl_guide_idx=list(range(12)))),
    ("survival-tiling", cli_args(selection="survival", library_design="tiling"),
     lambda: make_tiling_screen(n_guides=40, n_reps=3, seed=7, as_survival=True),
     dict(control_condition="D0", allele_df_key="allele_counts", condition_column="condition", time_column="time")),
    # the commands of the reference's tests/test_run.py on its own screens
    ("real-sorting-variant", cli_args(), lambda: golden_screen("real_var_mini_mixture"),
     dict(condition_column="condition", control_can_be_selected=True)),
    ("real-sorting-uniform-edit", cli_args(uniform_edit=True), lambda: golden_screen("real_var_mini_normal"),
     dict(condition_column="condition", control_can_be_selected=True, use_bcmatch=True)),
    ("real-sorting-tiling", cli_args(library_design="tiling"), lambda: golden_screen("tiling_real_mini"),
     dict(condition_column="condition", control_can_be_selected=True, allele_df_key="allele_counts", control_guide_tag=None)),
    ("real-survival-variant", cli_args(selection="survival"), lambda: golden_screen("survival_real_var_mixture"),
     dict(condition_column="condition", time_column="time", control_condition="D7", control_can_be_selected=True)),
    ("real-survival-tiling", cli_args(selection="survival", library_design="tiling"), lambda: golden_screen("survival_tiling_real_mini"),
     dict(condition_column="condition", time_column="time", control_condition="D0", control_can_be_selected=True,
          allele_df_key="allele_counts", control_guide_tag=None)),
]


@pytest.mark.parametrize("name,args,make_screen,data_kw", CASES, ids=[c[0] for c in CASES])
def test_bean_run_flow(cuda_device, tmp_path, name, args, make_screen, data_kw):
    label, model, guide = identify_model_guide(args)
    scr = make_screen()
    ndata = DATACLASS_DICT[args.selection][label](screen=scr, **data_kw)
    steps = 60
    params, hist = run_inference(model, guide, ndata, num_steps=steps, device=cuda_device)
    loss = np.asarray(hist["loss"])
    assert loss.shape == (steps,) and np.isfinite(loss).all()
    if not name.startswith("real-"):  # (10-sample screens: 60 steps of a one-particle ELBO are too noisy to order)
        assert loss[-10:].mean() < loss[:10].mean()
    tiling = args.library_design == "tiling"
    n_elem = ndata.n_edits if tiling else ndata.n_targets
    shape = (n_elem,) if tiling else (n_elem, 1)
    assert tuple(hist["params"]["mu_loc"].shape) == shape and tuple(hist["params"]["mu_scale"].shape) == shape
    assert (hist["params"]["mu_scale"] > 0).all()
    if args.selection == "sorting":
        assert tuple(hist["params"]["sd_loc"].shape) == shape and (hist["params"]["sd_scale"] > 0).all()
    if not args.uniform_edit:
        A = ndata.n_max_alleles if tiling else 2
        assert tuple(hist["params"]["alpha_pi"].shape) == (ndata.n_guides, A) and (hist["params"]["alpha_pi"] > 0).all()
    assert all(v.device.type == "cpu" for v in hist["params"].values())  # run.py:391-396 returns cpu tensors
    # negative-control fit on the control guides (variant designs), as cli/run.py:235-257 does
    neg = None
    if not tiling:
        idx = np.where(ndata.screen.guides["target_group"].to_numpy() == "NegCtrl")[0]
        nm, ng = identify_negctrl_model_guide(args, "X_bcmatch" in scr.layers)
        neg, _ = run_inference(nm, ng, ndata[idx], num_steps=30, device=cuda_device)
        assert neg["mu_loc"].dim() == 0
    names = list(ndata.edit_index) if tiling else list(pd.unique(ndata.screen.guides["target"]))
    info = pd.DataFrame({"n": range(n_elem)}, index=pd.Index(names, name="target"))
    guide_info = ndata.screen.guides.drop(columns=["accessibility"], errors="ignore")  # write_result_table inserts that column itself
    table = write_result_table(info, guide_info, params, model_label=label, prefix=f"{tmp_path}/",
                               negctrl_params=neg, adjust_confidence_by_negative_control=True,
                               adjust_confidence_negatives=np.arange(min(12, n_elem)), sd_is_fitted=args.selection == "sorting",
                               guide_acc=(ndata.guide_accessibility.cpu().numpy() if getattr(ndata, "guide_accessibility", None) is not None else None),
                               return_result=True, is_survival_screen=args.selection == "survival")
    z = "mu_z_adj" if n_elem >= 10 else "mu_z"  # fewer than 10 negatives: no confidence adjustment (readwrite.py:141-150)
    assert len(table) == n_elem and {"mu", "mu_sd", "mu_z", z, "CI[0.025", "0.975]"} <= set(table.columns)
    assert np.isfinite(table[z]).all()
    assert (np.diff(table[z].abs().to_numpy()) <= 1e-12).all()  # sorted by |z|, strongest first
    assert (tmp_path / f"bean_sgRNA_result.{label}.csv").exists()
