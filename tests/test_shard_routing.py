"""Which (model, screen) pairs `run_inference` splits over the ranks of torch.distributed (run.shards_over_ranks): host logic,
no GPU.  Variant designs shard by variant blocks; the tiling sorting design by guide blocks where the fused step serves it;
ControlNormal, the covariate model, `+Acc` tiling and wide allele tables run as replicas."""
from functools import partial

from crispr_bean_b200 import model as sm
from crispr_bean_b200 import survival_model as svm
from crispr_bean_b200.data_class import TilingSortingReporterScreenData
from crispr_bean_b200.run import shards_over_ranks
from crispr_bean_b200.synth import make_tiling_screen
from tests import helpers as H


def test_variant_designs_shard_and_control_model_does_not():
    data = H.make_small_mixture_data(n_variants=6, n_reps=2)
    assert shards_over_ranks(sm.MixtureNormalModel, data)
    assert shards_over_ranks(sm.NormalModel, data)
    assert not shards_over_ranks(sm.ControlNormalModel, data)
    assert shards_over_ranks(svm.MixtureNormalModel, data)


def test_tiling_design_shards_only_where_the_fused_step_serves_it():
    scr = make_tiling_screen(n_guides=12, max_alleles=5, n_reps=2, seed=1)
    data = TilingSortingReporterScreenData(scr, control_can_be_selected=True, allele_df_key="allele_counts")
    assert shards_over_ranks(sm.MultiMixtureNormalModel, data)
    assert not shards_over_ranks(partial(sm.MultiMixtureNormalModel, scale_by_accessibility=True), data)
    data.n_max_alleles = 40  # wider than one lane per allele: the site-kernel engine, unsharded
    assert not shards_over_ranks(sm.MultiMixtureNormalModel, data)
