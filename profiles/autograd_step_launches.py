"""Kernel launches of ONE eager SVI step of the autograd engines (tiling c3, survival c4), for `ncu --metrics
gpu__time_duration.sum`:  the step is bracketed by cudaProfilerStart/Stop so only it is listed.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/autograd_launches.csv python profiles/autograd_step_launches.py survival
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from crispr_bean_b200 import data_class as dc  # noqa: E402
from crispr_bean_b200.synth import make_survival_screen, make_tiling_screen  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "survival"
    if which == "survival":
        from crispr_bean_b200.survival import SurvivalSviEngine

        n_var = int(sys.argv[2]) if len(sys.argv) > 2 else 690  # 200000 -> the 1M-guide survival screen
        d = dc.VariantSurvivalReporterScreenData(make_survival_screen(n_var, "lognormal" if n_var == 690 else 5, n_reps=3, seed=21,
                                                                      n_negctrl_guides=101), control_condition="D7")
        eng = SurvivalSviEngine(d, "MixtureNormal", "cuda", num_steps=100)
    else:
        from crispr_bean_b200.generic import TilingSviEngine

        d = dc.TilingSortingReporterScreenData(make_tiling_screen(n_guides=800, max_alleles=16, n_reps=4, seed=3),
                                               control_can_be_selected=True, allele_df_key="allele_counts")
        eng = TilingSviEngine(d, "cuda", num_steps=100)
    eng.run(3, use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    eng.run(1, use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
