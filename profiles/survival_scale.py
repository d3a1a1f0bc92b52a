"""Step time of the survival MixtureNormal engine (autograd engine: C-ABI site kernels + torch glue, CUDA graph) as the
screen grows -- c4 shape (3 replicates x 3 timepoints, 5 guides per variant) from 3.5k to 1M guides.

    python profiles/survival_scale.py            # one JSON line per size
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from crispr_bean_b200 import data_class as dc  # noqa: E402
from crispr_bean_b200.survival import SurvivalSviEngine  # noqa: E402
from crispr_bean_b200.synth import make_survival_screen  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [700, 20_000, 200_000]
    for n_var in sizes:
        t0 = time.perf_counter()
        scr = make_survival_screen(n_var, 5, n_reps=3, seed=21, n_negctrl_guides=100)
        d = dc.VariantSurvivalReporterScreenData(scr, control_condition="D7")
        t_host = time.perf_counter() - t0
        eng = SurvivalSviEngine(d, "MixtureNormal", "cuda", num_steps=200)
        eng.run(10)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        eng.run(50)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 50
        loss = eng.losses()
        cells = d.n_guides * d.n_reps * d.n_condits
        print(json.dumps({"guides": int(d.n_guides), "ms_per_step": round(ms, 4), "cells_per_s": round(cells / ms * 1e3, 0),
                          "loss_first": float(loss[0]), "loss_last": float(loss[-1]), "finite": bool(torch.isfinite(loss).all()),
                          "synth_plus_tensorise_s": round(t_host, 2)}), flush=True)
        del eng, d, scr
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
