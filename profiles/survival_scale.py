"""Step time of the survival MixtureNormal step as the screen grows -- c4 shape (3 replicates x 3 timepoints, 5 guides per
variant) from 3.5k to 1M guides: the fused three-kernel step (`bean_svi_survival_run_*`, what run_inference uses) beside the
round-1 autograd engine (C-ABI site kernels + torch glue replayed from a CUDA graph).

    python profiles/survival_scale.py [n_variants ...]           # one JSON line per size and engine
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
        from crispr_bean_b200.survival_fused import SurvivalFusedEngine

        for name, make in (("fused", lambda: SurvivalFusedEngine(d, "cuda", num_steps=400)),
                           ("autograd", lambda: SurvivalSviEngine(d, "MixtureNormal", "cuda", num_steps=400))):
            eng = make()
            eng.run(110)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            eng.run(100)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 100
            loss = eng.losses()
            cells = d.n_guides * d.n_reps * d.n_condits
            row = {"engine": name, "guides": int(d.n_guides), "ms_per_step": round(ms, 4), "cells_per_s": round(cells / ms * 1e3, 0),
                   "loss_first": float(loss[0]), "loss_last": float(loss[-1]), "finite": bool(torch.isfinite(loss).all()),
                   "synth_plus_tensorise_s": round(t_host, 2)}
            if name == "fused":  # each kernel alone (BeanSviConfig.phases)
                for label, ph in (("guide_ms", 1), ("alpha_ms", 4), ("variant_ms", 2)):
                    eng.cfg.phases = ph
                    a.record()
                    eng.run(50)
                    b.record()
                    torch.cuda.synchronize()
                    row[label] = round(a.elapsed_time(b) / 50, 4)
                eng.cfg.phases = 0
            print(json.dumps(row), flush=True)
            del eng
            torch.cuda.empty_cache()
        del d, scr


if __name__ == "__main__":
    main()
