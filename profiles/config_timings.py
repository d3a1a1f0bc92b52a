"""SVI step time at the BASELINE.json shapes that are NOT the bench workload (c1-c4): GPU engine vs the CPU oracle port.

    python profiles/config_timings.py            # prints one JSON line per configuration

Parity at these shapes is tested in tests/test_gpu_configs.py; this script only reports how long a step takes
(CUDA events over 50 steps after 10 warm-up steps; CPU: 3 steps of the plain-torch oracle on all host threads).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from crispr_bean_b200 import data_class as dc  # noqa: E402
from crispr_bean_b200.synth import make_config, make_survival_screen, make_tiling_screen  # noqa: E402
from oracle import bean_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402


def gpu_ms(eng, steps=50, warm=10):
    eng.run(warm)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    eng.run(steps)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def cpu_ms(fn, data, steps=3, **kw):
    ps, opt, ts = O.ParamStore(), O.ClippedAdam(lr=0.01, lrd=0.1 ** (1 / 2000)), []
    for t in range(steps + 1):
        t0 = time.perf_counter()
        loss, _ = fn(data, ps, **kw)
        ps.zero_grad()
        loss.backward()
        opt.step(ps.unconstrained)
        float(loss.detach())
        if t:
            ts.append(time.perf_counter() - t0)
    return 1e3 * sum(ts) / len(ts)


def main():
    from crispr_bean_b200.generic import TilingSviEngine
    from crispr_bean_b200.survival import SurvivalSviEngine
    from crispr_bean_b200.svi import SviEngine

    torch.set_num_threads(os.cpu_count() or 1)
    dev = "cuda"
    rows = []
    d = H.load_var_mini()
    rows.append(("c1 var_mini Normal (30 guides x 2 x 5)", d, SviEngine(d, "Normal", dev, use_bcmatch=False, num_steps=100),
                 O.elbo_normal, dict(use_bcmatch=False)))
    d = dc.VariantSortingReporterScreenData(make_config("c2_ldlc_variant", seed=101), control_can_be_selected=True)
    rows.append((f"c2 LDL-C MixtureNormal ({d.n_guides} guides x 4 x 5)", d, SviEngine(d, "MixtureNormal", dev, num_steps=100),
                 O.elbo_mixture_normal, {}))
    d = dc.TilingSortingReporterScreenData(make_tiling_screen(n_guides=800, max_alleles=16, n_reps=4, seed=3),
                                           control_can_be_selected=True, allele_df_key="allele_counts")
    from crispr_bean_b200.tiling_fused import TilingFusedEngine

    rows.append((f"c3 tiling MultiMixtureNormal (800 guides x {d.n_max_alleles} alleles, {d.n_edits} edits), fused step", d,
                 TilingFusedEngine(d, dev, num_steps=100), O.elbo_multi_mixture_normal, {}))
    rows.append((f"c3 tiling MultiMixtureNormal (800 guides x {d.n_max_alleles} alleles, {d.n_edits} edits), site-kernel engine (round 1)", d,
                 TilingSviEngine(d, dev, num_steps=100), O.elbo_multi_mixture_normal, {}))
    d = dc.VariantSurvivalReporterScreenData(make_survival_screen(690, "lognormal", n_reps=3, seed=21, n_negctrl_guides=101),
                                             control_condition="D7")
    from crispr_bean_b200.survival_fused import SurvivalFusedEngine

    rows.append((f"c4 survival MixtureNormal ({d.n_guides} guides x 3 x 3), fused step", d, SurvivalFusedEngine(d, dev, num_steps=100),
                 O.elbo_survival_mixture_normal, {}))
    for name, data, eng, fn, kw in rows:
        g, c = gpu_ms(eng), cpu_ms(fn, data, **kw)
        print(json.dumps({"config": name, "gpu_ms_per_step": round(g, 4), "cpu_oracle_ms_per_step": round(c, 2),
                          "speedup": round(c / g, 1), "cpu_threads": torch.get_num_threads()}))


if __name__ == "__main__":
    main()
