"""Own kernels of the survival MixtureNormal step timed alone at 1M guides (CUDA events, 20 launches after 3 warm-ups):
`bean_pi_sites` (thread per guide, double), `bean_ll` (float), `bean_latent_sites`, and the whole graph-replayed step.

    python profiles/survival_kernel_times.py [n_variants=200000]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from crispr_bean_b200 import data_class as dc  # noqa: E402
from crispr_bean_b200.latent_sites import latent_sites  # noqa: E402
from crispr_bean_b200.ll_function import launch_ll  # noqa: E402
from crispr_bean_b200.pi_sites import pi_sites  # noqa: E402
from crispr_bean_b200.survival import SurvivalSviEngine  # noqa: E402
from crispr_bean_b200.synth import make_survival_screen  # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n, 4)


def main():
    n_var = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    d = dc.VariantSurvivalReporterScreenData(make_survival_screen(n_var, 5, n_reps=3, seed=21, n_negctrl_guides=100), control_condition="D7")
    eng = SurvivalSviEngine(d, "MixtureNormal", "cuda", num_steps=100)
    G, R = eng.G, eng.R
    out = {"guides": G}
    with torch.no_grad():
        alpha = eng.theta["alpha_pi"].exp()
        scaled = alpha / alpha.sum(-1, keepdim=True) * eng.pi_a0[:, None]
        conc_g, conc_m = scaled.clamp(min=1e-5).contiguous(), scaled.contiguous()
        pi = torch.distributions.Dirichlet(conc_g.unsqueeze(0).unsqueeze(0).expand(R, 1, -1, -1)).sample()
        mu = 0.1 * torch.randn((G, 2), device=eng.device, dtype=eng.dtype)
        out["pi_sites_ms"] = timed(lambda: pi_sites(conc_g, conc_m, pi, eng.pi_data, growth=mu, work_dtype=eng.pi_dtype))
        out["pi_sites_dtype"] = str(eng.pi_dtype)
        pi_g = pi[:, 0].permute(1, 0, 2).contiguous().to(eng.dtype)
        out["ll_ms"] = timed(lambda: launch_ll(eng.screen, mu, torch.ones_like(mu), pi_g, None))
        T = eng.theta["mu_loc"].numel()
        eps = torch.randn(eng.theta["mu_loc"].shape, device=eng.device, dtype=eng.dtype)
        out["latent_sites_ms"] = timed(lambda: latent_sites(eng.theta["mu_loc"].detach(), eng.theta["mu_scale"].detach(), eps, eng.latent_prior))
        out["variants"] = T
    out["step_ms"] = timed(lambda: eng.run(4), n=5, warm=2) / 4
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
