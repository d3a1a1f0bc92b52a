"""Seeded synthetic screens of the five BASELINE.json shapes.

Produces a `MiniScreen` (counts + `X_bcmatch` + `edits` layers, guide and sample tables) that goes
through the normal tensoriser, so the synthetic path exercises the same code as a real screen.
Generation model: variants with a true effect mu_v ~ Laplace(0,1) on 10 % of variants, sd 1; editing
rate pi_g ~ Beta(2,5); per-guide depth ~ LogNormal; per (rep, guide) bin fractions ~ Dirichlet(a0 * p)
and Poisson counts around depth * fraction; X_bcmatch ~ Binomial(X, 1/2); edits ~ Binomial(X_bcmatch, pi).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import pandas as pd
import torch

from .screen import MiniScreen

DEFAULT_BINS: Tuple[Tuple[float, float], ...] = ((0.0, 0.2), (0.2, 0.4), (0.6, 0.8), (0.8, 1.0))  # docs/model.rst:226


def _phi(z):
    return 0.5 * (1 + torch.erf(z / math.sqrt(2)))


def variant_lengths(n_variants: int, guides_per_variant, gen: torch.Generator) -> torch.Tensor:
    if isinstance(guides_per_variant, int):
        return torch.full((n_variants,), guides_per_variant, dtype=torch.int64)
    # c2 statistics: median 5, clipped to [3, 135] (tests/data/test_guide_info.csv)
    ln = torch.exp(math.log(5.0) + 0.35 * torch.randn(n_variants, generator=gen))
    return ln.round().clamp(3, 135).long()


def make_sorting_screen(n_variants: int, guides_per_variant=5, n_reps: int = 4,
                        bins: Sequence[Tuple[float, float]] = DEFAULT_BINS, depth: float = 300.0,
                        seed: int = 101, n_negctrl_guides: int = 0, frac_effect: float = 0.1,
                        accessibility: bool = False) -> MiniScreen:
    """Sorting screen with a bulk (control) sample per replicate and reporter layers."""
    gen = torch.Generator().manual_seed(seed)
    lengths = variant_lengths(n_variants, guides_per_variant, gen)
    if n_negctrl_guides:
        lengths = torch.cat([torch.tensor([n_negctrl_guides]), lengths])
    T = len(lengths)
    G = int(lengths.sum())
    gv = torch.repeat_interleave(torch.arange(T), lengths)
    mu_v = torch.distributions.Laplace(0.0, 1.0).sample((T,)) if False else None  # (global RNG not used)
    lap_u = torch.rand(T, generator=gen) - 0.5
    mu_v = -torch.sign(lap_u) * torch.log1p(-2 * lap_u.abs())
    mu_v = mu_v * (torch.rand(T, generator=gen) < frac_effect)
    if n_negctrl_guides:
        mu_v[0] = 0.0
    g1 = torch._standard_gamma(torch.full((G,), 2.0), generator=gen)
    g2 = torch._standard_gamma(torch.full((G,), 5.0), generator=gen)
    pi_g = g1 / (g1 + g2)
    n_g = torch.exp(math.log(depth) + 1.0 * torch.randn(G, generator=gen)).clamp(min=5.0)

    B = len(bins)
    uq = torch.tensor([b[1] for b in bins], dtype=torch.float64)
    lq = torch.tensor([b[0] for b in bins], dtype=torch.float64)
    tu = torch.where(uq == 1, torch.tensor(math.inf, dtype=torch.float64), torch.erfinv(2 * uq.clamp(max=1 - 1e-16) - 1) * math.sqrt(2))
    tl = torch.where(lq == 0, torch.tensor(-math.inf, dtype=torch.float64), torch.erfinv(2 * lq.clamp(min=1e-300) - 1) * math.sqrt(2))
    mu_g = mu_v[gv].double()
    p_edit = _phi(tu[:, None] - mu_g[None, :]) - _phi(tl[:, None] - mu_g[None, :])  # (B, G)
    p_wt = (uq - lq)[:, None]
    p = ((1 - pi_g.double())[None, :] * p_wt + pi_g.double()[None, :] * p_edit)
    p = (p / p.sum(0, keepdim=True)).float()  # (B, G)

    a0_g = torch.exp(-1.510 + 0.7861 * torch.log(n_g * B))  # get_alpha0.py:105 trend
    conc = (a0_g[None, :] * p).clamp(min=1e-3)  # (B, G)
    s_rb = 0.7 + 0.6 * torch.rand(n_reps, B + 1, generator=gen)  # per-sample depth factor, last = bulk
    frac = torch._standard_gamma(conc[None].expand(n_reps, B, G).contiguous(), generator=gen)
    frac = frac / frac.sum(1, keepdim=True)
    lam = n_g[None, None, :] * s_rb[:, :B, None] * frac * B
    X_sort = torch.poisson(lam, generator=gen)  # (R, B, G)
    X_bulk = torch.poisson(n_g[None, :] * s_rb[:, B, None], generator=gen)  # (R, G)
    X = torch.cat([X_sort, X_bulk[:, None, :]], dim=1)  # (R, B+1, G)
    Xbc = torch.binomial(X, torch.full_like(X, 0.5), generator=gen)
    edits = torch.binomial(Xbc, pi_g[None, None, :].expand_as(Xbc).contiguous(), generator=gen)

    cond_names = [f"bin{j}" for j in range(B)] + ["bulk"]
    rows = []
    for r in range(n_reps):
        for j in range(B + 1):
            lo, hi = (bins[j] if j < B else (0.0, 1.0))
            rows.append((f"rep{r}_{cond_names[j]}", f"rep{r}", cond_names[j], lo, hi, 1))
    samples = pd.DataFrame(rows, columns=["name", "replicate", "bin", "lower_quantile", "upper_quantile", "mask"]).set_index("name")
    width = max(6, len(str(T)))
    tnames = np.char.add("v", np.char.zfill(np.arange(T).astype(str), width))
    if n_negctrl_guides:
        tnames[0] = "CONTROL"
    guides = pd.DataFrame({
        "target": tnames[gv.numpy()],
        "target_group": np.where((gv.numpy() == 0) & (n_negctrl_guides > 0), "NegCtrl", "Variant"),
    }, index=pd.Index(np.char.add("g", np.arange(G).astype(str)), name="name"))
    guides["true_mu"] = mu_v[gv].numpy()
    guides["true_pi"] = pi_g.numpy()
    if accessibility:
        guides["accessibility"] = torch.exp(1.0 + 0.8 * torch.randn(G, generator=gen)).numpy()

    def flat(t):  # (R, S, G) -> (G, R*S) sample-major columns matching `samples`
        return t.permute(2, 0, 1).reshape(G, -1).numpy().astype(np.float32)

    return MiniScreen(flat(X), guides, samples, {"X_bcmatch": flat(Xbc), "edits": flat(edits)})


# The five BASELINE.json configurations (SURVEY section 8 shape table).  c1 is the reference's CSV
# fixture (loaded by tests from tests/golden, not generated); c3/c4 are defined where the tiling /
# survival paths are built.
CONFIGS = {
    "c2_ldlc_variant": dict(n_variants=690, guides_per_variant="lognormal", n_reps=4, n_negctrl_guides=101),
    "c5_genome_scale": dict(n_variants=200_000, guides_per_variant=5, n_reps=8),
}


def make_config(name: str, seed: int = 101, **override) -> MiniScreen:
    kw = dict(CONFIGS[name])
    kw.update(override)
    return make_sorting_screen(seed=seed, **kw)


def make_tiling_screen(n_guides: int = 200, window: int = 6, max_alleles: int = 5, n_reps: int = 3,
                       bins: Sequence[Tuple[float, float]] = DEFAULT_BINS, depth: float = 400.0, seed: int = 101,
                       frac_effect: float = 0.2, accessibility: bool = False, as_survival: bool = False) -> MiniScreen:
    """Tiling screen (c3 shape): guide g can edit positions [g, g + window); each of its 1..max_alleles-1 edited
    alleles is a small subset of those positions, so edits are shared by overlapping guides.
    `uns["allele_counts"]` holds the per-sample allele-count table `bean filter` would write."""
    gen = torch.Generator().manual_seed(seed)
    n_pos = n_guides + window
    lap_u = torch.rand(n_pos, generator=gen) - 0.5
    mu_edit = -torch.sign(lap_u) * torch.log1p(-2 * lap_u.abs()) * (torch.rand(n_pos, generator=gen) < frac_effect)
    B = len(bins)
    uq = torch.tensor([b[1] for b in bins], dtype=torch.float64)
    lq = torch.tensor([b[0] for b in bins], dtype=torch.float64)
    tu = torch.where(uq == 1, torch.tensor(math.inf, dtype=torch.float64), torch.erfinv(2 * uq.clamp(max=1 - 1e-16) - 1) * math.sqrt(2))
    tl = torch.where(lq == 0, torch.tensor(-math.inf, dtype=torch.float64), torch.erfinv(2 * lq.clamp(min=1e-300) - 1) * math.sqrt(2))
    n_all = torch.randint(1, max_alleles, (n_guides,), generator=gen)  # edited alleles per guide
    rows, alleles_of = [], []
    for g in range(n_guides):
        al = []
        for j in range(int(n_all[g])):
            k = int(torch.randint(1, 3, (1,), generator=gen))
            pos = (g + torch.randperm(window, generator=gen)[:k]).sort().values.tolist()
            if pos not in al:
                al.append(pos)
        alleles_of.append(al)
    n_g = torch.exp(math.log(depth) + 0.7 * torch.randn(n_guides, generator=gen)).clamp(min=20.0)
    s_rb = 0.7 + 0.6 * torch.rand(n_reps, B + 1, generator=gen)
    X = torch.zeros(n_reps, B + 1, n_guides)
    cond_names = [f"bin{j}" for j in range(B)] + ["bulk"]
    sample_names = [f"rep{r}_{c}" for r in range(n_reps) for c in cond_names]
    table = []
    frac_all = []
    for g in range(n_guides):
        al = alleles_of[g]
        conc = torch.cat([torch.tensor([6.0]), torch.full((len(al),), 1.5)])
        gam = torch._standard_gamma(conc, generator=gen)
        frac = gam / gam.sum()  # allele fractions incl. WT
        frac_all.append(frac)
        mus = torch.cat([torch.zeros(1), torch.stack([mu_edit[p].sum() for p in al])]).double()
        sds = torch.cat([torch.ones(1), torch.tensor([math.sqrt(len(p)) for p in al])]).double()
        P = _phi((tu[:, None] - mus[None]) / sds[None]) - _phi((tl[:, None] - mus[None]) / sds[None])  # (B, A_g)
        p = (P * frac.double()[None]).sum(-1)
        p = (p / p.sum()).float()
        lam = n_g[g] * s_rb[:, :B] * p[None, :] * B
        X[:, :B, g] = torch.poisson(lam, generator=gen)
        X[:, B, g] = torch.poisson(n_g[g] * s_rb[:, B], generator=gen)
    Xbc = torch.binomial(X, torch.full_like(X, 0.6), generator=gen)
    for g in range(n_guides):
        frac = frac_all[g]
        for j, pos in enumerate(alleles_of[g]):
            cnt = torch.binomial(Xbc[:, :, g], torch.full_like(Xbc[:, :, g], float(frac[j + 1])), generator=gen)  # (R, B+1)
            # the reference's Edit string: "<abs pos>:<pos relative to the guide>:<strand>:<ref>><alt>" (framework/Edit.py:129-134)
            table.append([f"g{g}", ",".join(f"{p}:{p - g}:+:A>G" for p in pos)] + cnt.reshape(-1).tolist())
    allele_df = pd.DataFrame(table, columns=["guide", "allele"] + sample_names)
    rows = []
    for r in range(n_reps):
        for j in range(B + 1):
            lo, hi = (bins[j] if j < B else (0.0, 1.0))
            rows.append((f"rep{r}_{cond_names[j]}", f"rep{r}", cond_names[j], lo, hi, 1))
    samples = pd.DataFrame(rows, columns=["name", "replicate", "bin", "lower_quantile", "upper_quantile", "mask"]).set_index("name")
    guides = pd.DataFrame({"start_pos": np.arange(n_guides)}, index=pd.Index([f"g{g}" for g in range(n_guides)], name="name"))
    if accessibility:
        guides["accessibility"] = torch.exp(1.0 + 0.8 * torch.randn(n_guides, generator=gen)).numpy()
    if as_survival:  # the same counts read as a time course: condition D<7j>, time 7j (shape / parity tests of the survival tiling model)
        samples["condition"] = [f"D{7 * j}" for _ in range(n_reps) for j in range(B + 1)]
        samples["time"] = [7.0 * j for _ in range(n_reps) for j in range(B + 1)]
        samples = samples.drop(columns=["bin", "lower_quantile", "upper_quantile"])

    def flat(t):
        return t.permute(2, 0, 1).reshape(n_guides, -1).numpy().astype(np.float32)

    scr = MiniScreen(flat(X), guides, samples, {"X_bcmatch": flat(Xbc), "edits": np.zeros((n_guides, len(samples)), dtype=np.float32)},
                     uns={"allele_counts": allele_df, "true_mu_edit": {f"{p}:A>G": float(mu_edit[p]) for p in range(n_pos)}})
    return scr


def make_survival_screen(n_variants: int, guides_per_variant=5, n_reps: int = 3, times: Sequence[float] = (0.0, 7.0, 14.0),
                         depth: float = 300.0, seed: int = 101, n_negctrl_guides: int = 0, frac_effect: float = 0.3,
                         accessibility: bool = False) -> MiniScreen:
    """Proliferation screen (c4 shape): samples = replicates x timepoints (`condition` D<t>, `time` t); a guide's
    abundance grows as exp(mu t / max t) in its edited cells.  Same layers as the sorting screens."""
    gen = torch.Generator().manual_seed(seed)
    lengths = variant_lengths(n_variants, guides_per_variant, gen)
    if n_negctrl_guides:
        lengths = torch.cat([torch.tensor([n_negctrl_guides]), lengths])
    T = len(lengths)
    G = int(lengths.sum())
    gv = torch.repeat_interleave(torch.arange(T), lengths)
    mu_v = 0.8 * torch.randn(T, generator=gen) * (torch.rand(T, generator=gen) < frac_effect)
    if n_negctrl_guides:
        mu_v[0] = 0.0
    g1 = torch._standard_gamma(torch.full((G,), 2.0), generator=gen)
    g2 = torch._standard_gamma(torch.full((G,), 5.0), generator=gen)
    pi_g = g1 / (g1 + g2)
    n_g = torch.exp(math.log(depth) + 0.8 * torch.randn(G, generator=gen)).clamp(min=5.0)
    tt = torch.tensor(times) / max(times)
    S = len(times)
    growth = (1 - pi_g)[None, :] + pi_g[None, :] * torch.exp(mu_v[gv][None, :] * tt[:, None])  # (S, G)
    s_rb = 0.7 + 0.6 * torch.rand(n_reps, S, generator=gen)
    over = torch._standard_gamma(torch.full((n_reps, S, G), 8.0), generator=gen) / 8.0  # overdispersion
    X = torch.poisson(n_g[None, None, :] * s_rb[:, :, None] * growth[None] * over, generator=gen)
    Xbc = torch.binomial(X, torch.full_like(X, 0.5), generator=gen)
    p_edit = (pi_g[None, :] * torch.exp(mu_v[gv][None, :] * tt[:, None]) / growth)[None].expand_as(Xbc).contiguous()
    edits = torch.binomial(Xbc, p_edit, generator=gen)
    cond = [f"D{int(t)}" for t in times]
    rows = [(f"rep{r}_{cond[j]}", f"rep{r}", cond[j], float(times[j]), 1) for r in range(n_reps) for j in range(S)]
    samples = pd.DataFrame(rows, columns=["name", "replicate", "condition", "time", "mask"]).set_index("name")
    width = max(6, len(str(T)))
    tnames = np.char.add("v", np.char.zfill(np.arange(T).astype(str), width))
    if n_negctrl_guides:
        tnames[0] = "CONTROL"
    guides = pd.DataFrame({
        "target": tnames[gv.numpy()],
        "target_group": np.where((gv.numpy() == 0) & (n_negctrl_guides > 0), "NegCtrl", "Variant"),
    }, index=pd.Index(np.char.add("g", np.arange(G).astype(str)), name="name"))
    guides["true_mu"] = mu_v[gv].numpy()
    guides["true_pi"] = pi_g.numpy()
    if accessibility:
        guides["accessibility"] = torch.exp(1.0 + 0.8 * torch.randn(G, generator=gen)).numpy()

    def flat(t):
        return t.permute(2, 0, 1).reshape(G, -1).numpy().astype(np.float32)

    return MiniScreen(flat(X), guides, samples, {"X_bcmatch": flat(Xbc), "edits": flat(edits)})
