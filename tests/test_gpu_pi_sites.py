"""`bean_pi_sites_f32/f64` against the torch.distributions expressions of the reference programs (model.py:652-670,
:938-950; survival_model.py:313-346): value and autograd gradients w.r.t. the concentrations, the draws and the growth
rates, evaluated by torch in the same dtype.  fp64: 1e-11 (value) / 1e-8 (gradients; (h - hbar) / S cancels at 231 alleles) relative; fp32: 2e-5 / 5e-4."""
import pytest
import torch
import torch.distributions as tdist

from crispr_bean_b200.generic import _multinomial_log_prob
from crispr_bean_b200.pi_sites import PiSiteData, pi_sites

pytestmark = pytest.mark.gpu


def make_case(G, R, A, C, seed, dev, growth):
    g = torch.Generator().manual_seed(seed)
    conc_m = torch.rand((G, A), generator=g, dtype=torch.float64) * 3 + 0.05
    conc_m[::3, -1] = 1e-5  # a non-existent allele (tiling: concentration eps)
    conc_m[1, 0] = 1.0      # xlogy(0, pi)
    conc_g = conc_m.clamp(min=1e-5) * (1 + 0.1 * torch.rand((G, A), generator=g, dtype=torch.float64))
    gam = torch._standard_gamma(conc_g.expand(R, 1, G, A).contiguous(), generator=g).clamp(min=1e-300)
    pi = gam / gam.sum(-1, keepdim=True)
    pi[0, 0, 0, 0] = 1e-20  # below the float64 eps after normalisation: clamped probability, zero gradient
    counts = torch.poisson(torch.rand((R, C, G, A), generator=g, dtype=torch.float64) * 30, generator=g)
    counts[:, :, ::3, -1] = 0
    mask = torch.rand((R, G), generator=g) < 0.8
    mu = 0.3 * torch.randn((G, A), generator=g, dtype=torch.float64) if growth else None
    tc = torch.tensor([0.5, 1.0][:C], dtype=torch.float64) if growth else None
    to = lambda t: None if t is None else t.to(dev)
    return tuple(map(to, (conc_g, conc_m, pi, counts, mask, mu, tc)))


def torch_value(conc_g, conc_m, pi, counts, mask, mu, tc, mask_guide_site):
    R, _, G, A = pi.shape
    C = counts.shape[1]
    m = mask.unsqueeze(1)  # (R, 1, G)
    cg, cm = conc_g.expand(R, 1, G, A), conc_m.expand(R, 1, G, A)
    lp_g = tdist.Dirichlet(cg, validate_args=False).log_prob(pi)
    lp_m = tdist.Dirichlet(cm, validate_args=False).log_prob(pi)
    q = pi.expand(-1, C, -1, -1)
    if mu is not None:
        q = q * torch.exp(mu.unsqueeze(0).unsqueeze(0) * tc.reshape(1, C, 1, 1))
    lp_x = _multinomial_log_prob(q, counts)
    z = torch.zeros((), dtype=pi.dtype, device=pi.device)
    val = torch.where(m, lp_m, z).sum() + torch.where(m.expand(lp_x.shape), lp_x, z).sum()
    return val - (torch.where(m, lp_g, z).sum() if mask_guide_site else lp_g.sum())


@pytest.mark.parametrize("G,R,A,C,growth,mask_guide_site", [(37, 3, 2, 1, True, False), (37, 3, 2, 2, True, False), (21, 4, 7, 1, False, True),
                                                          (9, 2, 231, 1, False, True), (9, 2, 45, 2, True, True)])
@pytest.mark.parametrize("dtype,tol_v,tol_g", [(torch.float64, 1e-11, 1e-8), (torch.float32, 2e-5, 5e-4)])
def test_pi_sites_value_and_gradients(cuda_device, G, R, A, C, growth, mask_guide_site, dtype, tol_v, tol_g):
    conc_g, conc_m, pi, counts, mask, mu, tc = make_case(G, R, A, C, seed=G + A, dev=cuda_device, growth=growth)
    if dtype == torch.float32:
        pi = pi.clamp(min=float(torch.finfo(torch.float32).tiny))
    # the reference expression in the SAME dtype (the probability clamp [eps, 1 - eps] follows it)
    leaves = [t.to(dtype).requires_grad_(True) for t in (conc_g, conc_m, pi)] + ([mu.to(dtype).requires_grad_(True)] if growth else [])
    ref = torch_value(*leaves[:3], counts.to(dtype), mask, leaves[3] if growth else None, None if tc is None else tc.to(dtype), mask_guide_site)
    ref_grads = torch.autograd.grad(ref, leaves)
    data = PiSiteData(counts.to(dtype), mask, tc, mask_guide_site)
    mine = [t.detach().to(dtype).requires_grad_(True) for t in leaves]
    got = pi_sites(mine[0], mine[1], mine[2], data, growth=mine[3] if growth else None)
    assert got.dtype == dtype
    assert abs(got.item() - ref.item()) <= tol_v * abs(ref.item()), (got.item(), ref.item())
    grads = torch.autograd.grad(got * 2.0, mine)  # upstream factor 2: backward must scale
    for name, a, b in zip(("conc_guide", "conc_model", "pi", "growth"), grads, ref_grads):
        a, b = a.double() / 2.0, b.double()
        if dtype == torch.float32 and name == "pi":
            sel = leaves[2].detach() > 1e-6  # d/d pi ~ (conc - 1) / pi: compare where float32 can resolve the draw
            a, b = a[sel], b[sel]
        assert torch.isfinite(a).all(), name
        err = ((a - b).abs() / (b.abs() + b.abs().mean())).max().item()
        assert err <= tol_g, (name, err)


def test_bad_arguments(cuda_device):
    import ctypes as C

    from crispr_bean_b200 import _lib

    args = _lib.BeanPiSitesArgs()
    assert _lib.lib().bean_pi_sites_f64(C.byref(args), None) == -1
    assert b"sizes must be positive" in _lib.lib().bean_last_error()
    with pytest.raises(AssertionError):
        PiSiteData(torch.zeros(2, 1, 3, 2), torch.ones(2, 3, dtype=torch.bool))  # CPU tensor: no fallback
