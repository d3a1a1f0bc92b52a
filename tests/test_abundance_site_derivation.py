"""The closed form DESIGN.md states for the survival abundance sites (Dirichlet over all guides, same concentration in model
and guide: survival_model.py:306-311, :660-669) against torch autograd -- the formula the fused survival step will use."""
import torch


def test_closed_form_gradient_of_the_abundance_site_pair():
    torch.manual_seed(0)
    G, R = 50, 3
    q0 = torch.randn(G, dtype=torch.float64).requires_grad_(True)
    c = q0.exp()
    conc = c.unsqueeze(0).expand(R, -1)
    obs = torch.rand(R, G, dtype=torch.float64)
    obs = obs / obs.sum(-1, keepdim=True)
    x = torch.distributions.Dirichlet(conc).rsample()
    d_model = torch.distributions.Dirichlet(conc).log_prob(obs).sum()
    d_guide = torch.distributions.Dirichlet(conc).log_prob(x).sum()
    elbo_part = d_model - d_guide
    # the normalisers cancel: only sum (c - 1)(log obs - log x) remains
    assert torch.allclose(elbo_part, ((conc - 1) * (obs.log() - x.log())).sum(), rtol=1e-12)
    (g,) = torch.autograd.grad(elbo_part, q0)
    with torch.no_grad():
        C = c.sum()
        D = torch._dirichlet_grad(x, conc.contiguous(), C.expand_as(conc).contiguous())
        dc = (obs.log() - x.log()).sum(0) + (D * (-(c - 1) / x + (C - G))).sum(0)  # sum_h x_h gout_h = -(C - G)
    assert torch.allclose(g, dc * c, rtol=1e-10, atol=1e-10)
