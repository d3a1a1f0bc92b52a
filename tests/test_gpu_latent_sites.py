"""`bean_latent_sites_*` (+ `_grad_*`) against torch.distributions: the `mu_targets` / `sd_targets` sites of the
reference programs (model.py:579-610 model, :893-921 guide).  fp64 1e-12, fp32 2e-5 relative."""
import pytest
import torch
import torch.distributions as tdist

from crispr_bean_b200.latent_sites import LatentPrior, latent_sites

pytestmark = pytest.mark.gpu


def torch_sites(mu_loc, mu_ls, sd_loc, sd_ls, eps_mu, eps_sd, prior, has_sd):
    s = mu_ls.exp()
    mu = mu_loc + s * eps_mu
    v = -tdist.Normal(mu_loc, s).log_prob(mu).sum()
    if prior.get("normal"):
        v = v + tdist.Normal(prior["mu_loc"], prior["mu_scale"]).log_prob(mu).sum()
    else:
        v = v + tdist.Laplace(torch.zeros_like(mu), torch.ones_like(mu)).log_prob(mu).sum()
    sd = None
    if has_sd:
        t = sd_ls.exp()
        sd = torch.exp(sd_loc + t * eps_sd)
        v = v + tdist.LogNormal(prior["sd_loc"], prior["sd_scale"]).log_prob(sd).sum() - tdist.LogNormal(sd_loc, t).log_prob(sd).sum()
    return mu, sd, v


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
@pytest.mark.parametrize("n,shape,has_sd,kind", [(1, (), False, "normal_scalar"), (7, (7, 1), False, "laplace"), (1000, (1000,), True, "laplace"),
                                                (333, (333, 1), True, "normal_scalar"), (333, (333, 1), True, "normal_vector")])
def test_latent_sites_value_draws_and_gradients(cuda_device, dtype, tol, n, shape, has_sd, kind):
    g = torch.Generator().manual_seed(n)
    r = lambda scale=1.0: (scale * torch.randn(shape, generator=g, dtype=torch.float64)).to(cuda_device, dtype)
    leaves = [r(), r(0.3), r(0.2), r(0.3)]
    eps_mu, eps_sd = r(), r()
    if n > 1:
        eps_mu.reshape(-1)[0] = 0.0
        leaves[0].reshape(-1)[0] = 0.0  # mu == 0 exactly: |.|' = 0
    pp, tp = None, {"sd_loc": torch.tensor(0.0, device=cuda_device, dtype=dtype), "sd_scale": torch.tensor(0.01, device=cuda_device, dtype=dtype)}
    if kind == "normal_scalar":
        pp = {"mu_loc": 0.1, "mu_scale": 2.0}
        tp.update(normal=True, mu_loc=torch.tensor(0.1, device=cuda_device, dtype=dtype), mu_scale=torch.tensor(2.0, device=cuda_device, dtype=dtype))
    elif kind == "normal_vector":
        pp = {"mu_loc": r(0.3), "mu_scale": r(0.1).abs() + 0.5, "sd_loc": r(0.05), "sd_scale": r(0.01).abs() + 0.02}
        tp = dict(pp, normal=True)
    prior = LatentPrior(n, pp, 0.01, cuda_device, dtype)
    leaves = [t.requires_grad_(True) for t in leaves]
    ref_mu, ref_sd, ref_v = torch_sites(*leaves, eps_mu, eps_sd, tp, has_sd)
    # a downstream use of the draws, so that all three upstream gradients are exercised
    w_mu, w_sd = r(), r()
    ref_loss = 0.7 * ref_v + (w_mu * ref_mu).sum() + ((w_sd * ref_sd).sum() if has_sd else 0.0)
    used = leaves if has_sd else leaves[:2]
    ref_grads = torch.autograd.grad(ref_loss, used)
    mine = [t.detach().clone().requires_grad_(True) for t in leaves]
    out = latent_sites(mine[0], mine[1], eps_mu, prior, *( (mine[2], mine[3], eps_sd) if has_sd else ()))
    mu, sd, v = (out if has_sd else (out[0], None, out[1]))
    assert mu.shape == leaves[0].shape
    rel = lambda a, b: ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()
    assert rel(mu, ref_mu) <= tol and (not has_sd or rel(sd, ref_sd) <= tol)
    assert abs(v.item() - ref_v.item()) <= tol * max(abs(ref_v.item()), 1.0), (v.item(), ref_v.item())
    loss = 0.7 * v + (w_mu * mu).sum() + ((w_sd * sd).sum() if has_sd else 0.0)
    grads = torch.autograd.grad(loss, mine if has_sd else mine[:2])
    for name, a, b in zip(("mu_loc", "mu_log_scale", "sd_loc", "sd_log_scale"), grads, ref_grads):
        assert a.shape == b.shape
        assert rel(a, b) <= 10 * tol, (name, rel(a, b))
