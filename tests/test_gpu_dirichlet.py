"""bean_dirichlet_rsample_* / bean_dirichlet_rsample_grad_*: the `pi` draws and their pathwise derivative for A >= 2 alleles.

Forward: distributional parity with torch.distributions.Dirichlet (the reference's own CPU and CUDA streams differ, SURVEY
App. B11) -- every component of a Dirichlet(c) draw is Beta(c_a, C - c_a), so its probability-integral transform must be
uniform (Kolmogorov-Smirnov per concentration band), and the rows must sum to one.  Backward: exactly torch's
`_Dirichlet_backward` (torch._dirichlet_grad evaluated in double)."""
import numpy as np
import pytest
import torch
from scipy import stats

from crispr_bean_b200.dirichlet import DirichletStream, dirichlet_rsample

pytestmark = pytest.mark.gpu


def _stream(dev, seed=5, step=0, offset=0, site=0):
    return DirichletStream(seed, torch.full((1,), step, dtype=torch.int64, device=dev), guide_offset=offset, site=site)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("A", [2, 3, 7, 16, 40])
def test_draws_have_beta_marginals(cuda_device, dtype, A):
    g = torch.Generator().manual_seed(A)
    G, R = 6000, 4
    # concentrations across the sampler's regimes: boosted (< 1), moderate, large.  float32 draws are clamped to
    # [1.2e-38, 1 - 6e-8] like torch's: with concentrations below ~0.3 a visible share of the mass sits beyond those clamps
    # (atoms in the transform), so the float32 case starts at 0.3
    lo_c = 0.05 if dtype == torch.float64 else 0.3
    conc = torch.exp(torch.empty((G, A)).uniform_(np.log(lo_c), np.log(60.0), generator=g)).to(dtype)
    x = dirichlet_rsample(conc.to(cuda_device), R, _stream(cuda_device)).cpu().double()
    assert x.shape == (R, G, A)
    assert torch.allclose(x.sum(-1), torch.ones(R, G, dtype=torch.float64), atol=1e-5 if dtype == torch.float32 else 1e-12)
    assert (x > 0).all() and (x < 1).all()
    c = conc.double()
    C = c.sum(-1, keepdim=True)
    u = stats.beta.cdf(x.numpy(), c.numpy()[None], (C - c).numpy()[None])  # (R, G, A)
    for lo, hi in ((0.05, 0.5), (0.5, 1.0), (1.0, 6.0), (6.0, 60.0)):
        sel = ((c >= lo) & (c < hi)).numpy()[None].repeat(R, 0)
        ub = u[sel]
        if ub.size < 2000:
            continue
        # KS critical value at alpha = 0.01 is 1.63 / sqrt(n); the draws of one row are weakly dependent (they share the row
        # sum), hence the slack
        assert stats.kstest(ub, "uniform").statistic < 1.5 * 1.63 / np.sqrt(ub.size), (A, lo, hi)


def test_draws_depend_on_global_ids_only(cuda_device):
    """A shard draws exactly the rows of the unsharded call; another step / site / seed gives other draws."""
    g = torch.Generator().manual_seed(0)
    conc = (0.2 + 5 * torch.rand((500, 5), generator=g)).to(cuda_device)
    full = dirichlet_rsample(conc, 3, _stream(cuda_device, step=7))
    part = dirichlet_rsample(conc[123:321].contiguous(), 3, _stream(cuda_device, step=7, offset=123))
    assert torch.equal(full[:, 123:321], part)
    assert torch.equal(full, dirichlet_rsample(conc, 3, _stream(cuda_device, step=7)))
    for other in (_stream(cuda_device, step=8), _stream(cuda_device, step=7, site=1), _stream(cuda_device, seed=6, step=7)):
        assert not torch.equal(full, dirichlet_rsample(conc, 3, other))


# fp64 tolerance: torch's saddle-point expression cancels from O((x - mean)^-2) to O(1), which amplifies the 1-ulp differences
# between glibc's and libdevice's log / pow by up to ~1e5 (observed 4e-10 at total concentration 170); 1e-10 holds for the small
# totals of variant designs (A = 2), 1e-8 is asserted for all
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-8), (torch.float32, 1e-5)])
@pytest.mark.parametrize("A", [2, 5, 33])
def test_backward_is_torchs_dirichlet_backward(cuda_device, dtype, tol, A):
    g = torch.Generator().manual_seed(10 + A)
    G, R = 300, 3
    conc = torch.exp(torch.empty((G, A), dtype=torch.float64).uniform_(np.log(0.05), np.log(40.0), generator=g))
    x = torch.distributions.Dirichlet(conc).sample((R,))  # (R, G, A), float64
    gout = torch.randn((R, G, A), generator=g, dtype=torch.float64)
    total = conc.sum(-1, keepdim=True).expand(G, A)
    D = torch._dirichlet_grad(x.contiguous(), conc.expand(R, G, A).contiguous(), total.expand(R, G, A).contiguous())
    ref = (D * (gout - (x * gout).sum(-1, keepdim=True))).sum(0)
    c = conc.to(cuda_device, dtype).requires_grad_(True)
    out = dirichlet_rsample(c, R, _stream(cuda_device), injected=x.to(dtype))
    out.backward(gout.to(cuda_device, dtype))
    got = c.grad.double().cpu()
    err = ((got - ref).abs() / (ref.abs() + ref.abs().mean()))
    i = int(err.argmax())
    assert err.max().item() <= tol, (err.max().item(), got.reshape(-1)[i].item(), ref.reshape(-1)[i].item(), conc.reshape(-1)[i].item(),
                                     conc[i // A].sum().item(), x[:, i // A, i % A].tolist())


def test_capturable_and_step_read_from_device(cuda_device):
    conc = (0.5 + torch.rand((64, 4), device=cuda_device))
    st = _stream(cuda_device, step=0)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(cuda_device)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        dirichlet_rsample(conc, 2, st)
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(graph):
        x = dirichlet_rsample(conc, 2, st)
    draws = []
    for t in range(3):
        st.step.fill_(t)
        graph.replay()
        draws.append(x.clone())
    assert not torch.equal(draws[0], draws[1]) and not torch.equal(draws[1], draws[2])
    st.step.fill_(1)
    assert torch.equal(draws[1], dirichlet_rsample(conc, 2, st))
