"""The pyro-primitive shim (tests/refharness/pyro) against pyro's documented semantics -- the part of the reference
stack that is restated rather than executed.  Each test states the pyro behaviour it pins (pyro-ppl 1.8.x docs / source)."""
import math
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refharness"))
import pyro  # noqa: E402  (the shim)
import pyro.distributions as dist  # noqa: E402
from pyro import poutine  # noqa: E402
from torch.distributions import constraints  # noqa: E402

assert os.path.dirname(os.path.abspath(pyro.__file__)).startswith(HERE), "these tests are about the shim, not a real pyro"


def trace_of(fn, *args):
    return poutine.trace(fn).get_trace(*args)


def test_plate_allocates_dims_from_the_right_and_broadcasts_the_site():
    """`with plate(a, 1): with plate(b, T): sample(Laplace(0,1))` -> batch shape (T, 1): un-dimmed plates take the first
    free dim from -1 leftwards, and a site's distribution is expanded to its plates' sizes (BroadcastMessenger)."""
    def model():
        with pyro.plate("a", 1):
            with pyro.plate("b", 7):
                return pyro.sample("x", dist.Laplace(0, 1))

    tr = trace_of(model)
    assert tr.nodes["x"]["value"].shape == (7, 1)
    assert tr.nodes["x"]["fn"].batch_shape == (7, 1)


def test_plate_with_explicit_dim_and_event_shape():
    """Plate dims index the BATCH shape (event dims excluded): Dirichlet((G,)) inside plate(R, dim=-1) -> sample (R, G)."""
    def model():
        with pyro.plate("r", 3, dim=-1):
            return pyro.sample("q", dist.Dirichlet(torch.ones(5)))

    assert trace_of(model).nodes["q"]["value"].shape == (3, 5)


def test_plate_size_mismatch_raises():
    def model():
        with pyro.plate("r", 3, dim=-1):
            pyro.sample("x", dist.Normal(torch.zeros(4), 1.0))

    with pytest.raises(ValueError):
        trace_of(model)


def test_plate_collision_and_reuse():
    """Two active plates may not share a dim; a plate object can be re-entered after it was left (the reference re-enters
    `replicate_plate` three times per model call)."""
    p = pyro.plate("r", 2, dim=-2)

    def ok():
        with p:
            pyro.sample("a", dist.Normal(torch.zeros(2, 1), 1.0))
        with p:
            pyro.sample("b", dist.Normal(torch.zeros(2, 1), 1.0))

    trace_of(ok)

    def bad():
        with pyro.plate("r", 2, dim=-1):
            with pyro.plate("s", 2, dim=-1):
                pyro.sample("a", dist.Normal(0.0, 1.0))

    with pytest.raises(ValueError):
        trace_of(bad)


def test_plate_yields_indices():
    with pyro.plate("b", 4, dim=-2) as idx:
        assert torch.equal(idx, torch.arange(4))


def test_mask_zeroes_log_prob_and_nests_by_and():
    """poutine.mask: log_prob -> where(mask, log_prob, 0); nested masks combine with &."""
    x = torch.tensor([0.5, -1.0, 2.0, 0.1])
    m1 = torch.tensor([True, True, False, True])
    m2 = torch.tensor([True, False, True, True])

    def model():
        with poutine.mask(mask=m1), poutine.mask(mask=m2):
            pyro.sample("x", dist.Normal(torch.zeros(4), 1.0), obs=x)

    tr = trace_of(model)
    tr.compute_log_prob()
    full = dist.Normal(torch.zeros(4), 1.0).log_prob(x)
    assert torch.equal(tr.nodes["x"]["log_prob"], torch.where(m1 & m2, full, torch.zeros(())))


def test_param_is_stored_unconstrained_and_initialised_once():
    """pyro.param(name, init, constraint=positive): the store keeps log(init); later calls ignore `init`."""
    pyro.clear_param_store()
    a = pyro.param("a", torch.tensor([2.0, 3.0]), constraint=constraints.positive)
    assert torch.allclose(a, torch.tensor([2.0, 3.0]))
    u = pyro.get_param_store().unconstrained("a")
    assert torch.allclose(u, torch.tensor([2.0, 3.0]).log()) and u.requires_grad
    b = pyro.param("a", torch.tensor([9.0, 9.0]), constraint=constraints.positive)
    assert torch.allclose(b, a)
    with torch.no_grad():
        u += 1.0
    assert torch.allclose(pyro.param("a"), torch.tensor([2.0, 3.0]) * math.e)


def test_replay_reuses_guide_values_but_leaves_observed_sites_alone():
    """Trace_ELBO runs the model under replay(guide_trace): latent sites take the guide's draw; a site the model OBSERVES
    keeps its observation even if the guide sampled the same name (the survival MixtureNormal quirk)."""
    obs = torch.tensor([0.2, 0.8])

    def guide():
        pyro.sample("z", dist.Normal(0.0, 1.0))
        pyro.sample("w", dist.Dirichlet(torch.ones(2)))

    def model():
        z = pyro.sample("z", dist.Normal(5.0, 1.0))
        w = pyro.sample("w", dist.Dirichlet(torch.ones(2)), obs=obs)
        u = pyro.sample("u", dist.Normal(0.0, 1.0))  # model-only latent: drawn from its prior
        return z, w, u

    gt = trace_of(guide)
    mt = poutine.trace(poutine.replay(model, trace=gt)).get_trace()
    assert mt.nodes["z"]["value"] is gt.nodes["z"]["value"]
    assert torch.equal(mt.nodes["w"]["value"], obs)
    assert "u" in mt.nodes and "u" not in gt.nodes


def test_trace_elbo_is_minus_model_plus_guide_log_probs_with_reparameterised_gradient():
    """One-particle Trace_ELBO for a fully reparameterised guide = -(sum log p - sum log q), differentiated pathwise.
    Conjugate check: model z ~ N(0,1), x ~ N(z,1) obs; guide N(loc, scale): analytic gradient of the EXPECTED loss w.r.t.
    loc is (2 loc - x); averaged over draws the estimator matches it."""
    pyro.clear_param_store()
    x = torch.tensor(1.3)

    def model():
        z = pyro.sample("z", dist.Normal(0.0, 1.0))
        pyro.sample("x", dist.Normal(z, 1.0), obs=x)

    def guide():
        loc = pyro.param("loc", torch.tensor(0.4))
        scale = pyro.param("scale", torch.tensor(0.7), constraint=constraints.positive)
        pyro.sample("z", dist.Normal(loc, scale))

    elbo = pyro.infer.Trace_ELBO()
    torch.manual_seed(0)
    grads, n = 0.0, 2000
    for _ in range(n):
        loss = elbo.differentiable_loss(model, guide)
        mt, gt = elbo.last_traces
        z = gt.nodes["z"]["value"]
        expect = -(dist.Normal(0.0, 1.0).log_prob(z) + dist.Normal(z, 1.0).log_prob(x)
                   - dist.Normal(pyro.param("loc"), pyro.param("scale")).log_prob(z))
        assert torch.allclose(loss, expect)
        loss.backward()
        g = pyro.get_param_store().unconstrained("loc").grad
        grads += g.item()
        g.zero_()
        pyro.get_param_store().unconstrained("scale").grad = None
    assert abs(grads / n - (2 * 0.4 - 1.3)) < 0.07


def test_clipped_adam_equals_torch_adam_on_clamped_gradients_with_decayed_lr():
    """pyro.optim.ClippedAdam: lr <- lr * lrd before every step, gradient clamped to [-clip_norm, clip_norm], then Adam."""
    torch.manual_seed(1)
    p = torch.randn(5, requires_grad=True)
    q = p.detach().clone().requires_grad_(True)
    opt = pyro.optim.ClippedAdam({"lr": 0.01, "lrd": 0.9, "clip_norm": 0.5})
    ref = torch.optim.Adam([q], lr=0.01)
    for t in range(6):
        g = torch.randn(5) * 3
        p.grad = g.clone()
        opt([p])
        for grp in ref.param_groups:
            grp["lr"] = 0.01 * 0.9 ** (t + 1)
        q.grad = g.clamp(-0.5, 0.5)
        ref.step()
        assert torch.allclose(p, q, atol=1e-7)


def test_dirichlet_multinomial_log_prob_is_the_compound_pmf():
    """pyro.distributions.DirichletMultinomial.log_prob == scipy's Dirichlet-multinomial pmf (uses value.sum(-1) as the
    total count, as pyro does with validation off)."""
    from scipy.stats import dirichlet_multinomial

    a = torch.tensor([[0.7, 2.0, 1.3], [5.0, 0.2, 0.9]], dtype=torch.float64)
    x = torch.tensor([[3.0, 0.0, 5.0], [1.0, 1.0, 0.0]], dtype=torch.float64)
    got = dist.DirichletMultinomial(a, validate_args=False).log_prob(x)
    ref = [dirichlet_multinomial.logpmf(x[i].numpy().astype(int), a[i].numpy(), int(x[i].sum())) for i in range(2)]
    assert torch.allclose(got, torch.tensor(ref, dtype=torch.float64), atol=1e-12)


def test_factor_adds_its_log_factor_to_the_model_side_of_the_elbo():
    """pyro.factor(name, t) contributes exactly t (summed) to the model log-density and differentiates through it."""
    def model():
        w = pyro.param("w", torch.tensor(2.0))
        pyro.sample("z", dist.Normal(0.0, 1.0))
        pyro.factor("extra", -3.0 * w * w)

    def guide():
        pyro.sample("z", dist.Normal(0.0, 1.0))

    pyro.clear_param_store()
    torch.manual_seed(0)
    loss = pyro.infer.Trace_ELBO().differentiable_loss(model, guide)
    assert abs(float(loss) - 12.0) < 1e-12  # -(log p(z) - 3 w^2 - log q(z)) = 3 w^2
    loss.backward()
    assert abs(float(pyro.get_param_store().unconstrained("w").grad) - 12.0) < 1e-12
