"""Minimal restatement of the pyro-ppl primitives the reference's model code calls.

TEST INFRASTRUCTURE ONLY.  pyro-ppl (>=1.8.5, bean setup.py:42) is not installable in this image, so the
reference's OWN, UNMODIFIED model / guide programs (bean/model/model.py, survival_model.py, run.py) are
executed on top of this shim to produce the golden vectors under tests/golden/ (see
tests/golden/make_reference_golden.py).  Only the effect-handler semantics the reference relies on are
restated, from pyro's published behaviour:

  * `pyro.param`    -- global store of UNCONSTRAINED tensors; value = transform_to(constraint)(unconstrained)
  * `pyro.sample`   -- message passed through the handler stack (trace / replay / mask / plate)
  * `pyro.plate`    -- dim allocation (explicit `dim=` or first free dim from -1) and BroadcastMessenger
                       expansion of the site's distribution to the enclosing plates' sizes
  * `poutine.mask`  -- site mask; log_prob -> where(mask, log_prob, 0)
  * `Trace_ELBO` (1 particle, reparameterised sites), `SVI.step`, `ClippedAdam`  (pyro.infer / pyro.optim)
"""
from __future__ import annotations

import torch
from torch.distributions import biject_to, constraints, transform_to  # noqa: F401

from . import poutine  # noqa: F401
from .poutine import _apply_stack, _PlateMessenger


# ---- parameter store ---------------------------------------------------------------------------
class ParamStoreDict:
    def __init__(self):
        self._unconstrained = {}
        self._constraints = {}

    def clear(self):
        self._unconstrained.clear()
        self._constraints.clear()

    def items(self):
        return [(k, self[k]) for k in self._unconstrained]

    def keys(self):
        return self._unconstrained.keys()

    def __contains__(self, k):
        return k in self._unconstrained

    def __getitem__(self, k):
        return transform_to(self._constraints[k])(self._unconstrained[k])

    def unconstrained(self, k):
        return self._unconstrained[k]

    def setdefault_param(self, name, init, constraint):
        if name not in self._unconstrained:
            if init is None:
                raise RuntimeError(f"param {name} has no initial value")
            value = init() if callable(init) else init
            value = torch.as_tensor(value)
            with torch.no_grad():
                u = transform_to(constraint).inv(value.detach()).clone()
            u.requires_grad_(True)
            self._unconstrained[name] = u
            self._constraints[name] = constraint
        return self[name]


_PARAM_STORE = ParamStoreDict()


def get_param_store():
    return _PARAM_STORE


def clear_param_store():
    _PARAM_STORE.clear()


def set_rng_seed(seed):
    import random

    import numpy as np

    torch.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)


def param(name, init_tensor=None, constraint=constraints.real, event_dim=None):
    msg = {"type": "param", "name": name, "fn": None, "value": None, "done": False, "is_observed": False,
           "mask": None, "scale": 1.0, "cond_indep_stack": (), "infer": {}}

    def default():
        return _PARAM_STORE.setdefault_param(name, init_tensor, constraint)

    msg["default"] = default
    _apply_stack(msg)
    return msg["value"]


def sample(name, fn, obs=None, infer=None):
    msg = {"type": "sample", "name": name, "fn": fn, "value": obs, "done": False, "is_observed": obs is not None,
           "mask": None, "scale": 1.0, "cond_indep_stack": (), "infer": infer or {}}

    def default():
        if msg["is_observed"]:
            return msg["value"]
        f = msg["fn"]
        return f.rsample() if f.has_rsample else f.sample()  # TorchDistributionMixin.__call__

    msg["default"] = default
    _apply_stack(msg)
    return msg["value"]


class _Unit:
    """pyro.distributions.Unit: a trivial distribution over the empty tensor whose log_prob IS the given log factor."""

    has_rsample = False

    def __init__(self, log_factor):
        self.log_factor = log_factor
        self.batch_shape, self.event_shape = tuple(log_factor.shape), (0,)

    def expand(self, batch_shape):
        return _Unit(self.log_factor.expand(tuple(batch_shape)))

    def sample(self, sample_shape=()):
        return self.log_factor.new_empty(tuple(self.log_factor.shape) + (0,))

    def log_prob(self, value):
        return self.log_factor


def factor(name, log_factor, *, has_rsample=None):
    """pyro.factor: adds an arbitrary log-probability term to the model (`sample(name, Unit(log_factor), obs=empty)` in
    pyro.primitives); plates and masks apply to it like to any other site."""
    import torch

    log_factor = torch.as_tensor(log_factor)
    unit = _Unit(log_factor)
    sample(name, unit, obs=unit.sample(), infer={"is_auxiliary": True})


def plate(name, size=None, subsample_size=None, subsample=None, dim=None):
    if subsample_size is not None or subsample is not None:
        raise NotImplementedError("subsampling plates are not used by the reference")
    return _PlateMessenger(name, size, dim)


from . import distributions, infer, optim  # noqa: E402,F401
