"""pyro.infer.SVI / Trace_ELBO restated for one particle and fully reparameterised guides."""
from __future__ import annotations

import torch

from . import poutine


class Trace_ELBO:
    def __init__(self, num_particles=1, **kw):
        if num_particles != 1:
            raise NotImplementedError

    def traces(self, model, guide, args, kwargs):
        guide_trace = poutine.trace(guide).get_trace(*args, **kwargs)
        model_trace = poutine.trace(poutine.replay(model, trace=guide_trace)).get_trace(*args, **kwargs)
        for site in guide_trace.nodes.values():
            if site["type"] == "sample" and not site["fn"].has_rsample:
                raise NotImplementedError("score-function terms: the reference's guides are fully reparameterised")
        model_trace.compute_log_prob()
        guide_trace.compute_log_prob()
        return model_trace, guide_trace

    def differentiable_loss(self, model, guide, *args, **kwargs):
        model_trace, guide_trace = self.traces(model, guide, args, kwargs)
        elbo = 0.0
        for site in model_trace.nodes.values():
            if site["type"] == "sample":
                elbo = elbo + site["log_prob_sum"]
        for site in guide_trace.nodes.values():
            if site["type"] == "sample":
                elbo = elbo - site["log_prob_sum"]
        self.last_traces = (model_trace, guide_trace)
        return -elbo


class SVI:
    def __init__(self, model, guide, optim, loss, **kw):
        self.model, self.guide, self.optim, self.loss = model, guide, optim, loss

    def step(self, *args, **kwargs):
        from . import get_param_store

        loss = self.loss.differentiable_loss(self.model, self.guide, *args, **kwargs)
        store = get_param_store()
        # parameters seen in either trace (pyro collects them from the param sites of the traces)
        model_trace, guide_trace = self.loss.last_traces
        names = []
        for tr in (model_trace, guide_trace):
            for site in tr.nodes.values():
                if site["type"] == "param" and site["name"] not in names:
                    names.append(site["name"])
        params = [store.unconstrained(n) for n in names]
        loss.backward()
        self.optim(params)
        for p in params:
            p.grad = None
        value = float(loss.detach())
        if value != value:
            import warnings

            warnings.warn("Encountered NaN loss")
        return value
