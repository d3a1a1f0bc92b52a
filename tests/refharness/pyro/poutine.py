"""Handler stack: trace, replay, mask, plate (+ broadcast).  See pyro/__init__.py of this shim."""
from __future__ import annotations

from collections import OrderedDict, namedtuple

import torch

_STACK = []  # innermost handler LAST (pyro's _PYRO_STACK)


def _apply_stack(msg):
    """pyro.poutine.runtime.apply_stack: handlers see the message innermost-first on the way down
    (`_process_message`), the default behaviour runs unless a handler marked it done, then every handler that
    saw it post-processes on the way up."""
    seen = []
    for h in reversed(_STACK):
        seen.append(h)
        h.process(msg)
        if msg.get("stop"):
            break
    if not msg["done"]:
        msg["value"] = msg["default"]()
        msg["done"] = True
    for h in reversed(seen):
        h.postprocess(msg)
    return msg


class Messenger:
    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, *exc):
        assert _STACK[-1] is self
        _STACK.pop()
        return False

    def process(self, msg):
        pass

    def postprocess(self, msg):
        pass

    def __call__(self, fn):
        def wrapped(*a, **k):
            with self:
                return fn(*a, **k)

        return wrapped


# ---- trace ---------------------------------------------------------------------------------------
class Trace:
    def __init__(self):
        self.nodes = OrderedDict()

    def __contains__(self, k):
        return k in self.nodes

    def compute_log_prob(self):
        """pyro.poutine.Trace.compute_log_prob: log_prob -> scale_and_mask -> sum."""
        for site in self.nodes.values():
            if site["type"] != "sample":
                continue
            lp = site["fn"].log_prob(site["value"])
            mask = site["mask"]
            if mask is not None and mask is not True:
                if mask is False:
                    lp = torch.zeros_like(lp)
                else:
                    lp = torch.where(mask, lp, lp.new_zeros(()))
            if site["scale"] != 1.0:
                lp = lp * site["scale"]
            site["log_prob"] = lp
            site["log_prob_sum"] = lp.sum()


class trace(Messenger):
    def __init__(self, fn=None):
        self.fn = fn
        self.trace = Trace()

    def postprocess(self, msg):
        if msg["type"] == "sample" and msg["name"] in self.trace.nodes:
            raise RuntimeError(f"Multiple sample sites named '{msg['name']}'")
        self.trace.nodes[msg["name"]] = dict(msg)

    def get_trace(self, *args, **kwargs):
        self.trace = Trace()
        with self:
            self.fn(*args, **kwargs)
        return self.trace


class replay(Messenger):
    """pyro.poutine.ReplayMessenger: latent sites take the guide's value; observed sites are left alone."""

    def __init__(self, fn=None, trace=None):
        self.fn, self.guide_trace = fn, trace

    def process(self, msg):
        if msg["type"] != "sample" or self.guide_trace is None or msg["name"] not in self.guide_trace:
            return
        if msg["is_observed"]:
            return
        g = self.guide_trace.nodes[msg["name"]]
        if g["type"] != "sample" or g["is_observed"]:
            raise RuntimeError(f"site {msg['name']} must be a latent sample in the guide")
        msg["done"] = True
        msg["value"] = g["value"]
        msg["infer"] = g["infer"]

    def __call__(self, *args, **kwargs):
        with self:
            return self.fn(*args, **kwargs)


# ---- mask ----------------------------------------------------------------------------------------
class mask(Messenger):
    def __init__(self, fn=None, mask=None):
        if mask is None:
            raise ValueError("mask is required")
        self.fn, self.mask = fn, mask

    def process(self, msg):
        if msg["type"] != "sample":
            return
        msg["mask"] = self.mask if msg["mask"] is None else msg["mask"] & self.mask


# ---- plate ---------------------------------------------------------------------------------------
CondIndepStackFrame = namedtuple("CondIndepStackFrame", ["name", "dim", "size"])


class _DimAllocator:
    """pyro.poutine.runtime._DimAllocator: plates without `dim=` take the first free dim counting from -1."""

    def __init__(self):
        self._stack = []  # index i <-> dim -1-i ; value = plate name or None

    def allocate(self, name, dim):
        if name in self._stack:
            raise ValueError(f"duplicate plate '{name}'")
        if dim is None:
            dim = -1
            while -dim <= len(self._stack) and self._stack[-1 - dim] is not None:
                dim -= 1
        elif dim >= 0:
            raise ValueError("plate dim must be negative")
        while len(self._stack) < -dim:
            self._stack.append(None)
        if self._stack[-1 - dim] is not None:
            raise ValueError(f"plates '{name}' and '{self._stack[-1 - dim]}' collide at dim={dim}")
        self._stack[-1 - dim] = name
        return dim

    def free(self, name, dim):
        assert self._stack[-1 - dim] == name
        self._stack[-1 - dim] = None
        while self._stack and self._stack[-1] is None:
            self._stack.pop()


_DIM_ALLOCATOR = _DimAllocator()


class _PlateMessenger(Messenger):
    def __init__(self, name, size, dim):
        self.name, self.size, self.requested_dim = name, size, dim
        self.dim = None
        self.indices = torch.arange(size) if size is not None else None

    def __enter__(self):
        self.dim = _DIM_ALLOCATOR.allocate(self.name, self.requested_dim)
        super().__enter__()
        return self.indices

    def __exit__(self, *exc):
        _DIM_ALLOCATOR.free(self.name, self.dim)
        return super().__exit__(*exc)

    def process(self, msg):
        if msg["type"] != "sample":
            return
        frame = CondIndepStackFrame(self.name, self.dim, self.size)
        msg["cond_indep_stack"] = (frame,) + tuple(msg["cond_indep_stack"])
        # BroadcastMessenger._pyro_sample (plate applies it at every sample site)
        dist = msg["fn"]
        actual = tuple(dist.batch_shape)
        target = [None if s == 1 else s for s in actual]
        for f in msg["cond_indep_stack"]:
            if f.dim is None or f.size == -1:
                continue
            target = [None] * (-f.dim - len(target)) + target
            if target[f.dim] is not None and target[f.dim] != f.size:
                raise ValueError(
                    f"Shape mismatch inside plate('{f.name}') at site {msg['name']} dim {f.dim}, {f.size} vs {target[f.dim]}")
            target[f.dim] = f.size
        for i in range(-len(target) + 1, 1):
            if target[i] is None:
                target[i] = actual[i] if len(actual) >= -i else 1
        msg["fn"] = dist.expand(torch.Size(target))
