"""pyro.distributions surface used by the reference: torch's distributions + DirichletMultinomial."""
import torch
from torch.distributions import Dirichlet, Laplace, LogNormal, Multinomial, Normal  # noqa: F401
from torch.distributions import constraints  # noqa: F401


def _log_beta_1(alpha, value):
    # pyro.distributions.conjugate._log_beta_1 (is_sparse=False branch)
    return torch.lgamma(1 + value) + torch.lgamma(alpha) - torch.lgamma(value + alpha)


class DirichletMultinomial(torch.distributions.Distribution):
    """pyro.distributions.DirichletMultinomial (pyro/distributions/conjugate.py), log_prob only: the reference
    uses it for observed sites exclusively.  `total_count` stays at its default 1 and is not used by log_prob."""

    arg_constraints = {}
    has_rsample = False

    def __init__(self, concentration, total_count=1, is_sparse=False, validate_args=None):
        self.concentration = concentration
        self.total_count = total_count
        self.is_sparse = is_sparse
        super().__init__(concentration.shape[:-1], concentration.shape[-1:], validate_args=False)

    def expand(self, batch_shape, _instance=None):
        batch_shape = torch.Size(batch_shape)
        return DirichletMultinomial(self.concentration.expand(batch_shape + self.event_shape), self.total_count, self.is_sparse)

    def log_prob(self, value):
        alpha = self.concentration
        if self.is_sparse:
            raise NotImplementedError
        return _log_beta_1(alpha.sum(-1), value.sum(-1)) - _log_beta_1(alpha, value).sum(-1)
