from torch.distributions.constraints import *  # noqa: F401,F403
from torch.distributions.constraints import positive, real, simplex, unit_interval  # noqa: F401
