from .metropolis_hastings import metropolis_hastings
from .emcee import emcee
from .priors import priors
from .polychord import polychord
from .utils import initialize_covariance, proposal_factor
