from .metropolis_hastings import metropolis_hastings
from .emcee import emcee
from .utils import initialize_covariance, proposal_factor
