from .priors import priors
from .gaussian import gaussian
from .bandpower import mspec_lnl
from .spt_lowl import spt_lowl, spt_lowl_egfs
from .sptsz_lowl import SPTSZ_lowl
