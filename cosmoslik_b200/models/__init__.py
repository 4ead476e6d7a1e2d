from .egfs import egfs, egfs_specs
from .clust_poisson_egfs import clust_poisson_egfs
