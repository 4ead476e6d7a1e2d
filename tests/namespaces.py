"""Plugin namespaces for cosmoslik_b200.synthetic.build_script: the product and the CPU oracle."""
import types

import cosmoslik_b200 as cb
from oracle import likes as olikes, runtime as ort

PRODUCT = types.SimpleNamespace(SlikPlugin=cb.SlikPlugin, param=cb.param, mspec_lnl=cb.likelihoods.mspec_lnl,
                                egfs=cb.models.egfs)
ORACLE = types.SimpleNamespace(SlikPlugin=ort.Plugin, param=ort.Param, mspec_lnl=olikes.MspecLike,
                               egfs=olikes.Egfs)
