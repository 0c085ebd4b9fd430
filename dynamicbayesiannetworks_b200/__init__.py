"""B200-native RJ-MCMC proposal-scoring path for dyban-style dynamic Bayesian networks (h / nh / seq / glob DBN).

`Network` mirrors `dyban.Network`; `DbnEngine` is the thin ctypes layer over the C ABI in include/dbn_b200.h.
Importing the package does not need a GPU; creating an engine does (there is no CPU fallback).
"""
from .network import Network  # noqa: F401
from .engine import DbnEngine, DbnError  # noqa: F401
from .methods import METHOD_IDS, run_params  # noqa: F401

__all__ = ['Network', 'DbnEngine', 'DbnError', 'METHOD_IDS', 'run_params']
