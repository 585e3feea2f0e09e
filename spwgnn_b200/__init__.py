"""spwgnn_b200 -- B200-native (sm_100a) hot path of SPWGNN's tower-stability propagation network.

Layout: csrc/ (CUDA kernels + C ABI, built into libspwgnn.so), _capi/_lib (ctypes binding),
graph (packed tower batches, GPU edge builder), engine (forward/backward/Adam driver),
Networks (drop-in for the reference's Networks.py), dp (data-parallel sharding + all-reduce),
synth (seeded synthetic layouts).  Importing the package does not load CUDA; the first call does,
and raises if the CUDA library or a GPU is missing.
"""
from ._capi import PARAM_SPECS, SpwError  # noqa: F401

__version__ = '0.1.0'
