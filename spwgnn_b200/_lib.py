"""Binds include/spwgnn.h to the nvcc-built libspwgnn.so that lives next to this file.

There is deliberately no fallback: if the CUDA library is missing or the machine has no GPU the
product raises.  (`python -c "import __graft_entry__ as g; g.build()"` builds it.)
"""
import os
import threading

from ._capi import CApi, EXPORTS, SpwError  # noqa: F401

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libspwgnn.so')
_api = None
_lock = threading.Lock()


def lib():
    global _api
    if _api is None:
        with _lock:
            if _api is None:
                if not os.path.exists(LIB_PATH):
                    raise SpwError(
                        'spwgnn_b200: %s not found. The CUDA extension is the only compute path; build it with '
                        '`python -c "import __graft_entry__ as g; g.build()"` (needs nvcc).' % LIB_PATH)
                _api = CApi(LIB_PATH)
    return _api


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise SpwError('spwgnn_b200 needs a CUDA device (B200 / sm_100a); there is no CPU path.')
