"""The 22 parameter tensors of the propagation network in one flat, 16-byte-aligned fp32 buffer.

Reference: the four MLPs created in /root/reference/src/Networks.py:46-50 through
Blocks.py:12-91 (Keras Dense: kernel[in, out] glorot_uniform, bias zeros).  One flat buffer
means one NCCL all-reduce for the gradients.
"""
import math

import torch

from ._capi import PARAM_SPECS, CApi

ALIGN = 64   # floats; every tensor starts on a 256-byte boundary


def _offsets():
    offs, off = [], 0
    for _, shape in PARAM_SPECS:
        offs.append(off)
        numel = 1
        for s in shape:
            numel *= s
        off = (off + numel + ALIGN - 1) // ALIGN * ALIGN
    return offs, off


OFFSETS, FLAT_SIZE = _offsets()
STATS_TAIL = 4   # floats behind the gradient buffer: the loss / accuracy sums ride in the gradient all-reduce (dp.py)
N_PARAMS = sum(int(torch.Size(s).numel()) for _, s in PARAM_SPECS)   # 209501


class ParamBuffer:
    """flat: (FLAT_SIZE,) fp32 tensor; views: name -> tensor view with the Keras shape."""

    def __init__(self, device, flat=None):
        self.flat = torch.zeros(FLAT_SIZE, dtype=torch.float32, device=device) if flat is None else flat
        assert self.flat.numel() == FLAT_SIZE and self.flat.dtype == torch.float32
        self.views = {}
        for (name, shape), off in zip(PARAM_SPECS, OFFSETS):
            numel = int(torch.Size(shape).numel())
            self.views[name] = self.flat[off:off + numel].view(*shape)

    def c_struct(self):
        base = self.flat.data_ptr()
        return CApi.params([base + 4 * off for off in OFFSETS])

    def names(self):
        return [n for n, _ in PARAM_SPECS]

    def load_dict(self, d):
        for name, _ in PARAM_SPECS:
            self.views[name].copy_(torch.as_tensor(d[name]).to(self.flat.device, torch.float32))
        return self

    def to_dict(self):
        return {k: v.detach().clone() for k, v in self.views.items()}

    def glorot_init(self, seed=0):
        """Keras defaults: glorot_uniform kernels, zero biases (Blocks.py:22-27)."""
        g = torch.Generator().manual_seed(seed)
        self.flat.zero_()
        for name, shape in PARAM_SPECS:
            if len(shape) == 2:
                lim = math.sqrt(6.0 / (shape[0] + shape[1]))
                k = (torch.rand(shape[0], shape[1], generator=g, dtype=torch.float64) * 2 - 1) * lim
                self.views[name].copy_(k.float())
        return self
