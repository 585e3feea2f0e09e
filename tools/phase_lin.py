"""Phase counters of csl::k_lin on one 150 -> 150 layer at the C2 edge count (development aid; needs tools/_build/libspwgnn_phase.so)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spwgnn_b200._lib as _lib
from spwgnn_b200._capi import CApi
_lib._api = CApi(os.path.join(os.path.dirname(os.path.abspath(__file__)), '_build', 'libspwgnn_phase.so'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_gpu_csl import to_csl, csl_linear
api = _lib.lib()
M = 368640
X = torch.randn(M, 152).cuda()
Xb = to_csl(X)
W = ((torch.rand(150, 150) * 2 - 1) * 0.15).cuda()
Yb = torch.empty(19 * M * 8, device='cuda')
bias = torch.zeros(150, device='cuda')
bits = torch.zeros(20 * M, dtype=torch.uint8, device='cuda')
for _ in range(2):
    csl_linear(api, M, Xb, 0, 150, W, 150, 160, Yb, bias=bias, act=1, ones_col=150, bits_out=bits)
