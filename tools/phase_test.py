import sys, torch, ctypes
sys.path.insert(0, '/root/repo')
from spwgnn_b200._capi import CApi
api = CApi('/root/repo/tools/_build/libspwgnn_phase.so')
import importlib
sys.path.insert(0, '/root/repo/tests')
from test_gpu_tc import _tc_linear
g = torch.Generator().manual_seed(0)
for (M, K, ldx, N, ldy, NB, tag) in [(368640, 150, 152, 150, 152, 160, 'E-level'), (40960, 100, 100, 100, 100, 112, 'node')]:
    X = torch.randn(M, ldx, generator=g).cuda()
    W = ((torch.rand(K, N, generator=g) * 2 - 1) * 0.2).cuda()
    b = torch.randn(N, generator=g).cuda()
    Y = torch.empty(M, ldy, device='cuda')
    print(tag, flush=True)
    for it in range(2):
        _tc_linear(api, M, X, K, None, 0, W, N, NB, bias=b, act=1, Y=Y, ldy=ldy, ones_col=150 if N == 150 else -1)
    torch.cuda.synchronize()
