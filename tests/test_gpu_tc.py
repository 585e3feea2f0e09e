"""tcgen05 / TMEM building blocks (3xTF32 split GEMM) against an fp64 product."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tc_selftest_matches_fp64():
    from spwgnn_b200._lib import lib
    api = lib()
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, 152, generator=g, dtype=torch.float32)
    W = (torch.rand(150, 150, generator=g, dtype=torch.float32) * 2 - 1) * 0.14
    ref = (A[:, :150].double() @ W.double()).numpy()
    Ad, Wd = A.cuda(), W.cuda()
    D = torch.full((128, 160), float('nan'), device='cuda')
    scratch = torch.empty(2 * 24320, device='cuda')
    status = torch.zeros(1, dtype=torch.int32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    api.check(api.dll.spw_tc_selftest(Ad.data_ptr(), Wd.data_ptr(), 150, 150, D.data_ptr(), scratch.data_ptr(),
                                      status.data_ptr(), st))
    torch.cuda.synchronize()
    assert int(status[0]) == 1, 'MMA completion barrier timed out'
    got = D.cpu().numpy()
    err = np.abs(got[:, :150] - ref).max() / np.abs(ref).max()
    print('3xTF32 tcgen05 GEMM rel err vs fp64: %.2e' % err)
    assert err < 2e-6
    assert np.all(got[:, 150:] == 0.0)      # zero-padded weight columns
    # plain fp32 FFMA-order product for comparison of magnitudes
    e32 = np.abs((A[:, :150] @ W).numpy() - ref).max() / np.abs(ref).max()
    print('torch fp32 matmul rel err vs fp64: %.2e' % e32)
