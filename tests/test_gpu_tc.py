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


def _tc_linear(api, M, X0, K0, X1, K1, W, N, NB, bias=None, rowscale=None, addend=None, act=0, mulsrc=None, mulmode=0,
               Y=None, ldy=None, accumulate=0, post_scale=1.0, ones_col=-1):
    ks = (K0 + K1 + 7) // 8
    scratch = torch.empty(2 * ks * 8 * NB, device='cuda')
    ptr = lambda t: 0 if t is None else t.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    api.check(api.dll.spw_tc_linear(M, X0.data_ptr(), X0.shape[1], K0, ptr(X1), 0 if X1 is None else X1.shape[1], K1,
                                    W.data_ptr(), N, NB, ptr(bias), ptr(rowscale), ptr(addend),
                                    0 if addend is None else addend.shape[1], act, ptr(mulsrc),
                                    0 if mulsrc is None else mulsrc.shape[1], mulmode, Y.data_ptr(), ldy, accumulate,
                                    post_scale, ones_col, scratch.data_ptr(), st))
    torch.cuda.synchronize()


@pytest.mark.parametrize('M', [1, 127, 128, 129, 1000, 40960, 80001])
def test_tc_linear_plain_shapes(M):
    """k_rows_tc: the four (K, N, NB) shapes the propagation network uses, no epilogue options."""
    from spwgnn_b200._lib import lib
    api = lib()
    g = torch.Generator().manual_seed(M)
    for (K, ldx, N, ldy, NB) in [(100, 100, 100, 100, 112), (150, 152, 100, 100, 112), (100, 100, 150, 152, 160),
                                 (150, 152, 150, 152, 160)]:
        X = torch.randn(M, ldx, generator=g).cuda()
        W = ((torch.rand(K, N, generator=g) * 2 - 1) * 0.2).cuda()
        Y = torch.full((M, ldy), float('nan'), device='cuda')
        _tc_linear(api, M, X, K, None, 0, W, N, NB, Y=Y, ldy=ldy)
        ref = X[:, :K].double() @ W.double()
        err = float((Y[:, :N].double() - ref).abs().max() / ref.abs().max())
        assert err < 2e-6, (K, N, NB, err)
        assert bool((Y[:, N:] == 0).all())


def test_tc_linear_epilogues():
    """k_rows_tc epilogue contract: bias*rowscale, addend, relu/tanh, the two multiplier modes, post-scale,
    accumulate, the ones column, two concatenated row segments."""
    from spwgnn_b200._lib import lib
    api = lib()
    g = torch.Generator().manual_seed(7)
    M = 777
    G = torch.randn(M, 100, generator=g).cuda()
    P = torch.randn(M, 100, generator=g).cuda()
    W = ((torch.rand(200, 100, generator=g) * 2 - 1) * 0.15).cuda()
    bias = torch.randn(101, generator=g).cuda()[1:]          # deliberately 4-byte aligned only
    rs = torch.randint(0, 3, (M,), generator=g).float().cuda()
    add = torch.randn(M, 100, generator=g).cuda()
    mul = (torch.rand(M, 100, generator=g) * 2 - 1).cuda()
    pre = torch.cat([G, P], 1).double() @ W.double() + rs.double()[:, None] * bias.double()[None] + add.double()
    # two segments, tanh, (1 - m^2) multiplier, post scale, accumulate
    Y0 = torch.randn(M, 100, generator=g).cuda()
    Y = Y0.clone()
    _tc_linear(api, M, G, 100, P, 100, W, 100, 112, bias=bias, rowscale=rs, addend=add, act=2, mulsrc=mul, mulmode=2, Y=Y,
               ldy=100, accumulate=1, post_scale=1.25)
    ref = torch.tanh(pre) * (1 - mul.double() ** 2) * 1.25 + Y0.double()
    assert float((Y.double() - ref).abs().max()) < 3e-6
    # relu + sign multiplier, no accumulate
    Y = torch.full((M, 100), float('nan'), device='cuda')
    _tc_linear(api, M, G, 100, P, 100, W, 100, 112, bias=bias, rowscale=rs, addend=add, act=1, mulsrc=mul, mulmode=1, Y=Y,
               ldy=100)
    ref = torch.relu(pre) * (mul.double() > 0)
    assert float((Y.double() - ref).abs().max()) < 5e-6
    # wide output with the ones column (relation-encoder layer)
    X = torch.randn(M, 152, generator=g).cuda()
    W2 = ((torch.rand(150, 150, generator=g) * 2 - 1) * 0.15).cuda()
    b2 = torch.randn(150, generator=g).cuda()
    Y = torch.full((M, 152), float('nan'), device='cuda')
    _tc_linear(api, M, X, 150, None, 0, W2, 150, 160, bias=b2, act=1, Y=Y, ldy=152, ones_col=150)
    ref = torch.relu(X[:, :150].double() @ W2.double() + b2.double())
    assert float((Y[:, :150].double() - ref).abs().max()) < 5e-6
    assert bool((Y[:, 150] == 1).all()) and bool((Y[:, 151] == 0).all())


def test_tc2_cta_pair_selftest_matches_fp64():
    """tcgen05 cta_group::2: two CTAs, half of the weight operand in each CTA's shared memory, M = 256."""
    from spwgnn_b200._lib import lib
    api = lib()
    g = torch.Generator().manual_seed(3)
    A = torch.randn(256, 152, generator=g, dtype=torch.float32)
    W = (torch.rand(150, 150, generator=g, dtype=torch.float32) * 2 - 1) * 0.14
    ref = (A[:, :150].double() @ W.double()).numpy()
    Ad, Wd = A.cuda(), W.cuda()
    D = torch.full((256, 160), float('nan'), device='cuda')
    scratch = torch.empty(4 * 19 * 8 * 80, device='cuda')
    status = torch.zeros(2, dtype=torch.int32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    api.check(api.dll.spw_tc2_selftest(Ad.data_ptr(), Wd.data_ptr(), 150, 150, D.data_ptr(), scratch.data_ptr(),
                                       status.data_ptr(), st))
    torch.cuda.synchronize()
    assert status.tolist() == [1, 1], 'MMA completion barrier timed out: %s' % status.tolist()
    got = D.cpu().numpy()
    err = np.abs(got[:, :150] - ref).max() / np.abs(ref).max()
    print('cta_group::2 3xTF32 GEMM rel err vs fp64: %.2e' % err)
    assert err < 2e-6
    assert np.all(got[:, 150:] == 0.0)
