"""Round-2 building block: the pipelined column-slab linear layer (csl::k_lin, spw_csl_linear) against an fp64 product.
Arrays are stored column-slab major ([C / 4][rows][4]); sign bits byte-slab major ([C / 8][rows] u8)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def to_csl(X, cols=None):
    """(M, C) row-major -> flat CSL buffer with ceil(cols / 8) * 2 quad slabs of M rows (zero padded to a multiple of 8 columns)."""
    M, C = X.shape
    cols = cols or C
    c4 = 2 * ((cols + 7) // 8)
    P = torch.zeros(M, c4 * 4, dtype=X.dtype, device=X.device)
    P[:, :C] = X
    return P.reshape(M, c4, 4).permute(1, 0, 2).contiguous().reshape(-1)


def from_csl(buf, M, cols):
    c4 = buf.numel() // (M * 4)
    return buf.reshape(c4, M, 4).permute(1, 0, 2).reshape(M, c4 * 4)[:, :cols]


def bits_to_bool(bits_u8, M, cols):
    """byte-slab sign bits [cols / 8][M] -> (M, cols) bool."""
    b = bits_u8.reshape(-1, M).cpu().numpy()
    out = np.zeros((M, cols), dtype=bool)
    for c in range(cols):
        out[:, c] = (b[c >> 3] >> (c & 7)) & 1
    return out


def bool_to_bits(mask):
    M, cols = mask.shape
    b = np.zeros(((cols + 7) // 8, M), dtype=np.uint8)
    for c in range(cols):
        b[c >> 3] |= (mask[:, c].astype(np.uint8) << (c & 7))
    return torch.as_tensor(b).reshape(-1)


def csl_linear(api, M, Xbuf, x_col0, K, W, N, NB, Ybuf, y_col0=0, bias=None, rowscale=None, addend=None, add_col0=0, act=0,
               mulsrc=None, mul_col0=0, mulmode=0, bits_in=None, bits_out=None, accumulate=0, post_scale=1.0, ones_col=-1,
               write_pad=1, sync=True, scratch=None):
    nks = (K + 7) // 8
    if scratch is None:
        scratch = torch.empty(2 * nks * 8 * NB, device='cuda')
    ptr = lambda t: 0 if t is None else t.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    slab = M * 4
    api.check(api.dll.spw_csl_linear(M, Xbuf.data_ptr(), slab, x_col0, K, W.data_ptr(), N, NB, ptr(bias), ptr(rowscale),
                                     ptr(addend), slab, add_col0, act, ptr(mulsrc), slab, mul_col0, mulmode, ptr(bits_in),
                                     ptr(bits_out), Ybuf.data_ptr(), slab, y_col0, accumulate, post_scale, ones_col, write_pad,
                                     scratch.data_ptr(), st))
    if sync:
        torch.cuda.synchronize()


@pytest.mark.parametrize('M', [1, 127, 128, 129, 1000, 40960, 80001])
def test_csl_linear_plain_shapes(M):
    from spwgnn_b200._lib import lib
    api = lib()
    g = torch.Generator().manual_seed(M)
    for (K, N, NB) in [(100, 100, 112), (150, 100, 112), (100, 150, 160), (150, 150, 160), (200, 100, 112)]:
        X = torch.randn(M, K, generator=g).cuda()
        W = ((torch.rand(K, N, generator=g) * 2 - 1) * 0.2).cuda()
        Xb = to_csl(X)
        Yb = torch.full((((N + 7) // 8) * M * 8,), float('nan'), device='cuda')
        csl_linear(api, M, Xb, 0, K, W, N, NB, Yb)
        Y = from_csl(Yb, M, ((N + 7) // 8) * 8)
        ref = X.double() @ W.double()
        err = float((Y[:, :N].double() - ref).abs().max() / ref.abs().max())
        assert err < 2e-6, (K, N, NB, err)
        assert bool((Y[:, N:] == 0).all())


def test_csl_linear_epilogues_and_views():
    """Epilogue contract (bias * rowscale, addend, relu / tanh, multiplier modes incl. sign bits, post-scale, accumulate, ones
    column, sign bits out) and column-offset views into a combined array ([g | p] is one 200-column array)."""
    from spwgnn_b200._lib import lib
    api = lib()
    g = torch.Generator().manual_seed(7)
    M = 777
    GP = torch.randn(M, 200, generator=g).cuda()
    GPb = to_csl(GP)
    W = ((torch.rand(200, 100, generator=g) * 2 - 1) * 0.15).cuda()
    bias = torch.randn(101, generator=g).cuda()[1:]          # deliberately 4-byte aligned only
    rs = torch.randint(0, 3, (M,), generator=g).float().cuda()
    add = torch.randn(M, 100, generator=g).cuda()
    mul = (torch.rand(M, 100, generator=g) * 2 - 1).cuda()
    addb, mulb = to_csl(add), to_csl(mul)

    # u = relu([g | p].W + add)      (the hidden layer of the object propagator)
    Yb = torch.full((13 * M * 8,), float('nan'), device='cuda')
    csl_linear(api, M, GPb, 0, 200, W, 100, 112, Yb, addend=addb, act=1)
    ref = torch.relu(GP.double() @ W.double() + add.double())
    Y = from_csl(Yb, M, 104)
    assert float((Y[:, :100].double() - ref).abs().max() / ref.abs().max()) < 2e-6
    assert bool((Y[:, 100:] == 0).all())

    # g = tanh(H.W + deg * bias)       (in-degree-scaled bias)
    H = torch.randn(M, 150, generator=g).cuda()
    Hb = to_csl(H, 152)
    Wg = ((torch.rand(150, 100, generator=g) * 2 - 1) * 0.1).cuda()
    Yb = torch.full((13 * M * 8,), float('nan'), device='cuda')
    csl_linear(api, M, Hb, 0, 150, Wg, 100, 112, Yb, bias=bias, rowscale=rs, act=2)
    ref = torch.tanh(H.double() @ Wg.double() + rs.double()[:, None] * bias.double())
    Y = from_csl(Yb, M, 104)
    assert float((Y[:, :100].double() - ref).abs().max()) < 2e-6

    # p-part view (columns 100..199) as the input; tanh output written INTO columns 100..199 of a second combined array,
    # with the residual addend read from the first one; the g part (columns 0..99, sharing a slab with column 100..103) stays
    W2 = ((torch.rand(100, 100, generator=g) * 2 - 1) * 0.15).cuda()
    OUT = torch.randn(M, 200, generator=g).cuda()
    OUTb = to_csl(OUT)
    csl_linear(api, M, GPb, 100, 100, W2, 100, 112, OUTb, y_col0=100, bias=bias, addend=GPb, add_col0=100, act=2, write_pad=0)
    ref = torch.tanh(GP[:, 100:].double() @ W2.double() + bias.double() + GP[:, 100:].double())
    got = from_csl(OUTb, M, 200)
    assert float((got[:, 100:].double() - ref).abs().max()) < 2e-6
    assert bool((got[:, :100] == OUT[:, :100]).all()), 'columns outside the output view were touched'

    # multiplier modes, post-scale, accumulate
    X = torch.randn(M, 100, generator=g).cuda()
    Xb = to_csl(X)
    Y0 = torch.randn(M, 100, generator=g).cuda()
    base = X.double() @ W2.double()
    Yb = to_csl(Y0)
    csl_linear(api, M, Xb, 0, 100, W2, 100, 112, Yb, mulsrc=mulb, mulmode=1, post_scale=1.0 / 0.9, accumulate=1)
    ref = Y0.double() + base * (mul.double() > 0).double() / 0.9
    assert float((from_csl(Yb, M, 100).double() - ref).abs().max() / ref.abs().max()) < 2e-6
    Yb = to_csl(Y0)
    csl_linear(api, M, Xb, 0, 100, W2, 100, 112, Yb, mulsrc=mulb, mulmode=2)
    ref = base * (1 - mul.double() ** 2)
    assert float((from_csl(Yb, M, 100).double() - ref).abs().max() / ref.abs().max()) < 2e-6
    # K = 150 -> 100: accumulate; addend + (1 - m^2)
    Wt = ((torch.rand(150, 100, generator=g) * 2 - 1) * 0.1).cuda()
    Yb = to_csl(Y0)
    csl_linear(api, M, Hb, 0, 150, Wt, 100, 112, Yb, accumulate=1)
    ref = Y0.double() + H.double() @ Wt.double()
    assert float((from_csl(Yb, M, 100).double() - ref).abs().max() / ref.abs().max()) < 2e-6
    Yb = torch.full((13 * M * 8,), float('nan'), device='cuda')
    csl_linear(api, M, Hb, 0, 150, Wt, 100, 112, Yb, addend=addb, mulsrc=mulb, mulmode=2)
    ref = (H.double() @ Wt.double() + add.double()) * (1 - mul.double() ** 2)
    assert float((from_csl(Yb, M, 100).double() - ref).abs().max() / ref.abs().max()) < 2e-6

    # 150-wide layer: relu, ones column, sign bits out; then a data-gradient style layer masked by those bits
    X = torch.randn(M, 150, generator=g).cuda()
    Xb = to_csl(X, 152)
    W3 = ((torch.rand(150, 150, generator=g) * 2 - 1) * 0.15).cuda()
    b3 = torch.randn(150, generator=g).cuda()
    Yb = torch.full((19 * M * 8,), float('nan'), device='cuda')
    bits = torch.zeros(20 * M, dtype=torch.uint8, device='cuda')
    csl_linear(api, M, Xb, 0, 150, W3, 150, 160, Yb, bias=b3, act=1, ones_col=150, bits_out=bits)
    ref = torch.relu(X.double() @ W3.double() + b3.double())
    Y = from_csl(Yb, M, 152)
    assert float((Y[:, :150].double() - ref).abs().max() / ref.abs().max()) < 2e-6
    assert bool((Y[:, 150] == 1).all()) and bool((Y[:, 151] == 0).all())
    got_bits = bits_to_bool(bits, M, 160)
    assert np.array_equal(got_bits[:, :150], (Y[:, :150] > 0).cpu().numpy())
    assert not got_bits[:, 150:].any()
    mask = np.random.default_rng(0).random((M, 150)) > 0.5
    mb = bool_to_bits(np.pad(mask, ((0, 0), (0, 10)))).cuda()
    Gb = torch.full((19 * M * 8,), float('nan'), device='cuda')
    csl_linear(api, M, Yb, 0, 150, W3, 150, 160, Gb, mulmode=3, bits_in=mb, post_scale=1.25)
    ref = (Y[:, :150].double() @ W3.double()) * torch.as_tensor(mask).cuda().double() * 1.25
    G = from_csl(Gb, M, 152)
    assert float((G[:, :150].double() - ref).abs().max() / ref.abs().max()) < 2e-6
    assert bool((G[:, 150:] == 0).all())


def test_csl_linear_timing_c2_edge_layer():
    """Not a parity test: prints the launch time of one 150 -> 150 relation-encoder layer at the C2 edge count."""
    from spwgnn_b200._lib import lib
    api = lib()
    M = 368640
    g = torch.Generator().manual_seed(1)
    X = torch.randn(M, 152, generator=g).cuda()
    Xb = to_csl(X)
    W = ((torch.rand(150, 150, generator=g) * 2 - 1) * 0.15).cuda()
    Yb = torch.empty(19 * M * 8, device='cuda')
    bias = torch.zeros(150, device='cuda')
    bits = torch.zeros(20 * M, dtype=torch.uint8, device='cuda')
    scratch = torch.empty(2 * 19 * 8 * 160, device='cuda')
    for _ in range(3):
        csl_linear(api, M, Xb, 0, 150, W, 150, 160, Yb, bias=bias, act=1, ones_col=150, bits_out=bits, scratch=scratch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        csl_linear(api, M, Xb, 0, 150, W, 150, 160, Yb, bias=bias, act=1, ones_col=150, bits_out=bits, sync=False, scratch=scratch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('k_lin 150->150 on %d rows: %.3f ms per launch (incl. weight packing), %.1f k cycles per tile at 1.965 GHz'
          % (M, ms, ms * 1e-3 * 1.965e9 / 20 / 1e3))
    ref = torch.relu(X[:4096, :150].double() @ W.double())
    Y = from_csl(Yb, M, 152)[:4096, :150]
    assert float((Y.double() - ref).abs().max() / ref.abs().max()) < 2e-6
