"""Kernel logic under the host emulator (tools/cuemu): the SAME .cu sources compiled with g++,
one pthread per CUDA thread, checked against the oracle.  CPU-only CI for indexing / barrier /
tile-boundary logic; the GPU parity tests (-m gpu) remain the proof for the real binary."""
import numpy as np
import pytest
import torch

from oracle import propnet as O


@pytest.fixture(scope='module')
def emu():
    from emu_harness import Emu
    return Emu()


def _towers(rng, sizes):
    node_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    n = int(node_off[-1])
    pos = np.zeros((n, 2))
    pos[:, 0] = rng.integers(300, 900, n)
    pos[:, 1] = 110 + 80 * rng.integers(0, 5, n)
    wid = rng.integers(50, 300, n).astype(np.float64)
    return node_off, pos, wid


@pytest.mark.parametrize('fc', [False, True])
def test_emu_edges_bit_exact(emu, fc):
    rng = np.random.default_rng(0)
    node_off, pos, wid = _towers(rng, [3, 7, 1, 10, 0, 5, 12, 64, 2])
    pos[3] = pos[4]                                   # coincident blocks
    pos[5] = pos[6] + np.array([170.0, 0.0])          # exactly on the threshold
    g = emu.edges(pos, node_off, fully_connected=fc)
    eo, snd, rcv, slot = O.edge_list(pos, node_off, fully_connected=fc)
    assert np.array_equal(eo, g.edge_off)
    assert np.array_equal(snd, g.snd) and np.array_equal(rcv, g.rcv) and np.array_equal(slot, g.slot)
    order = np.lexsort((np.arange(g.E), rcv))
    assert np.array_equal(g.in_snd, snd[order]) and np.array_equal(g.in_rcv, rcv[order])
    assert np.array_equal(g.out_pos[order], np.arange(g.E))
    assert np.array_equal(g.deg_in, np.bincount(rcv, minlength=g.n))
    assert np.array_equal(g.deg_out, np.bincount(snd, minlength=g.n))


def test_emu_forward_backward_match_oracle(emu):
    rng = np.random.default_rng(1)
    node_off, pos, wid = _towers(rng, [12, 9, 14, 2, 11, 1, 13])      # several edge tiles, split segments
    g = emu.edges(pos, node_off, thr=330.0)
    assert g.E > 2 * 128 and g.E < sum(s * (s - 1) for s in [12, 9, 14, 2, 11, 1, 13])
    eo, snd, rcv, slot = O.edge_list(pos, node_off, thr=330.0)
    obj = np.concatenate([pos, wid[:, None]], 1) / 170.0
    w64 = O.init_weights(5, nonzero_bias=True)
    wnp = {k: v.numpy() for k, v in w64.items()}
    logits, probs = emu.forward(wnp, g, obj, training=True)
    o64 = torch.as_tensor(obj.astype(np.float32).astype(np.float64))
    p64, l64 = O.forward_sparse(w64, o64, torch.as_tensor(snd), torch.as_tensor(rcv), return_logits=True)
    assert np.abs(logits - l64.numpy()).max() / np.abs(l64.numpy()).max() < 1e-5
    assert np.abs(probs - p64.numpy()).max() < 1e-5
    tgt = (rng.random(g.n) > 0.5).astype(np.float32)
    dl, stats = emu.bce_grad(logits, tgt, g.n)
    loss64, _, _, g64 = O.loss_and_grads_sparse(w64, o64, torch.as_tensor(snd), torch.as_tensor(rcv),
                                                torch.as_tensor(tgt.astype(np.float64)))
    assert abs(stats[0] / g.n - float(loss64)) < 1e-5
    grads = emu.backward(g, dl)
    for k in O.tensor_names():
        ref = g64[k].numpy()
        assert np.abs(grads[k] - ref).max() / max(np.abs(ref).max(), 1e-30) < 1e-5, k
