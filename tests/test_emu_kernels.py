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


def test_emu_dropout_matches_oracle_with_same_mask(emu):
    """Training-time inverted dropout (Networks.py:77-78): the kernels' hash mask is restated in numpy and
    injected into the fp64 oracle; forward and all gradients must then agree."""
    rng = np.random.default_rng(2)
    node_off, pos, wid = _towers(rng, [9, 7, 10, 3])
    g = emu.edges(pos, node_off, thr=330.0)
    eo, snd, rcv, slot = O.edge_list(pos, node_off, thr=330.0)
    obj = np.concatenate([pos, wid[:, None]], 1) / 170.0
    w64 = O.init_weights(4, nonzero_bias=True)     # small graph: a relu-kink flip is improbable
    wnp = {k: v.numpy() for k, v in w64.items()}
    rate, seed = 0.1, 0x1234567855AA77
    l0, _ = emu.forward(wnp, g, obj, training=True)
    l0 = l0.copy()
    logits, _ = emu.forward(wnp, g, obj, training=True, dropout_rate=rate, dropout_seed=seed)   # state kept for backward
    assert np.abs(logits - l0).max() > 0
    sc, sq = O.dropout_seeds(seed)
    keep = 1.0 / (1.0 - np.float32(rate))
    epos = g.out_pos.astype(np.uint64)                      # receiver-major position of slot-order edge e
    c_keep = O.dropout_keep_mask(sc, epos[:, None] * np.uint64(160) + np.arange(150, dtype=np.uint64)[None, :], rate)
    q_keep = O.dropout_keep_mask(sq, np.arange(g.n, dtype=np.uint64)[:, None] * np.uint64(128) + np.arange(100, dtype=np.uint64)[None, :], rate)
    assert 0.05 < 1 - c_keep.mean() < 0.15 and 0.05 < 1 - q_keep.mean() < 0.15
    cs, qs = torch.as_tensor(c_keep * float(keep)), torch.as_tensor(q_keep * float(keep))
    o64 = torch.as_tensor(obj.astype(np.float32).astype(np.float64))
    tgt = (rng.random(g.n) > 0.5).astype(np.float32)
    loss64, _, l64, g64 = O.loss_and_grads_sparse(w64, o64, torch.as_tensor(snd), torch.as_tensor(rcv),
                                                  torch.as_tensor(tgt.astype(np.float64)), c_scale=cs, q_scale=qs)
    assert np.abs(logits - l64.numpy()).max() / np.abs(l64.numpy()).max() < 1e-5
    dl, _ = emu.bce_grad(logits, tgt, g.n)
    grads = emu.backward(g, dl)
    for k in O.tensor_names():
        ref = g64[k].numpy()
        assert np.abs(grads[k] - ref).max() / max(np.abs(ref).max(), 1e-30) < 2e-5, k


def test_emu_device_sampler_matches_host_restatement(emu):
    """the device-side Jenga layout sampler equals synth.g_jenga_ctr bit for bit, and draws
    the same kind of layouts as the reference sampler restatement (synth.g_jenga)."""
    from spwgnn_b200 import synth
    seed, T, lo, hi = 0x5EED1234ABCD, 300, 2, 40
    node_off, raw, obj, pos = emu.sample_jenga(seed, T, lo, hi)
    sizes = synth.sizes_ctr(seed, T, lo, hi)
    assert np.array_equal(np.diff(node_off), sizes) and node_off[0] == 0
    ref = np.concatenate([synth.g_jenga_ctr(int(n), seed, t) for t, n in enumerate(sizes)])
    assert np.array_equal(raw, ref)                                  # float64, bit for bit
    assert np.array_equal(obj, (ref / 170.0).astype(np.float32))     # main.py:91
    assert np.array_equal(pos, ref[:, :2])
    _, _, _, pos_glue = emu.sample_jenga(seed, T, lo, hi, inference_glue=True)
    assert np.array_equal(pos_glue, ref[:, :2] / 170.0)              # JengaBuilder.py:309-323
    # same layout family as the reference sampler: layer heights, width range, first layer inside [400, 1100]
    assert set(np.unique((ref[:, 1] - 110.0) % 80.0)) == {0.0}
    assert ref[:, 2].min() >= 50 and ref[:, 2].max() <= 300
    first = ref[node_off[:-1]]
    assert (first[:, 0] >= 400).all() and (first[:, 0] <= 1100).all() and (first[:, 1] == 110.0).all()


def test_emu_device_tower_sampler_matches_host_restatement(emu):
    """TowerCreator layouts on the device (k_sample_tower) == synth.g_tower_ctr bit for bit; the dropped block is object 0
    and sits on top (TowerCreator.py:265-271, 451)."""
    from spwgnn_b200 import synth
    seed, T, lo, hi = 987654321, 400, 2, 33
    node_off, raw, obj, pos = emu.sample_jenga(seed, T, lo, hi, kind='tower')
    sizes = synth.sizes_ctr(seed, T, lo, hi)
    assert np.array_equal(np.diff(node_off), sizes)
    ref = np.concatenate([synth.g_tower_ctr(int(n), seed, t) for t, n in enumerate(sizes)])
    assert np.array_equal(raw, ref)
    assert np.array_equal(obj, (ref / 170.0).astype(np.float32))
    for t in range(T):
        tw = raw[node_off[t]:node_off[t + 1]]
        assert tw[0, 1] == tw[:, 1].max() and (tw[:, 2] == 150.0).all()


def test_emu_edges_property_sweep(emu):
    """Hypothesis-driven sweep of the edge builder (the reference has no tests; the property is
    equality with the numpy restatement of main.py:66-81): ragged tower sizes including 0 and 1, positions on a coarse
    lattice so that distances land exactly on / one ulp around the 170 threshold, arbitrary thresholds."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    lattice = st.integers(0, 12).map(lambda k: 85.0 * k)                    # multiples of 85: many pairs at exactly 170
    nudge = st.sampled_from([0.0, 0.0, 0.0, 2.0 ** -44, -2.0 ** -44, 0.5])   # one-ulp neighbours of the lattice

    @st.composite
    def batch(draw):
        sizes = draw(st.lists(st.integers(0, 9), min_size=1, max_size=6))
        n = sum(sizes)
        xs = [draw(lattice) + draw(nudge) for _ in range(n)]
        ys = [draw(lattice) + draw(nudge) for _ in range(n)]
        thr = draw(st.sampled_from([170.0, 170.0, 85.0, 1.0, 240.41630560342617]))   # last: 170 * sqrt(2)
        return sizes, np.array([xs, ys], dtype=np.float64).T.reshape(n, 2), thr

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(batch())
    def check(b):
        sizes, pos, thr = b
        node_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        g = emu.edges(pos, node_off, thr=thr)
        eo, snd, rcv, slot = O.edge_list(pos, node_off, thr=thr)
        assert np.array_equal(eo, g.edge_off)
        assert np.array_equal(snd, g.snd) and np.array_equal(rcv, g.rcv) and np.array_equal(slot, g.slot)
        if g.E:
            order = np.lexsort((np.arange(g.E), rcv))
            assert np.array_equal(g.in_snd, snd[order]) and np.array_equal(g.in_rcv, rcv[order])
            assert np.array_equal(g.out_pos[order], np.arange(g.E))

    check()
