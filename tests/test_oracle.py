"""The oracle against the fixtures produced by running the reference's own source
(oracle/make_golden.py), plus self-consistency checks."""
import os

import numpy as np
import pytest
import torch

from oracle import propnet as O

TRAIN = ['train_n2', 'train_n4', 'train_n7', 'train_n9']
PRED = ['predict_jenga_n5', 'predict_jenga_n9']


def _weights(golden_dir, dtype=torch.float64):
    z = np.load(os.path.join(golden_dir, 'weights.npz'))
    return {k: torch.as_tensor(z[k]).to(dtype) for k in z.files}


def _flat_batch(objects):
    B, N, _ = objects.shape
    return objects.reshape(B * N, 3), np.arange(B + 1) * N


@pytest.mark.parametrize('name', TRAIN)
def test_relations_match_reference_loops(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    rs, rr = O.build_relations_dense(g['raw_pos'])
    assert np.array_equal(rs, g['sender_relations'])
    assert np.array_equal(rr, g['receiver_relations'])
    # sparse edge list expands to the same one-hots, slot for slot
    B, N, _ = g['raw_pos'].shape
    node_off = np.arange(B + 1) * N
    eo, snd, rcv, slot = O.edge_list(g['raw_pos'].reshape(-1, 2), node_off)
    rs2, rr2 = O.relations_from_edges(N, eo, snd, rcv, slot, node_off)
    assert np.array_equal(rs2, g['sender_relations'])
    assert np.array_equal(rr2, g['receiver_relations'])
    # closed-form slot formula
    for e in range(len(snd)):
        t = np.searchsorted(eo, e, side='right') - 1
        assert slot[e] == O.slot_of(snd[e] - node_off[t], rcv[e] - node_off[t], N)


@pytest.mark.parametrize('name', TRAIN)
def test_normalisation_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    assert np.array_equal(O.normalise_objects(g['raw_pos']), g['objects'][:, :, 0:2])


@pytest.mark.parametrize('name', TRAIN + PRED)
def test_forward_dense_and_sparse_match_reference_graph(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    w = _weights(golden_dir)
    obj = torch.as_tensor(g['objects'])
    rs = torch.as_tensor(g['sender_relations']); rr = torch.as_tensor(g['receiver_relations'])
    probs = O.forward_dense(w, obj, rs, rr)
    assert probs.shape == g['probs'].shape
    assert np.abs(probs.numpy() - g['probs']).max() < 1e-13
    # sparse restatement on the edge list extracted from the one-hots
    B, N, R = g['sender_relations'].shape
    snd, rcv = [], []
    for b in range(B):
        for r in range(R):
            if g['sender_relations'][b, :, r].any():
                snd.append(b * N + int(np.argmax(g['sender_relations'][b, :, r])))
                rcv.append(b * N + int(np.argmax(g['receiver_relations'][b, :, r])))
    ps = O.forward_sparse(w, obj.reshape(B * N, 3), torch.tensor(snd, dtype=torch.long), torch.tensor(rcv, dtype=torch.long))
    assert np.abs(ps.numpy().reshape(B, N, 1) - g['probs']).max() < 1e-13


def test_inference_glue_is_fully_connected(golden_dir):
    # reference quirk F5: normalised positions thresholded against 170 => every slot active
    for name in PRED:
        g = np.load(os.path.join(golden_dir, name + '.npz'))
        assert g['sender_relations'].sum() == g['sender_relations'].shape[2]
        assert np.array_equal(g['objects'][0], g['raw'] / 170.0)
        rs, rr = O.build_relations_dense(g['objects'][:, :, 0:2], 170.0)
        assert np.array_equal(rs, g['sender_relations']) and np.array_equal(rr, g['receiver_relations'])


@pytest.mark.parametrize('name', ['train_n2', 'train_n4', 'train_n7'])
def test_gradients_match_reference_graph(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    w = _weights(golden_dir)
    B, N, _ = g['objects'].shape
    node_off = np.arange(B + 1) * N
    eo, snd, rcv, slot = O.edge_list(g['raw_pos'].reshape(-1, 2), node_off)
    obj = torch.as_tensor(g['objects']).reshape(B * N, 3)
    tgt = torch.as_tensor(g['target']).reshape(B * N)
    loss, probs, logits, grads = O.loss_and_grads_sparse(w, obj, torch.as_tensor(snd), torch.as_tensor(rcv), tgt)
    assert abs(float(loss) - float(g['loss'])) < 1e-13
    for k in O.tensor_names():
        ref = g['g:' + k].astype(np.float64)
        scale = max(np.abs(ref).max(), 1e-30)
        assert np.abs(grads[k].numpy() - ref).max() <= 2e-7 * scale + 1e-12, k   # fixture stored as fp32


def test_param_count_and_names():
    assert O.N_PARAMS == 209501
    assert len(O.tensor_names()) == 22


def test_dense_equals_sparse_fp64_random():
    torch.manual_seed(0)
    rng = np.random.default_rng(3)
    w = O.init_weights(1, nonzero_bias=True)
    B, N = 3, 6
    pos = rng.uniform(300, 900, size=(B, N, 2))
    pos[:, :, 1] = 110 + 80 * rng.integers(0, 3, size=(B, N))
    wid = rng.integers(50, 300, size=(B, N, 1)).astype(np.float64)
    objects = np.concatenate([pos, wid], axis=2) / 170.0
    rs, rr = O.build_relations_dense(pos)
    assert 0 < rs.sum() < B * N * (N - 1)
    pd = O.forward_dense(w, torch.as_tensor(objects), torch.as_tensor(rs), torch.as_tensor(rr))
    node_off = np.arange(B + 1) * N
    eo, snd, rcv, slot = O.edge_list(pos.reshape(-1, 2), node_off)
    ps = O.forward_sparse(w, torch.as_tensor(objects).reshape(-1, 3), torch.as_tensor(snd), torch.as_tensor(rcv))
    assert np.abs(pd.numpy().reshape(-1) - ps.numpy()).max() < 1e-14


def test_fp32_oracle_close_to_fp64():
    w64 = O.init_weights(2, nonzero_bias=True)
    w32 = {k: v.float() for k, v in w64.items()}
    rng = np.random.default_rng(5)
    N = 8
    pos = rng.uniform(0, 3, size=(N, 2)); wid = rng.uniform(0.3, 1.7, size=(N, 1))
    obj = torch.as_tensor(np.concatenate([pos, wid], 1))
    m, j = np.nonzero(~np.eye(N, dtype=bool))
    snd, rcv = torch.as_tensor(m), torch.as_tensor(j)
    _, l64 = O.forward_sparse(w64, obj, snd, rcv, return_logits=True)
    _, l32 = O.forward_sparse(w32, obj.float(), snd, rcv, return_logits=True)
    assert (l32.double() - l64).abs().max() / l64.abs().max() < 5e-6


def test_hand_cases():
    w = O.init_weights(4, nonzero_bias=True)
    # isolated nodes: g = tanh(0) = 0, output depends only on own [y, w]
    obj = torch.tensor([[0.1, 0.5, 0.9], [7.0, 0.5, 0.9]], dtype=torch.float64)
    e = torch.zeros(0, dtype=torch.long)
    p = O.forward_sparse(w, obj, e, e)
    assert abs(float(p[0] - p[1])) < 1e-15
    # permutation equivariance: relabelling blocks permutes outputs
    rng = np.random.default_rng(9)
    N = 5
    obj = torch.as_tensor(rng.uniform(0, 2, size=(N, 3)))
    m, j = np.nonzero(~np.eye(N, dtype=bool))
    p0 = O.forward_sparse(w, obj, torch.as_tensor(m), torch.as_tensor(j))
    perm = rng.permutation(N)
    p1 = O.forward_sparse(w, obj[perm], torch.as_tensor(m), torch.as_tensor(j))
    assert (p0[perm] - p1).abs().max() < 1e-12


def test_gradcheck_small():
    # autograd of the fp64 oracle agrees with finite differences on a few weight entries
    w = O.init_weights(6, nonzero_bias=True)
    rng = np.random.default_rng(1)
    N = 4
    obj = torch.as_tensor(rng.uniform(0, 2, size=(N, 3)))
    m, j = np.nonzero(~np.eye(N, dtype=bool))
    snd, rcv = torch.as_tensor(m), torch.as_tensor(j)
    tgt = torch.tensor([1.0, 0.0, 1.0, 1.0], dtype=torch.float64)
    loss, _, _, grads = O.loss_and_grads_sparse(w, obj, snd, rcv, tgt)
    for k, idx in [('rmp.w0', (200, 7)), ('rm.w0', (1, 3)), ('omp.w1', (5, 0)), ('om.b1', (9,)), ('rmp.b2', (50,))]:
        h = 1e-6
        wp = {a: b.clone() for a, b in w.items()}; wp[k][idx] += h
        wm = {a: b.clone() for a, b in w.items()}; wm[k][idx] -= h
        lp = O.bce_keras(O.forward_sparse(wp, obj, snd, rcv), tgt)
        lm = O.bce_keras(O.forward_sparse(wm, obj, snd, rcv), tgt)
        fd = float(lp - lm) / (2 * h)
        assert abs(fd - float(grads[k][idx])) < 1e-8 + 1e-5 * abs(fd), (k, fd, float(grads[k][idx]))
