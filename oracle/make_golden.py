"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference; the GPU box does not have it):

    python oracle/make_golden.py

How: `oracle/keras_shim/` (stand-ins for keras / tensorflow / pymunk / pyglet, torch-fp64
semantics of the handful of ops the reference uses) is put on sys.path, then the reference's
UNMODIFIED modules are imported from /root/reference/src:
  * `main.train_gnn`              (main.py:25-110)  -> frame padding, the relation loops
                                    (main.py:66-81), labels (main.py:8-23), /170 (main.py:91)
                                    and the `.fit(...)` call whose arguments the shim records;
  * `Networks.PropagationNetwork` (Networks.py:12-104) -> the graph wiring, replayed in fp64 to
                                    get probabilities and, through torch autograd, the gradient
                                    of the Keras BCE loss w.r.t. all 22 weight tensors;
  * `JengaBuilder.predict_stabilities` (JengaBuilder.py:301-329) and
    `TowerCreator.predict_stabilities` (TowerCreator.py:402-431) -> the inference-time glue
                                    (normalised positions thresholded against 170 => fully
                                    connected), called unbound on a plain namespace object.
Nothing of the reference is copied: its code runs where it lies and only its numeric
outputs are stored.  Keras' own arithmetic (TF1 kernels) stays unpinned -- see oracle/propnet.py.
"""
import json
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/src'
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit('reference not mounted at %s (golden vectors are generated in the '
                         'build container only)' % REF)
    sys.path.insert(0, os.path.join(HERE, 'keras_shim'))
    sys.path.insert(0, REF)
    import keras  # the shim
    import Networks, main, JengaBuilder, TowerCreator  # noqa: E401  (reference modules, unmodified)
    return keras, Networks, main, JengaBuilder, TowerCreator


def _jenga_like_tower(rng, n):
    """Small random layout in raw pixels with both near (<170) and far (>170) block pairs."""
    out, layer, x = [], 0, 400 + rng.randint(0, 200)
    for i in range(n):
        wdt = rng.randint(50, 300)
        x += wdt
        out.append([x - wdt / 2.0, 70 + 40 + 80.0 * layer, float(wdt)])
        x += rng.randint(0, 50)
        if rng.random() < 0.45:
            layer += 1
            x = 400 + rng.randint(0, 200)
    return out


def _weights_of(pn):
    names, tensors = [], []
    for net, mdl in (('rm', pn.relnet), ('om', pn.objnet), ('rmp', pn.relnetp), ('omp', pn.objnetp)):
        tw = mdl.trainable_weights
        for li in range(len(tw) // 2):
            names += ['%s.w%d' % (net, li), '%s.b%d' % (net, li)]
            tensors += [tw[2 * li], tw[2 * li + 1]]
    return names, tensors


def _randomise_biases(tensors, names, seed):
    g = torch.Generator().manual_seed(seed)
    for n, t in zip(names, tensors):
        if '.b' in n:
            t.data.copy_(((torch.rand(t.shape, generator=g, dtype=torch.float64) * 2 - 1) * 0.1).float().double())


WEIGHT_SEED = 7      # every case draws the same initial weights, so they are stored once


def case_train(keras, Networks, main, n_objects, n_traj, seed, tmpdir, with_grads=True):
    """Reference training-time path: main.train_gnn on a synthetic trajectory file."""
    rng = random.Random(seed)
    data = []
    for t in range(n_traj):
        tower = _jenga_like_tower(rng, n_objects)
        n_frames = rng.randint(1, 3)                       # ragged frame counts: exercises main.py:52-63
        traj = []
        for o in range(n_objects):
            frames = []
            for f in range(n_frames):
                # every other block drifts a little so calculate_stability yields both labels
                drift = (0.4 * f) if (o + t) % 2 else 0.0
                frames.append([tower[o][0] + drift, tower[o][1] - drift, tower[o][2]])
            traj.append(frames)
        data.append(traj)
    path = os.path.join(tmpdir, 'jenga_model_%d.txt' % seed)
    with open(path, 'w') as f:
        json.dump(data, f)

    created = []

    class Recording(Networks.PropagationNetwork):
        def __init__(self):
            super().__init__()
            created.append(self)
    main.PropagationNetwork = Recording
    keras.set_seed(WEIGHT_SEED)
    model = main.train_gnn(n_objects + 1, n_traj, path, jenga=True)   # jenga: n_objects = n-1, object_dim 3
    pn = created[-1]
    call = model.fit_calls[0]
    x, y = call['x'], call['y']['target']

    names, tensors = _weights_of(pn)
    _randomise_biases(tensors, names, WEIGHT_SEED + 1)
    probs = model.predict_torch(x)                                     # (T, N, 1) fp64, grad-enabled
    yt = torch.as_tensor(y, dtype=torch.float64)
    loss = keras.losses.binary_crossentropy(yt, probs).mean()
    grads = torch.autograd.grad(loss, tensors)

    raw_pos = np.array([[data[t][o][0][0:2] for o in range(n_objects)] for t in range(n_traj)], dtype=np.float64)
    out = dict(
        kind='train', n_objects=n_objects, raw_pos=raw_pos,
        objects=np.asarray(x['objects']), sender_relations=np.asarray(x['sender_relations']),
        receiver_relations=np.asarray(x['receiver_relations']), propagation_shape=np.asarray(x['propagation'].shape),
        target=np.asarray(y), probs=probs.detach().numpy(), loss=float(loss),
        fit_kwargs=json.dumps(call['kwargs']), traj_json=json.dumps(data),
    )
    for n, t, g in zip(names, tensors, grads):
        out['w:' + n] = t.detach().numpy()
        if with_grads:
            out['g:' + n] = g.numpy().astype(np.float32)      # fp32 storage: 6e-8 rel, tolerance is 1e-5
    return out


def case_predict(keras, Networks, Builder, n_objects, seed, jenga):
    """Reference inference-time glue: <Builder>.predict_stabilities on a plain namespace."""
    rng = random.Random(seed)
    tower = _jenga_like_tower(rng, n_objects)
    keras.set_seed(WEIGHT_SEED)
    pn = Networks.PropagationNetwork()
    model = pn.getModel(n_objects, 3 if jenga else 2)
    captured = {}
    orig_predict = model.predict

    def spy(x, **kw):
        captured.update(x)
        return orig_predict(x, **kw)
    model.predict = spy
    if jenga:
        fake = types.SimpleNamespace(n=n_objects + 1, relation_threshold=170.0, gnn_model=model,
                                     trajectories=[[[[b[0], b[1], b[2]]] for b in tower]])
    else:
        fake = types.SimpleNamespace(n=n_objects - 1, jenga=False, relation_threshold=170.0, gnn_model=model,
                                     trajectories=[[[[b[0], b[1]]] for b in tower]])
    names, tensors = _weights_of(pn)
    _randomise_biases(tensors, names, WEIGHT_SEED + 1)
    Builder.predict_stabilities(fake)
    out = dict(kind='predict', n_objects=n_objects, raw=np.array(tower, dtype=np.float64),
               objects=np.asarray(captured['objects']), sender_relations=np.asarray(captured['sender_relations']),
               receiver_relations=np.asarray(captured['receiver_relations']),
               probs=np.asarray(fake.stabilities))
    for n, t in zip(names, tensors):
        out['w:' + n] = t.detach().numpy()
    return out


def main_():
    keras, Networks, main, JengaBuilder, TowerCreator = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    import io, contextlib
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        cases = {
            'train_n2': case_train(keras, Networks, main, 2, 3, 11, tmp),
            'train_n4': case_train(keras, Networks, main, 4, 4, 12, tmp),
            'train_n7': case_train(keras, Networks, main, 7, 5, 13, tmp),     # C1 shape (N=7)
            'train_n9': case_train(keras, Networks, main, 9, 3, 14, tmp, with_grads=False),  # main.py default (n=10, jenga)
            'predict_jenga_n9': case_predict(keras, Networks, JengaBuilder.JengaBuilder, 9, 21, True),
            'predict_jenga_n5': case_predict(keras, Networks, JengaBuilder.JengaBuilder, 5, 22, True),
        }
    # identical weights in every case (same seed, same build order): store them once, as the
    # fp32 values they are
    wref = {k: v for k, v in cases['train_n7'].items() if k.startswith('w:')}
    for name, c in cases.items():
        for k in list(c):
            if k.startswith('w:'):
                assert np.array_equal(c[k], wref[k]), (name, k)
                assert np.array_equal(c[k].astype(np.float32).astype(np.float64), c[k])
                del c[k]
    np.savez_compressed(os.path.join(OUT, 'weights.npz'), **{k[2:]: v.astype(np.float32) for k, v in wref.items()})
    for name, c in cases.items():
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **c)
        act = c['sender_relations'].sum() / max(1, c['sender_relations'].shape[0])
        sys.stderr.write('%-18s N=%d  active edges/tower=%.1f of %d  probs[0,:3]=%s\n' % (
            name, c['n_objects'], act, c['n_objects'] * (c['n_objects'] - 1), np.round(c['probs'][0, :3, 0], 6)))


if __name__ == '__main__':
    main_()
