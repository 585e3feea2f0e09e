"""Boundary tests that EXECUTE THE REFERENCE'S OWN SOURCE (build container only: /root/reference is not on the GPU box, so
they skip there) plus CPU tests of the Keras-facing host logic.

  * /root/reference/src/main.py is imported UNMODIFIED with this repo's `Networks` / `Blocks` first on sys.path (and the
    pymunk / pyglet stand-ins of oracle/keras_shim, which is test infrastructure): `main.train_gnn` then loads a trajectory
    file, pads frames, runs its relation loops (main.py:66-81), labels, normalises and calls `.fit` on OUR model -- with no
    Keras or TensorFlow module anywhere in the process.  The dict it hands to `.fit` is checked.
  * the layout samplers of spwgnn_b200/synth.py are compared with JengaBuilder.create_world (JengaBuilder.py:137-192) and
    TowerCreator.create_world + drop_object (TowerCreator.py:106-213, 265-271) under the same seeded `random`.
  * the Keras-form Adam update and the validation split of `fit` against plain restatements.
"""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/src'
needs_reference = pytest.mark.skipif(not os.path.isdir(REF), reason='the reference source is only mounted in the build container')


def _run(script):
    env = dict(os.environ, PYTHONPATH='')
    res = subprocess.run([sys.executable, '-c', textwrap.dedent(script)], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


@needs_reference
def test_reference_main_py_runs_on_the_dropin(tmp_path):
    traj = tmp_path / 'jenga_model_5_6_test.txt'
    out = _run('''
        import contextlib, io, json, os, random, sys
        ROOT, REF = %r, %r
        sys.path[:0] = [os.path.join(ROOT, 'spwgnn_b200'), os.path.join(ROOT, 'oracle', 'keras_shim'), REF, ROOT]
        import numpy as np
        rng = random.Random(4)
        n, N, F = 5, 6, 7                                   # jenga: n_objects = n - 1 = 4 (main.py:30-31)
        data = []
        for t in range(N):
            tower = []
            for o in range(n - 1):
                x, y, w = 500 + rng.randint(0, 400), 110 + 80 * rng.randint(0, 3), rng.randint(50, 300)
                moves = rng.random() < 0.5
                frames = [[x + (0.3 * f if moves else 0.0), y - (0.2 * f if moves else 0.0), w] for f in range(F - (t %% 3))]
                tower.append(frames)
            data.append(tower)
        json.dump(data, open(%r, 'w'))
        with contextlib.redirect_stdout(io.StringIO()):
            import main                                        # the reference's main.py, unmodified
        import Networks, Blocks
        assert os.path.dirname(os.path.abspath(Networks.__file__)) == os.path.join(ROOT, 'spwgnn_b200'), Networks.__file__
        assert os.path.dirname(os.path.abspath(Blocks.__file__)) == os.path.join(ROOT, 'spwgnn_b200'), Blocks.__file__
        assert main.PropagationNetwork is Networks.PropagationNetwork and main.RelationalModel is Blocks.RelationalModel
        assert not any(m == 'keras' or m.startswith('keras.') or m.startswith('tensorflow') for m in sys.modules), 'Keras / TF got imported'
        seen = {}
        def fake_fit(self, x, y, **kw):
            seen.update(x=x, y=y, kw=kw, n_objects=self.n_objects)
        Networks.PropagationModel.fit = fake_fit
        with contextlib.redirect_stdout(io.StringIO()):
            model = main.train_gnn(n, N, %r, jenga=True)
        assert isinstance(model, Networks.PropagationModel) and seen['n_objects'] == n - 1
        assert seen['kw'] == dict(batch_size=32, epochs=10, validation_split=0.2, shuffle=True, verbose=1), seen['kw']
        from spwgnn_b200 import data as D
        from oracle import propnet as O
        raw0, y = D.training_arrays(%r, n, jenga=True)
        x = seen['x']
        assert np.array_equal(x['objects'], raw0 / 170.0) and np.array_equal(seen['y']['target'], y)
        rs, rr = O.build_relations_dense(raw0[:, :, :2], 170.0)
        assert np.array_equal(x['sender_relations'], rs) and np.array_equal(x['receiver_relations'], rr)
        assert x['propagation'].shape == (N, n - 1, 100) and not x['propagation'].any()
        active = int((rs.sum(1) > 0).sum())
        print(json.dumps(dict(ok=True, towers=N, active_relations=active, slots=int(rs.shape[0] * rs.shape[2]))))
    ''' % (ROOT, REF, str(traj), str(traj), str(traj)))
    assert out['ok'] and 0 < out['active_relations'] < out['slots']


@needs_reference
def test_layout_samplers_are_pinned_to_the_reference():
    out = _run('''
        import contextlib, io, json, os, random, sys
        ROOT, REF = %r, %r
        sys.path[:0] = [os.path.join(ROOT, 'oracle', 'keras_shim'), REF, ROOT]
        import numpy as np
        _ri = random.randint
        def randint36(a, b):                                   # Python 3.6 (the reference's interpreter) accepted integral floats
            ia, ib = int(a), int(b)
            assert ia == a and ib == b
            return _ri(ia, ib)
        random.randint = randint36
        with contextlib.redirect_stdout(io.StringIO()):
            import JengaBuilder as JB, TowerCreator as TC
        from spwgnn_b200 import synth
        checked = 0
        for n in (7, 9, 10, 18, 32, 54, 64):
            for s in range(25):
                random.seed(s)
                with contextlib.redirect_stdout(io.StringIO()):
                    jb = JB.JengaBuilder(n)                    # __init__ calls create_world (JengaBuilder.py:76)
                ref = np.array([[b.body.position[0], b.body.position[1], jb.get_rect_width(b)] for b in jb.flat_boxes])
                got = synth.g_jenga(n, random.Random(s))
                assert ref.shape == got.shape and np.array_equal(ref, got), ('jenga', n, s)
                checked += 1
        for n in (6, 8, 10, 12, 20):
            for s in range(60):
                random.seed(s)
                with contextlib.redirect_stdout(io.StringIO()):
                    tc = TC.TowerCreator(n)                    # __init__ calls create_world (TowerCreator.py:63)
                    tc.drop_object()                           # TowerCreator.py:265-271
                d = tc.dropped_object.body.position
                ref = np.array([[d[0], d[1], 150.0]] + [[b.body.position[0], b.body.position[1], 150.0] for b in tc.flat_boxes])
                got = synth.g_tower(n, random.Random(s))       # dropped block first: object 0 (TowerCreator.py:451)
                assert ref.shape == got.shape and np.array_equal(ref, got), ('tower', n, s)
                checked += 1
        print(json.dumps(dict(ok=True, checked=checked)))
    ''' % (ROOT, REF))
    assert out['ok'] and out['checked'] == 7 * 25 + 5 * 60


def test_keras_adam_matches_fp64_restatement():
    """engine.keras_adam_update_ against a numpy fp64 restatement of Keras 2.2 Adam (Networks.py:101: lr 5e-4, betas (0.9, 0.999),
    epsilon 1e-7 outside the square root, lr_t folding both bias corrections), ten steps."""
    from spwgnn_b200.engine import keras_adam_update_
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(5000)
    gs = [rng.standard_normal(5000) * 10.0 ** rng.integers(-4, 1) for _ in range(10)]
    p = torch.as_tensor(p0, dtype=torch.float32).clone()
    st = dict(t=0, m=torch.zeros(5000), v=torch.zeros(5000))
    p64, m, v = p0.astype(np.float32).astype(np.float64), np.zeros(5000), np.zeros(5000)
    lr, b1, b2, eps = 5e-4, 0.9, 0.999, 1e-7
    for t, g in enumerate(gs, 1):
        g32 = g.astype(np.float32)
        keras_adam_update_(p, torch.as_tensor(g32), st, lr, b1, b2, eps)
        g64 = g32.astype(np.float64)
        lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        m = b1 * m + (1 - b1) * g64
        v = b2 * v + (1 - b2) * g64 * g64
        p64 = p64 - lr_t * m / (np.sqrt(v) + eps)
    assert st['t'] == 10
    assert np.abs(p.numpy() - p64).max() < 2e-6
    # the first step moves every weight by lr * g / (|g| + eps sqrt(1 - b2)-ish): close to lr, the Keras signature
    p1 = torch.zeros(3)
    keras_adam_update_(p1, torch.tensor([1.0, -2.0, 1e-3]), dict(t=0, m=torch.zeros(3), v=torch.zeros(3)))
    assert np.allclose(p1.numpy(), [-5e-4, 5e-4, -5e-4], rtol=1e-2)


def test_fit_holds_out_the_last_fraction_like_keras(monkeypatch):
    """Keras: split_at = int(len(x) * (1 - validation_split)); samples [split_at:] are held out BEFORE shuffling (main.py:92-98
    relies on it).  B = 999, split 0.2 -> 799 training samples, 200 validation samples."""
    from spwgnn_b200.Networks import PropagationNetwork, PropagationModel
    model = PropagationNetwork().getModel(4)                        # no GPU needed until something computes
    trained, validated = [], []

    class _B:
        n_nodes = 1
    monkeypatch.setattr(PropagationModel, 'engine', property(lambda self: type('E', (), {'device': 'cpu'})()))
    monkeypatch.setattr(PropagationModel, '_batch_from_dict', lambda self, x, sel=None: (_B(), list(sel))[0] if not trained.append(list(sel)) else None)
    monkeypatch.setattr(PropagationModel, 'train_on_batch', lambda self, batch, tgt: (0.5, 1.0))
    monkeypatch.setattr(PropagationModel, 'test_on_batch', lambda self, batch, tgt: (validated.append(int(tgt.numel())) or 0.25, 0.5))
    B = 999
    y = {'target': np.zeros((B, 4, 1), np.float32)}
    x = {'objects': np.zeros((B, 4, 3))}
    hist = model.fit(x, y, batch_size=32, epochs=1, validation_split=0.2, shuffle=True, verbose=0, seed=1)
    calls = trained
    train_idx = sorted(i for c in calls[:25] for i in c)
    val_idx = sorted(i for c in calls[25:] for i in c)
    assert train_idx == list(range(799)) and val_idx == list(range(799, 999))
    assert hist.history['loss'] == [0.5] and hist.history['val_loss'] == [0.25]
