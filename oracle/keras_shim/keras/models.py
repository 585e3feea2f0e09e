"""keras.models stand-in (TEST INFRASTRUCTURE ONLY): Model replays the recorded graph."""
import numpy as np
import torch
from keras import Node, _evaluate, is_symbolic, DTYPE


def _walk(outputs):
    seen, order = set(), []

    def rec(n):
        if n.uid in seen:
            return
        seen.add(n.uid)
        for p in n.parents:
            rec(p)
        order.append(n)
    for o in outputs:
        rec(o)
    return order


class Model:
    def __init__(self, inputs, outputs, name=None):
        self.inputs = list(inputs)
        self.outputs = list(outputs)
        self.fit_calls = []
        self.compiled = None

    # -- Keras API used by the reference ------------------------------------------------
    def build(self, input_shape):
        pass

    @property
    def trainable_weights(self):
        ws = []
        for n in _walk(self.outputs):
            owner = getattr(n, 'owner', None)
            if owner is not None:
                for w in owner.trainable_weights:
                    if not any(w is x for x in ws):
                        ws.append(w)
        return ws

    def call(self, x):
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        out_shape = self.outputs[0].shape
        if any(is_symbolic(t) for t in xs):
            return Node(lambda ts: self._run(ts)[0], xs, out_shape)
        return self._run(xs)[0]

    def compile(self, optimizer=None, loss=None, metrics=None):
        self.compiled = dict(optimizer=optimizer, loss=loss, metrics=metrics)

    def fit(self, x=None, y=None, **kwargs):
        self.fit_calls.append(dict(x=x, y=y, kwargs=kwargs))
        return None

    def predict(self, x, **kwargs):
        with torch.no_grad():
            return self.predict_torch(x).numpy()

    # -- helpers -------------------------------------------------------------------------
    def _feed(self, x):
        if isinstance(x, dict):
            vals = [x[i.name] for i in self.inputs]
        else:
            vals = list(x)
        return [v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v), dtype=DTYPE) for v in vals]

    def _run(self, vals):
        feed = {i.uid: v for i, v in zip(self.inputs, vals)}
        return _evaluate(self.outputs, feed)

    def predict_torch(self, x):
        return self._run(self._feed(x))[0]


Sequential = Model
