"""Minimal stand-in for standalone Keras 2.2/2.3 (TEST INFRASTRUCTURE ONLY).

Purpose: the reference (irmakguzey/SPWGNN) builds its propagation network with the
Keras functional API (`/root/reference/src/Networks.py:16-104`,
`/root/reference/src/Blocks.py:12-91`).  Keras/TensorFlow-1 are not installable in
this image, so `oracle/make_golden.py` puts THIS package on `sys.path` and imports
the reference's *unmodified* `Networks.py` / `main.py`.  The reference source then
wires its own graph (slices, concat order, dot axes, residual channels, head channel)
and this shim only supplies the documented semantics of each Keras op, evaluated with
torch float64 on the CPU.  Nothing under `spwgnn_b200/` imports this package.

Semantics implemented (Keras 2.2.x docs):
  Input, Dense (x @ kernel[in,out] + bias, glorot_uniform / zeros), Activation,
  Dropout (identity at inference), Permute, Subtract, Add, Concatenate (last axis),
  Lambda, dot(axes=(2,1)) == batched matmul, backend.reshape, Model (graph replay),
  Layer (build-once / call protocol), losses.binary_crossentropy (clip 1e-7, mean).
"""
import math
import torch

DTYPE = torch.float64
_GEN = torch.Generator().manual_seed(0)


def set_seed(seed):
    _GEN.manual_seed(int(seed))


class Node:
    """Symbolic tensor: evaluated lazily by Model._run."""
    _count = 0

    def __init__(self, op, parents, shape, name=None):
        self.op = op            # callable(list of torch tensors) -> torch tensor
        self.parents = parents
        self.shape = tuple(shape)
        self.name = name
        Node._count += 1
        self.uid = Node._count

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        shape = []
        for d, s in zip(self.shape, idx):
            if isinstance(s, slice):
                if d is None:
                    shape.append(None)
                else:
                    shape.append(len(range(*s.indices(d))))
            else:
                raise NotImplementedError('integer indexing on symbolic tensors')
        shape += list(self.shape[len(idx):])
        return Node(lambda xs, idx=idx: xs[0][idx], [self], shape)


def _evaluate(outputs, feed):
    cache = dict(feed)

    def ev(n):
        if n.uid in cache:
            return cache[n.uid]
        if n.op is None:
            raise KeyError('unfed Input %r' % (n.name,))
        v = n.op([ev(p) for p in n.parents])
        cache[n.uid] = v
        return v
    return [ev(o) for o in outputs]


def is_symbolic(x):
    return isinstance(x, Node)


def apply_op(fn, inputs, shape):
    """Apply fn eagerly on torch tensors, or record a Node on symbolic ones."""
    if any(is_symbolic(i) for i in inputs):
        return Node(lambda xs: fn(*xs), list(inputs), shape)
    return fn(*inputs)


def shape_of(x):
    if is_symbolic(x):
        return x.shape
    return (None,) + tuple(x.shape[1:])


from . import backend, layers, models, optimizers, regularizers, activations, utils, losses  # noqa: E402,F401
