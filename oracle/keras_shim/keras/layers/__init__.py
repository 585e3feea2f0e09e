"""keras.layers stand-in (TEST INFRASTRUCTURE ONLY; see keras_shim/keras/__init__.py)."""
import math
import torch
import keras as _k
from keras import Node, apply_op, shape_of, is_symbolic


class Layer:
    """Keras Layer protocol: first __call__ runs build(input_shape), then call(x)."""

    def __init__(self, **kwargs):
        self.name = kwargs.get('name')
        self.built = False
        if not hasattr(self, 'trainable_weights'):
            self.trainable_weights = []

    def build(self, input_shape):
        self.built = True

    def call(self, x):
        raise NotImplementedError

    def compute_output_shape(self, input_shape):
        return input_shape

    def __call__(self, x):
        in_shape = [shape_of(t) for t in x] if isinstance(x, (list, tuple)) else shape_of(x)
        if not self.built:
            self.build(in_shape)
            self.built = True
        out = self.call(x)
        if is_symbolic(out) and self.name is not None:
            out.name = self.name
        return out


def Input(shape, name=None):
    return Node(None, [], (None,) + tuple(shape), name=name)


class Dense(Layer):
    def __init__(self, units, activation=None, kernel_regularizer=None, bias_regularizer=None,
                 activity_regularizer=None, **kwargs):
        super().__init__(**kwargs)
        self.units = int(units)
        self.activation = activation
        self.kernel = None
        self.bias = None

    def build(self, input_shape):
        fan_in = int(input_shape[-1])
        limit = math.sqrt(6.0 / (fan_in + self.units))          # glorot_uniform
        k = (torch.rand(fan_in, self.units, generator=_k._GEN, dtype=torch.float64) * 2 - 1) * limit
        # weights are fp32 in the real reference: round so every consumer sees identical values
        self.kernel = k.float().double().requires_grad_(True)
        self.bias = torch.zeros(self.units, dtype=torch.float64, requires_grad=True)
        self.trainable_weights = [self.kernel, self.bias]

    def call(self, x):
        act = _ACT[self.activation or 'linear']
        shp = shape_of(x)[:-1] + (self.units,)
        out = apply_op(lambda t: act(t @ self.kernel + self.bias), [x], shp)
        if is_symbolic(out):
            out.owner = self
        return out


_ACT = {
    'linear': lambda t: t,
    'relu': torch.relu,
    'tanh': torch.tanh,
    'sigmoid': torch.sigmoid,
}


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.fn = _ACT[activation]

    def call(self, x):
        return apply_op(self.fn, [x], shape_of(x))


class Dropout(Layer):
    """Inference semantics (identity).  Training-mode masks cannot match TF's RNG; parity is
    defined with dropout off."""

    def __init__(self, rate, **kwargs):
        super().__init__(**kwargs)
        self.rate = rate

    def call(self, x):
        return apply_op(lambda t: t, [x], shape_of(x))


class Permute(Layer):
    def __init__(self, dims, **kwargs):
        super().__init__(**kwargs)
        self.dims = tuple(dims)

    def call(self, x):
        s = shape_of(x)
        shp = (s[0],) + tuple(s[d] for d in self.dims)
        perm = (0,) + self.dims
        return apply_op(lambda t: t.permute(*perm), [x], shp)


class Subtract(Layer):
    def call(self, xs):
        return apply_op(lambda a, b: a - b, list(xs), shape_of(xs[0]))


class Add(Layer):
    def call(self, xs):
        def f(*ts):
            out = ts[0]
            for t in ts[1:]:
                out = out + t
            return out
        return apply_op(f, list(xs), shape_of(xs[0]))


class Multiply(Layer):
    def call(self, xs):
        return apply_op(lambda a, b: a * b, list(xs), shape_of(xs[0]))


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, xs):
        shapes = [shape_of(t) for t in xs]
        shp = list(shapes[0])
        shp[self.axis] = sum(s[self.axis] for s in shapes)
        return apply_op(lambda *ts: torch.cat(ts, dim=self.axis), list(xs), shp)


class Lambda(Layer):
    def __init__(self, function, output_shape=None, **kwargs):
        super().__init__(**kwargs)
        self.function = function

    def call(self, x):
        return self.function(x)


class Reshape(Layer):
    def __init__(self, target_shape, **kwargs):
        super().__init__(**kwargs)
        self.target_shape = tuple(target_shape)

    def call(self, x):
        return apply_op(lambda t: t.reshape((t.shape[0],) + self.target_shape), [x],
                        (None,) + self.target_shape)


def dot(inputs, axes, normalize=False):
    """keras.layers.dot: batch_dot contracting axes[0] of inputs[0] with axes[1] of inputs[1].
    Only the (2, 1) case used by the reference (== batched matmul) is provided."""
    a, b = inputs
    assert tuple(axes) == (2, 1), axes
    sa, sb = shape_of(a), shape_of(b)
    assert sa[2] == sb[1], (sa, sb)
    return apply_op(lambda x, y: torch.matmul(x, y), [a, b], (None, sa[1], sb[2]))


class _Unavailable(Layer):
    def __init__(self, *a, **k):
        raise NotImplementedError('not used on the SPWGNN hot path')


LSTM = CuDNNLSTM = TimeDistributed = Conv1D = MaxPooling1D = GlobalAveragePooling1D = _Unavailable
BatchNormalization = _Unavailable
