"""Drop-in for /root/reference/src/Blocks.py: the names `RelationalModel` and `ObjectModel` that `main.py:4` star-imports.

In the reference these are Keras `Layer`s that wrap a row-wise MLP (flatten (B, S, F) -> (B.S, F), Dense + relu for every
filter but the last, linear last layer, reshape back; Blocks.py:12-47, 51-89) and `Networks.py:46-50` instantiates four of
them.  Here the four MLPs are fused into the sm_100a kernels of libspwgnn.so and their 16 tensors live in ONE flat
parameter buffer (spwgnn_b200/params.py), so these classes are DESCRIPTORS: they keep the constructor signature and the
attributes callers can see (`input_size`, `n_of_features`, `output_size`, `getRelnet()` / `getObjnet()` for the weight
sharing across model sizes, Networks.py:40-56) and give access to the tensors of "their" MLP inside a network's buffer.
They do no arithmetic and import no Keras.
"""

__all__ = ['RelationalModel', 'ObjectModel', 'regul']

regul = 0.001     # Blocks.py:9 (L2 factors the reference declares; whether they reach its loss is Keras-version dependent)


class _MlpDescriptor:
    """input_size: tuple of leading dimensions; n_of_features: MLP input width; filters: layer widths."""

    def __init__(self, input_size, n_of_features, filters, shared=None, reuse_model=False, **kwargs):
        self.input_size = tuple(input_size) if isinstance(input_size, (tuple, list)) else (input_size,)
        self.n_of_features = int(n_of_features)
        self.filters = [int(f) for f in filters]
        self.output_size = self.filters[-1]
        self._net = shared if reuse_model else self      # the object that owns the weights (weight sharing: Blocks.py:18-19)
        self.prefix = kwargs.get('prefix')               # 'rm' | 'om' | 'rmp' | 'omp' once bound to a PropagationNetwork
        self.network = kwargs.get('network')

    def compute_output_shape(self, input_shape=None):
        return (None,) + self.input_size + (int(self.output_size),)

    def layer_shapes(self):
        """[(in, out), ...] of the Dense kernels (Blocks.py:22-27 / 62-66)."""
        dims = [self.n_of_features] + self.filters
        return list(zip(dims[:-1], dims[1:]))

    def weights(self):
        """The tensors of this MLP inside the bound network's parameter buffer: [kernel0, bias0, kernel1, bias1, ...]."""
        if self.network is None or self.prefix is None:
            raise RuntimeError('this descriptor is not bound to a PropagationNetwork')
        views = self.network.engine.params.views
        out = []
        for i in range(len(self.filters)):
            out += [views['%s.w%d' % (self.prefix, i)], views['%s.b%d' % (self.prefix, i)]]
        return out


class RelationalModel(_MlpDescriptor):
    """Blocks.py:12-50."""

    def __init__(self, input_size, n_of_features, filters, rm=None, reuse_model=False, **kwargs):
        super().__init__(input_size, n_of_features, filters, rm, reuse_model, **kwargs)

    def getRelnet(self):
        return self._net


class ObjectModel(_MlpDescriptor):
    """Blocks.py:51-91."""

    def __init__(self, input_size, n_of_features, filters, om=None, reuse_model=False, **kwargs):
        super().__init__(input_size, n_of_features, filters, om, reuse_model, **kwargs)

    def getObjnet(self):
        return self._net
