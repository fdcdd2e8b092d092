"""Minimal TensorFlow-API shim over PyTorch-CPU, used ONLY by tests/golden/make_golden_loop.py to execute the
reference's own, unmodified ``GNN/Models/*.py`` (TensorFlow cannot be installed in the build image).

It implements exactly the calls those files make on the hot path, with the TF/Keras-2 semantics listed in
SURVEY.md App. B (eager ``tf.constant`` = identity, ``while_loop`` = Python loop with ``bool(cond)``,
``sparse_dense_matmul(adjoint_a=True)`` = sequential accumulation in stored order, Keras BatchNormalization:
batch mean / biased variance in training, eps 1e-3, momentum 0.99 with assign_sub updates).  It pins the
reference's Python-level control flow and op order - not TensorFlow's kernels.
"""
import types

import numpy as np
import torch

_FLOATX = ["float32"]
_DT = {"float32": torch.float32, "float64": torch.float64, "bool": torch.bool, "int64": torch.int64,
       "int32": torch.int32}
bool = torch.bool          # noqa: A001  (tf.bool)
int64 = torch.int64
float32 = torch.float32
Tensor = torch.Tensor
newaxis = None
RANDOM_DRAWS = []          # every tf.random.normal draw, in order (the goldens store them as explicit state0)
_GEN = torch.Generator().manual_seed(1234)


def _dt(d):
    if d is None:
        return _DT[_FLOATX[0]]
    if isinstance(d, str):
        return _DT[d]
    return d


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(_dt(dtype))
    return torch.as_tensor(np.asarray(x), dtype=None if dtype is None else _dt(dtype))


def constant(x, dtype=None, **kw):
    if isinstance(x, torch.Tensor):       # eager: tf.constant(EagerTensor) returns the tensor itself
        return x if dtype is None or x.dtype == _dt(dtype) else x.to(_dt(dtype))
    return torch.as_tensor(np.asarray(x), dtype=_dt(dtype) if dtype is not None or isinstance(x, float) else None)


def zeros(shape, dtype=None):
    return torch.zeros(tuple(int(s) for s in shape), dtype=_dt(dtype))


def ones_like(x, dtype=None):
    return torch.ones_like(x, dtype=_dt(dtype) if dtype is not None else None)


def squeeze(x, axis=None):
    x = _t(x)
    return x.squeeze() if axis is None else x.squeeze(axis)


def sqrt(x): return torch.sqrt(x)
def square(x): return torch.square(x)
def subtract(a, b): return torch.subtract(a, b)
def greater(a, b): return torch.gt(a, b)
def less(a, b): return torch.lt(_t(a), b) if not isinstance(b, torch.Tensor) else torch.lt(a, b)
def logical_and(a, b): return torch.logical_and(_t(a), _t(b))
def reduce_any(x): return torch.any(x)
def reduce_sum(x, axis=None): return torch.stack(list(x)).sum(0) if isinstance(x, (list, tuple)) else (x.sum() if axis is None else x.sum(dim=axis))
def reduce_mean(x, axis=None):
    x = torch.stack(list(x)) if isinstance(x, (list, tuple)) else x
    return x.mean() if axis is None else x.mean(dim=axis)
def concat(xs, axis): return torch.cat([_t(x) for x in xs], dim=axis)
def boolean_mask(x, mask): return x[_t(mask).to(torch.bool)]
def gather(x, idx): return x[_t(idx).to(torch.int64)]
def reshape(x, shape): return x.reshape(tuple(int(s) for s in shape))
def cast(x, dtype): return x.to(_dt(dtype)) if isinstance(x, torch.Tensor) else x
def where(m): return torch.nonzero(_t(m).to(torch.bool))          # [n, 1] for a vector mask


def scatter_nd(indices, updates, shape):
    out = torch.zeros(tuple(int(s) for s in shape), dtype=updates.dtype)
    return out.index_put((indices[:, 0],), updates, accumulate=True)


def while_loop(cond, body, loop_vars):
    vars_ = list(loop_vars)
    while __builtins__["bool"](cond(*vars_)) if isinstance(__builtins__, dict) else __builtins__.bool(cond(*vars_)):
        vars_ = list(body(*vars_))
    return vars_


class SparseTensor:
    def __init__(self, indices, values, dense_shape):
        self.indices = _t(indices).to(torch.int64)
        self.values = _t(values)
        self.dense_shape = [int(s) for s in _t(dense_shape).reshape(-1)]
        self.shape = self.dense_shape


def _sdm(sp_a, b, adjoint_a=False):
    r, c = sp_a.indices[:, 0], sp_a.indices[:, 1]
    rows_out = sp_a.dense_shape[1] if adjoint_a else sp_a.dense_shape[0]
    out = torch.zeros((rows_out, b.shape[1]), dtype=b.dtype)
    if adjoint_a:
        return out.index_add(0, c, sp_a.values.to(b.dtype)[:, None] * b[r])
    return out.index_add(0, r, sp_a.values.to(b.dtype)[:, None] * b[c])


sparse = types.SimpleNamespace(sparse_dense_matmul=_sdm, reorder=lambda x: x)
math = types.SimpleNamespace(scalar_mul=lambda s, x: s * x)


def _normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None):
    x = (torch.randn(tuple(int(s) for s in shape), generator=_GEN, dtype=torch.float64) * stddev + mean).to(_dt(dtype))
    RANDOM_DRAWS.append(x.clone())
    return x


random = types.SimpleNamespace(normal=_normal)


class GradientTape:
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def gradient(self, loss, sources):
        flat = [v for grp in sources for v in grp]
        g = torch.autograd.grad(loss, flat, allow_unused=True)
        out, i = [], 0
        for grp in sources:
            out.append(list(g[i:i + len(grp)])); i += len(grp)
        return out


# ---- keras ------------------------------------------------------------------------------------------------
SELU_SCALE, SELU_ALPHA = 1.0507009873554805, 1.6732632423543772


def _activation(name, x):
    if name in (None, "linear"): return x
    if name == "tanh": return torch.tanh(x)
    if name == "sigmoid": return torch.sigmoid(x)
    if name == "relu": return torch.relu(x)
    if name == "selu": return torch.where(x < 0, (SELU_SCALE * SELU_ALPHA) * (torch.exp(x) - 1), SELU_SCALE * x)
    if name == "softmax": return torch.softmax(x, dim=-1)
    raise ValueError(name)


class _Dense:
    def __init__(self, W, b, activation):
        self.kernel, self.bias, self.activation = W, b, activation
        self.trainable_variables = [self.kernel, self.bias]
    def __call__(self, x, training=False): return _activation(self.activation, x @ self.kernel + self.bias)


class _BatchNormalization:
    def __init__(self, gamma, beta, mm, mv, eps=1e-3, momentum=0.99):
        self.gamma, self.beta, self.moving_mean, self.moving_variance = gamma, beta, mm, mv
        self.epsilon, self.momentum = eps, momentum
        self.trainable_variables = [self.gamma, self.beta]
    def __call__(self, x, training=False):
        if __builtins__["bool"](training) if isinstance(__builtins__, dict) else __builtins__.bool(training):
            mean = x.mean(dim=0)
            var = ((x - mean.detach()) ** 2).mean(dim=0)
            with torch.no_grad():
                decay = 1.0 - self.momentum
                self.moving_mean -= (self.moving_mean - mean) * decay
                self.moving_variance -= (self.moving_variance - var) * decay
        else:
            mean, var = self.moving_mean, self.moving_variance
        inv = torch.rsqrt(var + self.epsilon) * self.gamma
        return x * inv + (self.beta - mean * inv)


class _Sequential:
    def __init__(self, layers, name=None):
        self.layers, self.name = layers, name
    @property
    def trainable_variables(self): return [v for l in self.layers for v in l.trainable_variables]
    def __call__(self, x, training=False):
        for l in self.layers: x = l(x, training=training)
        return x
    def compile(self, *a, **k): pass


class _Model:
    def __init__(self, name=None, **kw): self._name = name
    def __call__(self, inputs, training=False, **kw): return self.call(inputs, training=training)
    def compile(self, *a, **k): pass


keras = types.SimpleNamespace(
    Model=_Model,
    backend=types.SimpleNamespace(floatx=lambda: _FLOATX[0]),
    models=types.SimpleNamespace(Sequential=_Sequential, clone_model=lambda m: m),
    layers=types.SimpleNamespace(Dense=_Dense, BatchNormalization=_BatchNormalization),
    utils=types.SimpleNamespace(Sequence=object),
)


def set_floatx(name):
    _FLOATX[0] = name


def net_from_dict(net, dtype):
    """oracle-format net dict -> shim Sequential with torch leaves (Keras variable order)."""
    leaf = lambda a, rg=True: torch.tensor(np.asarray(a), dtype=dtype).requires_grad_(rg)
    layers = []
    if net.get("bn") is not None:
        b = net["bn"]
        layers.append(_BatchNormalization(leaf(b["gamma"]), leaf(b["beta"]), leaf(b["moving_mean"], False),
                                          leaf(b["moving_var"], False), b["eps"], b["momentum"]))
    for l in net["layers"]:
        layers.append(_Dense(leaf(l["W"]), leaf(l["b"]), l["act"]))
    return _Sequential(layers)
