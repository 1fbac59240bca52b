"""Minimal eager `tensorflow` API shim backed by torch CPU float64.

PURPOSE: TensorFlow is not installable in the build container, so the reference's own Python
source (/root/reference/vfegpssm/*.py) is executed UNMODIFIED on top of this shim to produce
golden vectors (tests/golden/make_reference_golden.py).  Only the ~40 tf symbols that the hot
path touches are provided; each maps to the torch op with the same documented semantics
(matmul/cholesky/triangular_solve/solve/logdet are LAPACK-backed in both).  `tf.gradients` is
torch reverse-mode autograd.  Test infrastructure only.
"""
import contextlib
import types

import numpy as np
import torch

float64 = torch.float64
float32 = torch.float32
int32 = torch.int32
int64 = torch.int64

_TRAINABLE = []
_RNG = torch.Generator().manual_seed(0)
NOISE_LOG = []          # every tf.random.normal draw, in call order
UNIFORM_LOG = []        # every uniform draw behind tfp Categorical.sample, in call order
PLACEHOLDER_VALUES = {}  # dtype -> value returned by tf.compat.v1.placeholder
# TF1 graph re-execution: a `session.run(op)` of the reference evaluates the graph on the CURRENT variable values.  The
# eager shim emulates that by re-running the model constructor; VARIABLE_OVERRIDES maps the creation index of a Variable
# (deterministic: the constructor creates them in a fixed order) to the value it holds now.
VARIABLE_OVERRIDES = {}
_VAR_COUNT = [0]


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None or x.dtype == dtype else x.to(dtype)
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        return torch.stack([_t(v, dtype) for v in x])
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def Variable(initial_value, dtype=None, trainable=True, name=None):
    idx = _VAR_COUNT[0]
    _VAR_COUNT[0] += 1
    if idx in VARIABLE_OVERRIDES:
        initial_value = VARIABLE_OVERRIDES[idx]
    v = _t(initial_value, dtype).detach().clone()
    v._tf_index = idx
    if v.dtype.is_floating_point:
        v.requires_grad_(True)
    v._tf_name = name
    v._tf_trainable = bool(trainable)
    if trainable:
        _TRAINABLE.append(v)
    return v


def constant(value, dtype=None):
    return _t(value, dtype)


def convert_to_tensor(value, dtype=None):
    return _t(value, dtype)


def cast(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(x, dtype=dtype)


def identity(x):
    return x


exp = torch.exp
sqrt = torch.sqrt
square = torch.square


def maximum(a, b):
    return torch.maximum(_t(a), torch.as_tensor(b, dtype=_t(a).dtype))


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    if axis is None:
        return torch.sum(x)
    return torch.sum(x, dim=axis, keepdim=keepdims)


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


def transpose(x, perm=None):
    x = _t(x)
    if perm is None:
        return x.permute(*reversed(range(x.dim())))
    return x.permute(*perm)


def shape(x):
    return _t(x).shape


def squeeze(x):
    return torch.squeeze(_t(x))


def fill(dims, value):
    return torch.ones(tuple(int(d) for d in dims), dtype=torch.float64) * value


def eye(n, dtype=torch.float64):
    return torch.eye(int(n), dtype=dtype)


def ones(shape_, dtype=torch.float64):
    if isinstance(shape_, int):
        shape_ = (shape_,)
    return torch.ones(tuple(int(s) for s in shape_), dtype=dtype)


def zeros(shape_, dtype=torch.float64):
    if isinstance(shape_, int):
        shape_ = (shape_,)
    return torch.zeros(tuple(int(s) for s in shape_), dtype=dtype)


def ones_like(x):
    return torch.ones_like(_t(x).detach())


def zeros_like(x):
    return torch.zeros_like(_t(x).detach())


def tile(x, multiples):
    return _t(x).repeat(*[int(m) for m in multiples])


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def stack(values, axis=0):
    if all(not isinstance(v, (torch.Tensor, list, tuple)) for v in values):
        return [int(v) for v in values]
    return torch.stack([_t(v) for v in values], dim=axis)


def concat(values, axis):
    return torch.cat([_t(v) for v in values], dim=axis)


def gather(x, idx, axis=None):
    """tf.gather: axis defaults to 0 (the reference's particle resampling); kernels_multi_output.py passes axis=-1."""
    x = _t(x)
    idx = torch.as_tensor(idx, dtype=torch.long)
    ax = 0 if axis is None else axis
    if idx.dim() == 0:
        return x.select(ax, int(idx))
    return torch.index_select(x, ax, idx)


def unstack(x, axis=0):
    return list(torch.unbind(_t(x), dim=axis))


def assert_equal(a, b):
    assert int(a) == int(b)
    return None


@contextlib.contextmanager
def control_dependencies(deps):
    yield


def gradients(ys, xs):
    xs = list(xs)
    if not xs:
        return []
    gs = torch.autograd.grad(ys, xs, allow_unused=True, retain_graph=True)
    return [torch.zeros_like(x) if g is None else g for x, g in zip(xs, gs)]


# ---- tf.linalg
linalg = types.SimpleNamespace()
linalg.cholesky = torch.linalg.cholesky
linalg.matrix_transpose = lambda x: _t(x).transpose(-1, -2)
linalg.triangular_solve = lambda matrix, rhs, lower=True, adjoint=False: torch.linalg.solve_triangular(
    _t(matrix), _t(rhs), upper=not lower)
linalg.solve = lambda a, b: torch.linalg.solve(_t(a), _t(b))
linalg.logdet = lambda a: torch.linalg.slogdet(_t(a))[1]
linalg.diag_part = lambda a: torch.diagonal(_t(a), dim1=-2, dim2=-1)
linalg.tensor_diag_part = linalg.diag_part
linalg.diag = lambda v: torch.diag(_t(v))
linalg.inv = lambda a: torch.linalg.inv(_t(a))


def _svd(x, full_matrices=False):
    u, s, vh = torch.linalg.svd(_t(x).detach(), full_matrices=full_matrices)
    return s, u, vh.transpose(-1, -2)


linalg.svd = _svd

# ---- tf.math
math = types.SimpleNamespace(log=lambda x: torch.log(_t(x)), exp=torch.exp, sqrt=torch.sqrt)

# ---- tf.random
random = types.SimpleNamespace()


def _normal(shape_, dtype=torch.float64):
    z = torch.randn(tuple(int(s) for s in shape_), dtype=dtype, generator=_RNG)
    NOISE_LOG.append(z.clone())
    return z


random.normal = _normal
random.set_seed = lambda s: _RNG.manual_seed(int(s))


# ---- tf.compat.v1
class _Session:
    def __init__(self, config=None):
        pass

    def run(self, fetches, feed_dict=None):
        # TF returns NumPy values; variables keep their current (construction-time) value in this eager shim
        def conv(f):
            if isinstance(f, torch.Tensor):
                return f.detach().numpy().copy()
            if isinstance(f, (list, tuple)):
                return type(f)(conv(v) for v in f)
            return f
        return conv(fetches)


class _Adam:
    def __init__(self, lr):
        self.lr = lr

    def minimize(self, loss):
        tv = [v for v in _TRAINABLE]
        if not tv:
            raise ValueError("No variables to optimize.")
        gs = torch.autograd.grad(loss, tv, allow_unused=True, retain_graph=True)
        self.grads_and_vars = [(torch.zeros_like(v) if g is None else g, v) for g, v in zip(gs, tv)]
        return self.grads_and_vars


class _HList(list):
    """A placeholder value that can also be a feed_dict key (the reference builds {placeholder: value} dicts)."""
    __hash__ = object.__hash__


def _placeholder(dtype, shape=None):
    v = PLACEHOLDER_VALUES[dtype]
    return _HList(v) if isinstance(v, list) else v


def _config_proto():
    return types.SimpleNamespace(gpu_options=types.SimpleNamespace(allow_growth=False))


v1 = types.SimpleNamespace(
    placeholder=_placeholder,
    assign=lambda var, val: (var, val),
    Session=_Session,
    ConfigProto=_config_proto,
    global_variables_initializer=lambda: None,
    trainable_variables=lambda: list(_TRAINABLE),
    set_random_seed=lambda s: _RNG.manual_seed(int(s)),
    disable_eager_execution=lambda: None,
    train=types.SimpleNamespace(AdamOptimizer=_Adam),
)
compat = types.SimpleNamespace(v1=v1)
keras = types.SimpleNamespace(backend=types.SimpleNamespace(clear_session=lambda: None))


def reset_shim(seed=0, keep_overrides=False):
    _TRAINABLE.clear()
    NOISE_LOG.clear()
    UNIFORM_LOG.clear()
    _VAR_COUNT[0] = 0
    if not keep_overrides:
        VARIABLE_OVERRIDES.clear()
    _RNG.manual_seed(seed)


# tf.Tensor.get_shape().ndims (conditionals_multi_output.py:368): the shim's tensors are torch tensors
class _ShapeView:
    def __init__(self, t):
        self.ndims = t.dim()
        self._s = tuple(t.shape)

    def __getitem__(self, i):
        return self._s[i]

    def __len__(self):
        return len(self._s)

    def as_list(self):
        return list(self._s)


torch.Tensor.get_shape = lambda self: _ShapeView(self)
