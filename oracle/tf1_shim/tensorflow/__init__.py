"""A tiny eager stand-in for the TensorFlow 1.x API surface that bzhangGo/zero's Transformer path touches.

TEST INFRASTRUCTURE ONLY.  TensorFlow <= 1.13 / Python 2.7 cannot be installed in this container, so the
reference cannot be *run* as shipped.  This package lets the reference's own, unmodified Python files
(func.py, models/transformer*.py, modules/rpr.py, modules/rela.py, search.py, utils/util.py, utils/dtype.py)
be imported and executed: every `tf.*` call they make is answered here with the documented TF1 semantics,
evaluated eagerly on torch CPU tensors.  tests/golden/make_golden.py uses it to produce the golden vectors
that pin oracle/zero_oracle.py (the independent restatement used on the GPU box) to the reference's code.

Only the ops listed by `grep -oh "tf\\.[A-Za-z_.0-9]*"` over those files are provided.  Semantics restated
from the TF 1.13 API documentation:
  * tf.nn.top_k: values sorted descending, ties -> lower index first.
  * tf.where(cond[B], x[B,...], y[B,...]): row selection when cond is rank-1.
  * tf.nn.dropout(x, keep_prob): identity at keep_prob == 1 (all golden vectors use dropout 0).
  * variable scopes: name prefixing, reuse, inherited initializer / dtype / custom_getter.
Nothing under zero_b200/ imports this.
"""
from __future__ import annotations

import collections
import collections.abc
import contextlib
import math

import numpy as np
import torch

# the reference is Python-2 era code: utils/util.py:120 uses collections.Mapping
if not hasattr(collections, "Mapping"):
    collections.Mapping = collections.abc.Mapping

__version__ = "1.13.2-shim"
AUTO_REUSE = "AUTO_REUSE"


# ----------------------------------------------------------------------------------------------- dtypes
class DType(object):
    def __init__(self, name, tdtype):
        self.name = name
        self.t = tdtype

    @property
    def min(self):
        return float(torch.finfo(self.t).min) if self.t.is_floating_point else int(np.iinfo(self.name).min)

    @property
    def max(self):
        return float(torch.finfo(self.t).max) if self.t.is_floating_point else int(np.iinfo(self.name).max)

    def __eq__(self, other):
        return isinstance(other, DType) and other.t == self.t

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return "tf." + self.name


float16 = DType("float16", torch.float16)
float32 = DType("float32", torch.float32)
float64 = DType("float64", torch.float64)
int32 = DType("int32", torch.int64)  # ids/indices are carried as int64; values are identical
int64 = DType("int64", torch.int64)
bool = DType("bool", torch.bool)  # noqa: A001  (mirrors tf.bool)
_BY_NAME = {"float16": float16, "float32": float32, "float64": float64, "int32": int32, "int64": int64,
            "bool": bool}
_BY_TORCH = {torch.float16: float16, torch.float32: float32, torch.float64: float64, torch.int64: int64,
             torch.int32: int32, torch.bool: bool}


def as_dtype(d):
    if isinstance(d, DType):
        return d
    if isinstance(d, str):
        return _BY_NAME[d]
    if isinstance(d, torch.dtype):
        return _BY_TORCH[d]
    raise TypeError("unknown dtype %r" % (d,))


def _td(d, default=None):
    if d is None:
        return default
    return as_dtype(d).t


# ----------------------------------------------------------------------------------------------- tensors
class TensorShape(tuple):
    def __new__(cls, dims=()):
        return super().__new__(cls, tuple(None if d is None else int(d) for d in dims))

    @property
    def ndims(self):
        return len(self)

    @property
    def dims(self):
        return list(self)

    def as_list(self):
        return list(self)


class Tensor(torch.Tensor):
    """torch.Tensor with the handful of tf.Tensor attributes the reference uses."""

    def get_shape(self):
        return TensorShape(torch.Tensor.size(self))

    @property
    def shape(self):
        return TensorShape(torch.Tensor.size(self))

    def set_shape(self, shape):
        return None

    # TF tensors are immutable: `q *= s` in the reference rebinds, it never mutates a view
    def __imul__(self, o):
        return self * o

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __itruediv__(self, o):
        return self / o

    def __bool__(self):
        return builtins_bool(torch.Tensor.item(self))


import builtins as _b  # noqa: E402

builtins_bool = _b.bool


def _wrap(t):
    if isinstance(t, Tensor):
        return t
    return t.as_subclass(Tensor)


def convert_to_tensor(x, dtype=None, name=None):
    if isinstance(x, torch.Tensor):
        t = x
        if dtype is not None and t.dtype != _td(dtype):
            t = t.to(_td(dtype))
        return _wrap(t)
    if isinstance(x, (list, tuple)) and any(isinstance(e, torch.Tensor) for e in x):
        t = torch.stack([torch.as_tensor(e) for e in x])
    else:
        arr = np.asarray(x)
        if arr.dtype == np.float64 and dtype is None:
            arr = arr.astype(np.float32)
        t = torch.from_numpy(np.ascontiguousarray(arr)) if arr.ndim else torch.tensor(arr.item())
        if arr.ndim == 0 and arr.dtype.kind == "f" and dtype is None:
            t = t.to(torch.float32)
        if t.dtype == torch.int32:
            t = t.to(torch.int64)
    if dtype is not None:
        t = t.to(_td(dtype))
    return _wrap(t)


_c = convert_to_tensor


def _ints(shape):
    """Shape argument (list / tensor / mixture of ints and 0-d tensors) -> list of python ints."""
    if isinstance(shape, torch.Tensor):
        return [int(v) for v in shape.reshape(-1).tolist()]
    if isinstance(shape, (int, np.integer)):
        return [int(shape)]
    return [int(s) for s in shape]


def _axis_kw(axis, kw):
    if axis is None:
        axis = kw.get("reduction_indices", None)
    keep = kw.get("keepdims", kw.get("keep_dims", False))
    return axis, builtins_bool(keep)


# ----------------------------------------------------------------------------------------------- logging
class _Logging(object):
    INFO = 20

    def info(self, msg, *a):
        pass

    def warn(self, msg, *a):
        pass

    warning = warn

    def set_verbosity(self, v):
        pass


logging = _Logging()


class _GFile(object):
    @staticmethod
    def Exists(path):
        import os
        return os.path.exists(path)


gfile = _GFile()

# ----------------------------------------------------------------------------------------------- variables
_VARIABLES = collections.OrderedDict()
_RNG = torch.Generator().manual_seed(1234)


def reset_default_graph(seed=1234):
    _VARIABLES.clear()
    _RNG.manual_seed(seed)
    del _SCOPES[1:]


def set_random_seed(seed):
    _RNG.manual_seed(seed)


def all_variables():
    return _VARIABLES


def trainable_variables():
    return list(_VARIABLES.values())


class _Scope(object):
    def __init__(self, name, reuse, initializer, dtype, custom_getter):
        self.name, self.reuse, self.initializer, self.dtype, self.custom_getter = \
            name, reuse, initializer, dtype, custom_getter


_SCOPES = [_Scope("", None, None, float32, None)]


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, values=None, initializer=None, reuse=None, dtype=None,
                   custom_getter=None, **_unused):
    parent = _SCOPES[-1]
    name = name_or_scope if name_or_scope is not None else default_name
    full = name if not parent.name else parent.name + "/" + name
    sc = _Scope(full,
                reuse if reuse is not None else parent.reuse,
                initializer if initializer is not None else parent.initializer,
                dtype if dtype is not None else parent.dtype,
                custom_getter if custom_getter is not None else parent.custom_getter)
    _SCOPES.append(sc)
    try:
        yield sc
    finally:
        _SCOPES.pop()


@contextlib.contextmanager
def name_scope(name, default_name=None, values=None):
    yield name or default_name


def get_variable_scope():
    return _SCOPES[-1]


def _true_getter(name, shape=None, dtype=None, initializer=None, regularizer=None, trainable=True, **_kw):
    if name in _VARIABLES:
        return _VARIABLES[name]
    dt = as_dtype(dtype) if dtype is not None else float32
    if initializer is None:
        initializer = glorot_uniform_initializer()
    val = initializer(_ints(shape), dtype=dt)
    val = torch.as_tensor(val).detach().to(dt.t).clone()
    var = _wrap(val)
    var.requires_grad_(builtins_bool(trainable) and dt.t.is_floating_point)
    var.var_name = name
    _VARIABLES[name] = var
    return var


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **_kw):
    sc = _SCOPES[-1]
    full = name if not sc.name else sc.name + "/" + name
    if initializer is None:
        initializer = sc.initializer
    if dtype is None:
        dtype = sc.dtype
    if sc.custom_getter is not None:
        return sc.custom_getter(_true_getter, full, shape=shape, dtype=as_dtype(dtype), initializer=initializer,
                                trainable=trainable)
    return _true_getter(full, shape=shape, dtype=dtype, initializer=initializer, trainable=trainable)


# ----------------------------------------------------------------------------------------------- initializers
def _fans(shape):
    if len(shape) < 1:
        return 1.0, 1.0
    if len(shape) == 1:
        return float(shape[0]), float(shape[0])
    if len(shape) == 2:
        return float(shape[0]), float(shape[1])
    rf = float(np.prod(shape[:-2]))
    return shape[-2] * rf, shape[-1] * rf


def zeros_initializer(dtype=None):
    return lambda shape, dtype=float32, partition_info=None: torch.zeros(_ints(shape), dtype=_td(dtype))


def ones_initializer(dtype=None):
    return lambda shape, dtype=float32, partition_info=None: torch.ones(_ints(shape), dtype=_td(dtype))


def random_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=None):
    def init(shape, dtype=float32, partition_info=None):
        return (torch.randn(_ints(shape), generator=_RNG, dtype=torch.float64) * stddev + mean).to(_td(dtype))
    return init


def random_uniform_initializer(minval=0.0, maxval=None, seed=None, dtype=None):
    def init(shape, dtype=float32, partition_info=None):
        hi = 1.0 if maxval is None else maxval
        return (torch.rand(_ints(shape), generator=_RNG, dtype=torch.float64) * (hi - minval) + minval).to(_td(dtype))
    return init


def variance_scaling_initializer(scale=1.0, mode="fan_in", distribution="truncated_normal", seed=None, dtype=None):
    def init(shape, dtype=float32, partition_info=None):
        shape_ = _ints(shape)
        fan_in, fan_out = _fans(shape_)
        s = scale
        if mode == "fan_in":
            s /= max(1.0, fan_in)
        elif mode == "fan_out":
            s /= max(1.0, fan_out)
        else:
            s /= max(1.0, (fan_in + fan_out) / 2.0)
        if distribution == "uniform":
            limit = math.sqrt(3.0 * s)
            return ((torch.rand(shape_, generator=_RNG, dtype=torch.float64) * 2 - 1) * limit).to(_td(dtype))
        std = math.sqrt(s)
        if distribution in ("normal", "truncated_normal"):
            # TF 1.13: "normal" is an alias of "truncated_normal" — a normal truncated at two standard deviations
            # (out-of-range draws are redrawn), its stddev divided by 0.8796... so that the variance stays `s`
            std /= 0.87962566103423978
            z = torch.randn(shape_, generator=_RNG, dtype=torch.float64)
            bad = z.abs() > 2.0
            while builtins_bool(bad.any()):
                z[bad] = torch.randn(int(bad.sum()), generator=_RNG, dtype=torch.float64)
                bad = z.abs() > 2.0
            return (z * std).to(_td(dtype))
        return (torch.randn(shape_, generator=_RNG, dtype=torch.float64) * std).to(_td(dtype))
    return init


def glorot_uniform_initializer(seed=None, dtype=None):
    return variance_scaling_initializer(1.0, "fan_avg", "uniform")


# ----------------------------------------------------------------------------------------------- creation
def constant(value, dtype=None, shape=None, name=None):
    t = _c(value, dtype)
    if shape is not None:
        t = _wrap(t.expand(_ints(shape)).clone())
    return t


def zeros(shape, dtype=float32, name=None):
    return _wrap(torch.zeros(_ints(shape), dtype=_td(dtype)))


def ones(shape, dtype=float32, name=None):
    return _wrap(torch.ones(_ints(shape), dtype=_td(dtype)))


def zeros_like(x, dtype=None):
    return _wrap(torch.zeros_like(_c(x), dtype=_td(dtype)))


def ones_like(x, dtype=None):
    return _wrap(torch.ones_like(_c(x), dtype=_td(dtype)))


def fill(dims, value):
    v = _c(value)
    return _wrap(torch.full(_ints(dims), v.item(), dtype=v.dtype))


def range(start, limit=None, delta=1, dtype=None, name=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    return _wrap(torch.arange(int(start), int(limit), int(delta), dtype=_td(dtype, torch.int64)))


def eye(n, dtype=float32):
    return _wrap(torch.eye(int(n), dtype=_td(dtype)))


def one_hot(indices, depth, on_value=None, off_value=None, axis=None, dtype=None, name=None):
    idx = _c(indices).to(torch.int64)
    depth = int(depth)
    if dtype is not None:
        dt = _td(dtype)
    elif on_value is not None:
        dt = _c(on_value).dtype
    else:
        dt = torch.float32
    on = 1.0 if on_value is None else float(_c(on_value))
    off = 0.0 if off_value is None else float(_c(off_value))
    oh = torch.nn.functional.one_hot(idx, depth).to(torch.bool)
    return _wrap(torch.where(oh, torch.tensor(on, dtype=dt), torch.tensor(off, dtype=dt)))


def random_uniform(shape, minval=0, maxval=None, dtype=float32, seed=None):
    hi = 1.0 if maxval is None else maxval
    return _wrap((torch.rand(_ints(shape), generator=_RNG) * (hi - minval) + minval).to(_td(dtype)))


# ----------------------------------------------------------------------------------------------- shape ops
def shape(x, name=None):  # noqa: A001
    return _wrap(torch.tensor(list(torch.Tensor.size(_c(x))), dtype=torch.int64))


def reshape(x, shape_, name=None):
    return _wrap(torch.reshape(_c(x), _ints(shape_)))


def expand_dims(x, axis=None, name=None, dim=None):
    return _wrap(torch.unsqueeze(_c(x), axis if axis is not None else dim))


def squeeze(x, axis=None, name=None):
    x = _c(x)
    return _wrap(torch.squeeze(x) if axis is None else torch.squeeze(x, axis))


def transpose(x, perm=None, name=None):
    x = _c(x)
    if perm is None:
        perm = list(_b.range(x.dim()))[::-1]
    return _wrap(x.permute(*_ints(perm)))


def concat(values, axis, name=None):
    ts = [_c(v) for v in values]
    if all(not t.dtype.is_floating_point for t in ts):
        ts = [t.to(torch.int64) for t in ts]
    return _wrap(torch.cat(ts, dim=int(axis)))


def stack(values, axis=0, name=None):
    return _wrap(torch.stack([_c(v) for v in values], dim=axis))


def split(value, num_or_size_splits, axis=0, name=None):
    value = _c(value)
    if isinstance(num_or_size_splits, int):
        return [_wrap(t.contiguous()) for t in torch.chunk(value, num_or_size_splits, dim=axis)]
    return [_wrap(t.contiguous()) for t in torch.split(value, _ints(num_or_size_splits), dim=axis)]


def tile(x, multiples, name=None):
    return _wrap(_c(x).repeat(*_ints(multiples)))


def pad(x, paddings, mode="CONSTANT", name=None, constant_values=0):
    x = _c(x)
    p = [list(_ints(pp)) for pp in paddings]
    flat = []
    for lo, hi in reversed(p):
        flat += [lo, hi]
    return _wrap(torch.nn.functional.pad(x, flat, value=constant_values))


def gather(params, indices, axis=0, name=None):
    return _wrap(_c(params)[_c(indices).to(torch.int64)])


def gather_nd(params, indices, name=None):
    params, indices = _c(params), _c(indices).to(torch.int64)
    n = indices.shape[-1]
    return _wrap(params[tuple(indices[..., i] for i in _b.range(n))])


def boolean_mask(tensor, mask, name=None, axis=None):
    tensor, mask = _c(tensor), _c(mask).to(torch.bool)
    axis = 0 if axis is None else axis
    idx = torch.nonzero(mask, as_tuple=False).reshape(-1)
    return _wrap(torch.index_select(tensor, axis, idx))


def where(condition, x=None, y=None, name=None):
    c = _c(condition).to(torch.bool)
    x, y = _c(x), _c(y)
    if c.dim() == 1 and x.dim() > 1:
        c = c.reshape([-1] + [1] * (x.dim() - 1))
    return _wrap(torch.where(c, x, y))


def matrix_band_part(x, num_lower, num_upper, name=None):
    x = _c(x)
    m, n = x.shape[-2], x.shape[-1]
    i = torch.arange(m).reshape(-1, 1)
    j = torch.arange(n).reshape(1, -1)
    keep = torch.ones(m, n, dtype=torch.bool)
    if num_lower >= 0:
        keep &= (i - j) <= num_lower
    if num_upper >= 0:
        keep &= (j - i) <= num_upper
    return _wrap(x * keep.to(x.dtype))


# ----------------------------------------------------------------------------------------------- math
def cast(x, dtype, name=None):
    t = _c(x)
    return _wrap(t.to(_td(dtype)))


def to_float(x, name=None):
    return cast(x, float32)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _c(a), _c(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return _wrap(torch.matmul(a, b))


def add_n(inputs, name=None):
    out = _c(inputs[0])
    for t in inputs[1:]:
        out = out + _c(t)
    return _wrap(out)


def _unary(fn):
    return lambda x, name=None: _wrap(fn(_c(x)))


exp = _unary(torch.exp)
log = _unary(torch.log)
sin = _unary(torch.sin)
cos = _unary(torch.cos)
tanh = _unary(torch.tanh)
sigmoid = _unary(torch.sigmoid)
rsqrt = _unary(torch.rsqrt)
sqrt = _unary(torch.sqrt)
logical_not = _unary(torch.logical_not)


def pow(x, y, name=None):  # noqa: A001
    return _wrap(torch.pow(_c(x), _c(y) if isinstance(y, torch.Tensor) else y))


def mod(x, y, name=None):
    if isinstance(x, torch.Tensor) or isinstance(y, torch.Tensor):
        return _wrap(torch.remainder(_c(x), _c(y)))
    return x % y


def clip_by_value(x, lo, hi, name=None):
    return _wrap(torch.clamp(_c(x), lo, hi))


def cumsum(x, axis=0, name=None):
    return _wrap(torch.cumsum(_c(x), dim=axis))


def _cmp(fn):
    return lambda x, y, name=None: _wrap(fn(_c(x), _c(y)))


equal = _cmp(torch.eq)
less = _cmp(torch.lt)
less_equal = _cmp(torch.le)
greater = _cmp(torch.gt)
greater_equal = _cmp(torch.ge)
logical_and = _cmp(torch.logical_and)
logical_or = _cmp(torch.logical_or)


def _reduce(fn_all, fn_axis):
    def red(x, axis=None, keepdims=None, name=None, **kw):
        if keepdims is not None:
            kw["keepdims"] = keepdims
        axis, keep = _axis_kw(axis, kw)
        x = _c(x)
        if axis is None:
            return _wrap(fn_all(x))
        return _wrap(fn_axis(x, axis, keep))
    return red


reduce_sum = _reduce(torch.sum, lambda x, a, k: torch.sum(x, dim=a, keepdim=k))
reduce_mean = _reduce(torch.mean, lambda x, a, k: torch.mean(x, dim=a, keepdim=k))
reduce_min = _reduce(torch.min, lambda x, a, k: torch.min(x, dim=a, keepdim=k).values)
reduce_max = _reduce(torch.max, lambda x, a, k: torch.max(x, dim=a, keepdim=k).values)
reduce_all = _reduce(lambda x: torch.all(x.to(torch.bool)), lambda x, a, k: torch.all(x.to(torch.bool), dim=a, keepdim=k))
reduce_any = _reduce(lambda x: torch.any(x.to(torch.bool)), lambda x, a, k: torch.any(x.to(torch.bool), dim=a, keepdim=k))
reduce_logsumexp = _reduce(lambda x: torch.logsumexp(x.reshape(-1), 0),
                           lambda x, a, k: torch.logsumexp(x, dim=a, keepdim=k))


# ----------------------------------------------------------------------------------------------- control flow
def cond(pred, true_fn=None, false_fn=None, name=None):
    return true_fn() if builtins_bool(_c(pred).item()) else false_fn()


def while_loop(cond, body, loop_vars, shape_invariants=None, parallel_iterations=10, back_prop=True, **_kw):  # noqa: A002
    lv = list(loop_vars)
    with torch.no_grad() if not back_prop else contextlib.nullcontext():
        while builtins_bool(_c(cond(*lv)).item()):
            lv = list(body(*lv))
    return lv


# ----------------------------------------------------------------------------------------------- tf.nn
class _NN(object):
    @staticmethod
    def softmax(logits, axis=-1, name=None, dim=None):
        if dim is not None:
            axis = dim
        return _wrap(torch.softmax(_c(logits), dim=axis))

    @staticmethod
    def relu(x, name=None):
        return _wrap(torch.relu(_c(x)))

    @staticmethod
    def sigmoid(x, name=None):
        return _wrap(torch.sigmoid(_c(x)))

    @staticmethod
    def bias_add(value, bias, name=None):
        return _wrap(_c(value) + _c(bias))

    @staticmethod
    def dropout(x, keep_prob, name=None):
        kp = float(keep_prob)
        if kp >= 1.0:
            return _c(x)
        x = _c(x)
        keep = (torch.rand(x.shape, generator=_RNG) < kp).to(x.dtype)
        return _wrap(x * keep / kp)

    @staticmethod
    def top_k(x, k=1, sorted=True, name=None):  # noqa: A002
        x = _c(x)
        vals, idx = torch.sort(x, dim=-1, descending=True, stable=True)
        k = int(k)
        return _wrap(vals[..., :k].contiguous()), _wrap(idx[..., :k].contiguous())

    @staticmethod
    def softmax_cross_entropy_with_logits_v2(labels=None, logits=None, name=None, **_kw):
        logits, labels = _c(logits), _c(labels)
        lsm = torch.log_softmax(logits, dim=-1)
        return _wrap(-(labels * lsm).sum(-1))


nn = _NN()


class ConfigProto(object):  # referenced by utils/util.get_session only
    def __init__(self, **kw):
        pass


class Session(object):
    def __init__(self, **kw):
        raise RuntimeError("tf.Session is not available in the TF1 shim (eager only)")
