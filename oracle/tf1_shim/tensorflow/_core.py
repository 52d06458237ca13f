"""TEST INFRASTRUCTURE — a minimal eager stand-in for the TensorFlow r1.4 Python API.

Why it exists: the reference (``/root/reference/models/*.py``) is TF-1.x graph code and TensorFlow cannot be installed
here.  To pin the CPU oracle (``oracle/tacotron_oracle.py``) against *the reference's own model-building code* rather
than against a second reading of it, ``tools/make_reference_golden.py`` imports the UNMODIFIED reference modules with
this package on ``sys.path`` as ``tensorflow``.  Every ``tf.*`` call the reference makes on the hot path lands here and
is evaluated eagerly on torch CPU tensors (autograd gives ``optimizer.compute_gradients``).

What this pins and what it does not:
  * pinned by the reference's code, executed as is: which layers exist, their sizes, order, wiring, scopes, the
    decoder/attention wrapper stack (rnn_wrappers.py), the helpers' feeding rules (helpers.py), the loss and the
    optimizer recipe (tacotron.py:274-336), hyper-parameter defaults (hparams.py);
  * restated here from the TF r1.4 sources (file names cited per function, from memory — TF is not in this image):
    the *library* semantics — GRUCell, OutputProjectionWrapper, ResidualWrapper, MultiRNNCell, dynamic_decode,
    BasicDecoder, Bahdanau(Monotonic)Attention, bidirectional_dynamic_rnn, tf.layers.{dense,conv1d,
    batch_normalization,max_pooling1d,dropout}, clip_by_global_norm, AdamOptimizer.
The shim is written independently of the oracle (it shares no code with it), so agreement of the two is a real check
of both readings of TF; it is still not TensorFlow, and DESIGN.md §4 says so.

Graph-vs-eager: TF traces a ``while_loop`` body once; here the body runs once per step.  Variable creation is therefore
get-or-create, and every object that owns a default-named scope (Layer / RNNCell / attention mechanism) captures it on
first use — the variable names that result are the ones TF would produce.

Only ``tools/make_reference_golden.py`` and ``tests/`` may import this package.
"""
from __future__ import annotations

import collections
import contextlib
import math
import re
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------------------------
# dtypes, shapes, tensors
# ------------------------------------------------------------------------------------------------------------------
class DType:
    def __init__(self, name, tdtype):
        self.name, self.torch = name, tdtype

    def __repr__(self):
        return "tf." + self.name


float32 = DType("float32", torch.float32)
float64 = DType("float64", torch.float64)
int32 = DType("int32", torch.int32)
int64 = DType("int64", torch.int64)
bool_ = DType("bool", torch.bool)


def set_float_precision(double: bool):
    """Evaluate everything the reference declares as tf.float32 in float64 (for rounding-free structural comparisons)."""
    float32.torch = torch.float64 if double else torch.float32


def _dtype_of(t: torch.Tensor) -> DType:
    if t.dtype in (torch.float32, torch.float64):
        return float32 if t.dtype == float32.torch else float64
    if t.dtype == torch.bool:
        return bool_
    return int32 if t.dtype == torch.int32 else int64


class Dimension:
    def __init__(self, value):
        self.value = None if value is None else int(value)

    def __int__(self):
        return self.value

    __index__ = __int__

    def __eq__(self, other):
        return self.value == (other.value if isinstance(other, Dimension) else other)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return "Dimension(%r)" % self.value

    def __str__(self):
        return str(self.value)


class TensorShape:
    def __init__(self, dims):
        if isinstance(dims, TensorShape):
            dims = dims.as_list()
        self._dims = [d if isinstance(d, Dimension) else Dimension(d) for d in dims]

    @property
    def ndims(self):
        return len(self._dims)

    @property
    def dims(self):
        return self._dims

    def __len__(self):
        return len(self._dims)

    def __iter__(self):
        return iter(self._dims)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return TensorShape(self._dims[i])
        return self._dims[i]

    def as_list(self):
        return [d.value for d in self._dims]

    def __eq__(self, other):
        return self.as_list() == TensorShape(other).as_list()

    def __repr__(self):
        return "TensorShape(%r)" % self.as_list()


def _scalar_int(x) -> int:
    if isinstance(x, Tensor):
        return int(x.t.item())
    if isinstance(x, Dimension):
        return x.value
    if isinstance(x, torch.Tensor):
        return int(x.item())
    return int(x)


def _ints(xs) -> List[int]:
    if isinstance(xs, Tensor):
        return [int(v) for v in xs.t.reshape(-1).tolist()]
    if isinstance(xs, TensorShape):
        return xs.as_list()
    if isinstance(xs, (list, tuple)):
        return [_scalar_int(v) for v in xs]
    return [_scalar_int(xs)]


def _c(x, dtype: Optional[DType] = None) -> torch.Tensor:
    """Anything the reference passes where TF expects a tensor -> torch tensor."""
    if isinstance(x, Tensor):
        t = x.t
    elif isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(x)
        if t.dtype in (torch.float32, torch.float64):
            t = t.to(float32.torch)
    elif isinstance(x, (list, tuple)) and any(isinstance(v, (Tensor, torch.Tensor)) for v in x):
        t = torch.stack([_c(v) for v in x])
    else:
        probe = np.asarray(x)
        if probe.dtype == np.bool_:
            t = torch.tensor(probe, dtype=torch.bool)
        elif np.issubdtype(probe.dtype, np.integer):
            t = torch.tensor(probe, dtype=torch.int32)
        else:
            t = torch.tensor(probe, dtype=float32.torch)
    if dtype is not None:
        t = t.to(dtype.torch)
    return t


def _promote(a, b):
    """Python scalars adopt the tensor operand's dtype (as TF's op overloads do)."""
    ta = a.t if isinstance(a, Tensor) else None
    tb = b.t if isinstance(b, Tensor) else None
    if ta is None:
        ta = torch.as_tensor(a, dtype=tb.dtype) if not isinstance(a, (torch.Tensor, np.ndarray, list, tuple)) else _c(a).to(tb.dtype)
    if tb is None:
        tb = torch.as_tensor(b, dtype=ta.dtype) if not isinstance(b, (torch.Tensor, np.ndarray, list, tuple)) else _c(b).to(ta.dtype)
    return ta, tb


class Tensor:
    __array_priority__ = 100

    def __init__(self, t: torch.Tensor, name: Optional[str] = None):
        self.t = t
        self.name = name

    # -- static information ------------------------------------------------
    @property
    def shape(self):
        return TensorShape(list(self.t.shape))

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return _dtype_of(self.t)

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __repr__(self):
        return "<shim.Tensor %s %s %s>" % (self.name or "", tuple(self.t.shape), self.t.dtype)

    def __bool__(self):
        return bool(self.t.item())

    def __int__(self):
        return int(self.t.item())

    __index__ = __int__

    def __float__(self):
        return float(self.t.item())

    # -- operators ---------------------------------------------------------
    def _bin(self, other, fn, swap=False):
        a, b = _promote(self, other)
        return Tensor(fn(b, a) if swap else fn(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __rpow__(self, o): return self._bin(o, torch.pow, True)
    def __neg__(self): return Tensor(-self.t)
    def __ge__(self, o): return self._bin(o, torch.ge)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __le__(self, o): return self._bin(o, torch.le)
    def __lt__(self, o): return self._bin(o, torch.lt)
    __hash__ = object.__hash__

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        out = []
        for i in idx:
            if isinstance(i, (Tensor, Dimension)):
                out.append(_scalar_int(i))
            elif isinstance(i, slice):
                out.append(slice(*[None if v is None else _scalar_int(v) for v in (i.start, i.stop, i.step)]))
            else:
                out.append(i)
        return Tensor(self.t[tuple(out)])


class Variable(Tensor):
    """tf.Variable / the result of tf.get_variable: a named, mutable leaf."""

    def __init__(self, initial_value, name=None, trainable=True, dtype=None):
        t = _c(initial_value, dtype).clone()
        if trainable and t.dtype.is_floating_point:
            t.requires_grad_(True)
        super().__init__(t, name)
        self.trainable = trainable

    def assign(self, value):
        with torch.no_grad():
            self.t.copy_(_c(value).to(self.t.dtype))
        return self


# ------------------------------------------------------------------------------------------------------------------
# graph state: variable store, scopes, collections, placeholders
# ------------------------------------------------------------------------------------------------------------------
class GraphKeys:
    UPDATE_OPS = "update_ops"
    TRAINABLE_VARIABLES = "trainable_variables"
    GLOBAL_VARIABLES = "variables"


class _State:
    def __init__(self):
        self.reset()

    def reset(self, provider: Optional[Callable] = None, feeds: Optional[Dict[str, object]] = None):
        self.scope: List[str] = []
        self.vars: "collections.OrderedDict[str, Variable]" = collections.OrderedDict()
        self.opened: Dict[str, int] = collections.defaultdict(int)   # variable_scope_count of python/ops/variable_scope.py
        self.collections: Dict[str, list] = collections.defaultdict(list)
        self.provider = provider
        self.feeds = dict(feeds or {})
        self.placeholders: Dict[str, Tensor] = {}


_S = _State()


def reset_default_graph(provider=None, feeds=None):
    """provider(full_variable_name, shape, initializer) -> torch tensor | None (None: use the initializer)."""
    _S.reset(provider, feeds)


def shim_state() -> _State:
    return _S


class VariableScope:
    def __init__(self, path: List[str]):
        self.path = list(path)

    @property
    def name(self):
        return "/".join(self.path)


def get_variable_scope():
    return VariableScope(_S.scope)


def _unique_scope(prefix: str) -> str:
    """variable_scope._get_unique_variable_scope: first of prefix, prefix_1, ... never opened under the current scope."""
    base = "/".join(_S.scope + [prefix])
    if _S.opened[base] == 0:
        return prefix
    i = 1
    while _S.opened[base + "_%d" % i] > 0:
        i += 1
    return prefix + "_%d" % i


@contextlib.contextmanager
def variable_scope(name_or_scope=None, default_name=None, values=None, reuse=None, **_kw):
    saved = list(_S.scope)
    if isinstance(name_or_scope, VariableScope):
        _S.scope = list(name_or_scope.path)
    else:
        name = name_or_scope if name_or_scope is not None else _unique_scope(default_name)
        _S.scope = saved + [p for p in name.split("/") if p]
    _S.opened["/".join(_S.scope)] += 1
    try:
        yield VariableScope(_S.scope)
    finally:
        _S.scope = saved


@contextlib.contextmanager
def name_scope(name=None, default_name=None, values=None):
    yield name or default_name


@contextlib.contextmanager
def control_dependencies(ops):
    """Eager stand-in: pending update closures (batch-norm moving averages) run when the dependency is declared."""
    for op in ops or []:
        if callable(op):
            op()
    yield


def get_collection(key):
    return list(_S.collections[key])


def add_to_collection(key, value):
    _S.collections[key].append(value)


def trainable_variables():
    return [v for v in _S.vars.values() if v.trainable]


def global_variables():
    return list(_S.vars.values())


# initializers ------------------------------------------------------------------------------------------------------
class _Init:
    def __init__(self, kind, **kw):
        self.kind, self.kw = kind, kw

    def __call__(self, shape, dtype):
        shape = list(shape)
        if self.kind == "constant":
            return torch.full(shape, float(self.kw["value"]), dtype=dtype)
        if self.kind == "zeros":
            return torch.zeros(shape, dtype=dtype)
        if self.kind == "ones":
            return torch.ones(shape, dtype=dtype)
        g = self.kw.setdefault("_gen", torch.Generator().manual_seed(self.kw.get("seed") or 0))
        if self.kind == "truncated_normal":
            t = torch.randn(shape, generator=g, dtype=torch.float64)
            for _ in range(32):
                bad = t.abs() > 2
                if not bad.any():
                    break
                t = torch.where(bad, torch.randn(shape, generator=g, dtype=torch.float64), t)
            return (t * self.kw["stddev"] + self.kw.get("mean", 0.0)).to(dtype)
        if self.kind == "glorot_uniform":
            if len(shape) == 1:
                fi = fo = shape[0]
            else:
                rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
                fi, fo = shape[-2] * rf, shape[-1] * rf
            lim = math.sqrt(6.0 / (fi + fo))
            return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(dtype)
        raise ValueError(self.kind)


def truncated_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=None):
    return _Init("truncated_normal", mean=mean, stddev=stddev, seed=seed)


def constant_initializer(value=0, dtype=None):
    return _Init("constant", value=value)


def zeros_initializer(dtype=None):
    return _Init("zeros")


def ones_initializer(dtype=None):
    return _Init("ones")


def glorot_uniform_initializer(seed=None, dtype=None):
    return _Init("glorot_uniform", seed=seed)


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **_kw):
    """Get-or-create (see module docstring).  Default initializer: glorot_uniform (variable_scope.py, float dtypes)."""
    full = "/".join(_S.scope + [name])
    if full in _S.vars:
        return _S.vars[full]
    dtype = dtype or float32
    value = None
    if shape is None and initializer is not None and not callable(initializer):
        value = _c(initializer, dtype)                       # initializer given as a value (attention_g, score bias)
        shape = list(value.shape)
    shape = _ints(shape) if shape is not None else []
    if _S.provider is not None:
        got = _S.provider(full, tuple(shape), initializer)
        if got is not None:
            value = torch.as_tensor(got).to(dtype.torch).reshape(shape)
    if value is None:
        init = initializer if callable(initializer) else glorot_uniform_initializer()
        value = init(shape, dtype.torch)
    v = Variable(value, name=full, trainable=trainable)
    _S.vars[full] = v
    return v


def placeholder(dtype, shape=None, name=None):
    """Fed at construction time from ``reset_default_graph(feeds={name: value})`` (eager: there is no later session.run)."""
    if name in _S.feeds:
        t = _c(_S.feeds[name], dtype)
    elif dtype is bool_:
        t = torch.tensor(False)
    else:
        t = torch.zeros([1 if d is None else d for d in (shape or [])], dtype=dtype.torch)
    out = Tensor(t, name)
    _S.placeholders[name] = out
    return out


# ------------------------------------------------------------------------------------------------------------------
# ops used by the reference (python/ops/array_ops.py, math_ops.py, nn_ops.py)
# ------------------------------------------------------------------------------------------------------------------
def shape(x, name=None):
    return Tensor(torch.tensor(list(_c(x).shape), dtype=torch.int32))


def tile(x, multiples, name=None):
    return Tensor(_c(x).repeat(*_ints(multiples)))


def concat(values, axis, name=None):
    return Tensor(torch.cat([_c(v) for v in values], dim=_scalar_int(axis)))


def expand_dims(x, axis=None, name=None, dim=None):
    ax = axis if axis is not None else dim
    if isinstance(ax, (list, tuple)):
        ax = ax[0]
    return Tensor(_c(x).unsqueeze(_scalar_int(ax)))


def squeeze(x, axis=None, name=None):
    t = _c(x)
    for ax in sorted(_ints(axis), reverse=True) if axis is not None else []:
        t = t.squeeze(ax)
    return Tensor(t if axis is not None else t.squeeze())


def reshape(x, shp, name=None):
    return Tensor(_c(x).reshape(_ints(shp)))


def transpose(x, perm=None, name=None):
    t = _c(x)
    return Tensor(t.permute(*_ints(perm)) if perm is not None else t.permute(*reversed(range(t.dim()))))


def split(value, num_or_size_splits, axis=0, name=None):
    t = _c(value)
    ax = _scalar_int(axis)
    if isinstance(num_or_size_splits, int):
        return [Tensor(p) for p in torch.chunk(t, num_or_size_splits, dim=ax)]
    return [Tensor(p) for p in torch.split(t, _ints(num_or_size_splits), dim=ax)]


def identity(x, name=None):
    return x if isinstance(x, Tensor) else Tensor(_c(x))


def zeros(shp, dtype=float32, name=None):
    return Tensor(torch.zeros(_ints(shp) if not (isinstance(shp, (list, tuple)) and len(shp) == 0) else [], dtype=dtype.torch))


def ones(shp, dtype=float32, name=None):
    return Tensor(torch.ones(_ints(shp), dtype=dtype.torch))


def zeros_like(x):
    return Tensor(torch.zeros_like(_c(x)))


def fill(dims, value):
    return Tensor(torch.full(_ints(dims), value))


def cast(x, dtype, name=None):
    return Tensor(_c(x).to(dtype.torch))


def one_hot(indices, depth, dtype=float32, **_kw):
    return Tensor(F.one_hot(_c(indices).long(), _scalar_int(depth)).to(dtype.torch))


def reduce_mean(x, axis=None, keep_dims=False, name=None):
    t = _c(x)
    return Tensor(t.mean() if axis is None else t.mean(dim=_ints(axis), keepdim=keep_dims))


def reduce_sum(x, axis=None, keep_dims=False, name=None):
    t = _c(x)
    return Tensor(t.sum() if axis is None else t.sum(dim=_ints(axis), keepdim=keep_dims))


def reduce_all(x, axis=None, name=None):
    t = _c(x)
    return Tensor(t.all() if axis is None else t.all(dim=_scalar_int(axis)))


def equal(a, b, name=None):
    ta, tb = _promote(a if isinstance(a, Tensor) else Tensor(_c(a)), b)
    return Tensor(torch.eq(ta, tb))


def logical_or(a, b):
    return Tensor(torch.logical_or(_c(a), _c(b)))


def logical_and(a, b):
    return Tensor(torch.logical_and(_c(a), _c(b)))


def logical_not(a):
    return Tensor(torch.logical_not(_c(a)))


def abs(x, name=None):  # noqa: A001 (the reference calls tf.abs)
    return Tensor(_c(x).abs())


def minimum(a, b, name=None):
    ta, tb = _promote(a if isinstance(a, Tensor) else Tensor(_c(a)), b)
    return Tensor(torch.minimum(ta, tb))


def maximum(a, b, name=None):
    ta, tb = _promote(a if isinstance(a, Tensor) else Tensor(_c(a)), b)
    return Tensor(torch.maximum(ta, tb))


def matmul(a, b, name=None):
    return Tensor(torch.matmul(_c(a), _c(b)))


def tanh(x):
    return Tensor(torch.tanh(_c(x)))


def sigmoid(x):
    return Tensor(torch.sigmoid(_c(x)))


def sqrt(x):
    return Tensor(torch.sqrt(_c(x)))


def rsqrt(x):
    return Tensor(torch.rsqrt(_c(x)))


def square(x):
    t = _c(x)
    return Tensor(t * t)


def exp(x):
    return Tensor(torch.exp(_c(x)))


def log(x):
    return Tensor(torch.log(_c(x)))


def cumsum(x, axis=0, exclusive=False, reverse=False):
    t = _c(x)
    ax = _scalar_int(axis)
    if reverse:
        t = t.flip(ax)
    out = torch.cumsum(t, ax)
    if exclusive:
        out = out - t
    return Tensor(out.flip(ax) if reverse else out)


def clip_by_value(x, lo, hi):
    return Tensor(torch.clamp(_c(x), lo, hi))


def where(cond, a, b):
    return Tensor(torch.where(_c(cond), _c(a), _c(b)))


def reverse_sequence(x, seq_lengths, seq_axis=1, batch_axis=0):
    t = _c(x)
    assert seq_axis == 1 and batch_axis == 0
    n, T = t.shape[:2]
    pos = torch.arange(T).unsqueeze(0).expand(n, T)
    L = _c(seq_lengths).long().view(n, 1)
    idx = torch.where(pos < L, L - 1 - pos, pos)
    while idx.dim() < t.dim():
        idx = idx.unsqueeze(-1)
    return Tensor(torch.gather(t, 1, idx.expand_as(t)))


def reverse(x, axis):
    return Tensor(_c(x).flip(_ints(axis)))


def cond(pred, true_fn=None, false_fn=None, fn1=None, fn2=None, name=None):
    """Eager: evaluates exactly one branch (the value of the predicate is known at construction, see placeholder())."""
    tf_, ff_ = true_fn or fn1, false_fn or fn2
    return tf_() if bool(_c(pred).item()) else ff_()


def assert_equal(a, b, message=None, **_kw):
    if _scalar_int(a) != _scalar_int(b):
        raise ValueError(message or "assert_equal failed")
    return None


def clip_by_global_norm(t_list, clip_norm, use_norm=None, name=None):
    """python/ops/clip_ops.py: global_norm = sqrt(sum ||t||^2); t * clip_norm * min(1/norm, 1/clip_norm)."""
    ts = [None if t is None else _c(t) for t in t_list]
    norm = torch.sqrt(sum((t.double() ** 2).sum() for t in ts if t is not None)).to(float32.torch)
    scale = clip_norm * torch.minimum(1.0 / norm, torch.tensor(1.0 / clip_norm, dtype=norm.dtype))
    return [None if t is None else Tensor(t * scale) for t in ts], Tensor(norm)


class TensorArray:
    """python/ops/tensor_array_ops.py — functional write/stack, all the reference uses (rnn_wrappers.py:212,284)."""

    def __init__(self, dtype=None, size=0, dynamic_size=False, _items=None, **_kw):
        self._items = dict(_items or {})

    def write(self, index, value):
        items = dict(self._items)
        items[_scalar_int(index)] = _c(value)
        return TensorArray(_items=items)

    def stack(self):
        return Tensor(torch.stack([self._items[i] for i in range(len(self._items))], 0))

    def size(self):
        return len(self._items)


# ------------------------------------------------------------------------------------------------------------------
# nest (python/util/nest.py)
# ------------------------------------------------------------------------------------------------------------------
def _is_seq(x):
    return isinstance(x, (list, tuple)) and not isinstance(x, (str, bytes))


def nest_flatten(s):
    if not _is_seq(s):
        return [s]
    out = []
    for v in s:
        out += nest_flatten(v)
    return out


def nest_map_structure(fn, *structs):
    s0 = structs[0]
    if not _is_seq(s0):
        return fn(*structs)
    mapped = [nest_map_structure(fn, *vs) for vs in zip(*structs)]
    if hasattr(s0, "_fields"):
        return type(s0)(*mapped)
    return type(s0)(mapped)


# ------------------------------------------------------------------------------------------------------------------
# tf.nn
# ------------------------------------------------------------------------------------------------------------------
def relu(x, name=None):
    return Tensor(torch.relu(_c(x)))


def softsign(x, name=None):
    t = _c(x)
    return Tensor(t / (t.abs() + 1))


def softmax(x, axis=-1, name=None):
    return Tensor(torch.softmax(_c(x), dim=axis))


def embedding_lookup(params, ids, name=None):
    return Tensor(_c(params)[_c(ids).long()])


# ------------------------------------------------------------------------------------------------------------------
# tf.layers (python/layers/base.py, core.py, convolutional.py, normalization.py, pooling.py)
# ------------------------------------------------------------------------------------------------------------------
def _to_snake_case(name):
    intermediate = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    insecure = re.sub("([a-z])([A-Z])", r"\1_\2", intermediate).lower()
    return insecure if insecure[0] != "_" else "private" + insecure


class Layer:
    """base.py Layer: owns one variable scope, captured at the first call.  Functional wrappers pass ``_scope=name`` when
    a name is given (no uniquification); unnamed / class-constructed layers take ``variable_scope(None, default_name)``."""

    def __init__(self, name=None, trainable=True, dtype=None, _scope=None, _reuse=None, **_kw):
        self._base_name = name or _to_snake_case(type(self).__name__)
        self._name = name
        self._fixed_scope = _scope
        self._scope: Optional[VariableScope] = None
        self.built = False

    @property
    def name(self):
        return self._scope.path[-1] if self._scope is not None else self._base_name

    @property
    def scope_name(self):
        return self._scope.name

    def _enter(self, scope=None):
        if self._scope is not None:
            return variable_scope(self._scope)
        if scope is not None:
            return variable_scope(scope)
        if self._fixed_scope is not None:
            return variable_scope(self._fixed_scope)
        return variable_scope(None, default_name=self._base_name)

    def __call__(self, *args, **kwargs):
        scope = kwargs.pop("scope", None)
        with self._enter(scope) as vs_:
            if self._scope is None:
                self._scope = vs_
            out = self.call(*args, **kwargs)
            self.built = True
            return out

    apply = __call__


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None, name=None, **kw):
        super().__init__(name=name, **kw)
        self.units, self.activation, self.use_bias = units, activation, use_bias
        self.kernel_initializer = kernel_initializer
        self.bias_initializer = bias_initializer or zeros_initializer()

    def call(self, inputs):
        x = _c(inputs)
        kernel = get_variable("kernel", [x.shape[-1], self.units], dtype=float32, initializer=self.kernel_initializer)
        y = torch.matmul(x, kernel.t)                                  # core.py: tensordot over the last axis for rank > 2
        if self.use_bias:
            y = y + get_variable("bias", [self.units], dtype=float32, initializer=self.bias_initializer).t
        out = Tensor(y)
        return self.activation(out) if self.activation is not None else out


def dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None, name=None, reuse=None, **_kw):
    return Dense(units, activation=activation, use_bias=use_bias, kernel_initializer=kernel_initializer,
                 bias_initializer=bias_initializer, name=name, _scope=name, _reuse=reuse)(inputs)


class Conv1D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None, use_bias=True, name=None, **kw):
        super().__init__(name=name, **kw)
        self.filters, self.k, self.padding, self.activation, self.use_bias = filters, int(kernel_size), padding.lower(), activation, use_bias
        assert strides == 1

    def call(self, inputs):
        x = _c(inputs)                                                 # [N, T, Cin]  (channels_last)
        kernel = get_variable("kernel", [self.k, x.shape[-1], self.filters], dtype=float32)
        xt = x.transpose(1, 2)
        if self.padding == "same":                                     # SAME: total k-1, the extra one goes to the END
            left = (self.k - 1) // 2
            xt = F.pad(xt, (left, self.k - 1 - left))
        y = F.conv1d(xt, kernel.t.permute(2, 1, 0))                    # cross-correlation, as TF
        y = y.transpose(1, 2)
        if self.use_bias:
            y = y + get_variable("bias", [self.filters], dtype=float32, initializer=zeros_initializer()).t
        out = Tensor(y)
        return self.activation(out) if self.activation is not None else out


def conv1d(inputs, filters, kernel_size, strides=1, padding="valid", activation=None, use_bias=True, name=None, **_kw):
    return Conv1D(filters, kernel_size, strides, padding, activation, use_bias, name=name, _scope=name)(inputs)


class BatchNormalization(Layer):
    """normalization.py, non-fused path (rank-3 input): biased batch moments over all but the last axis in training,
    moving statistics otherwise; moving <- moving - (moving - batch) * (1 - momentum), registered in UPDATE_OPS."""

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, name=None, **kw):
        super().__init__(name=name, **kw)
        self.momentum, self.epsilon = momentum, epsilon

    def call(self, inputs, training=False):
        x = _c(inputs)
        c = x.shape[-1]
        gamma = get_variable("gamma", [c], initializer=ones_initializer())
        beta = get_variable("beta", [c], initializer=zeros_initializer())
        mm = get_variable("moving_mean", [c], initializer=zeros_initializer(), trainable=False)
        mv = get_variable("moving_variance", [c], initializer=ones_initializer(), trainable=False)
        if training:
            red = list(range(x.dim() - 1))
            mean = x.mean(dim=red)
            var = ((x - mean) ** 2).mean(dim=red)                      # nn.moments: biased
            bm, bv, mom = mean.detach(), var.detach(), self.momentum
            done = [False]

            def update():
                if not done[0]:
                    done[0] = True
                    with torch.no_grad():
                        mm.t.sub_((mm.t - bm) * (1 - mom))
                        mv.t.sub_((mv.t - bv) * (1 - mom))
            add_to_collection(GraphKeys.UPDATE_OPS, update)
        else:
            mean, var = mm.t, mv.t
        inv = torch.rsqrt(var + self.epsilon) * gamma.t                # nn.batch_normalization
        return Tensor(x * inv + (beta.t - mean * inv))


def batch_normalization(inputs, axis=-1, momentum=0.99, epsilon=1e-3, training=False, name=None, **_kw):
    return BatchNormalization(axis, momentum, epsilon, name=name, _scope=name)(inputs, training=bool(training))


def max_pooling1d(inputs, pool_size, strides, padding="valid", name=None, **_kw):
    """pooling.py -> nn.max_pool; SAME pads (pool-1) with -inf, the extra one at the END.  Ties: the first maximum of the
    window owns the value (and the gradient), as the CPU MaxPool/MaxPoolGrad kernels do."""
    x = _c(inputs)
    p = int(pool_size)
    assert strides == 1
    T = x.shape[1]
    if padding.lower() == "same":
        left = (p - 1) // 2
        neg = torch.full_like(x[:, :1], -float("inf"))
        xp = torch.cat([neg] * left + [x] + [neg] * (p - 1 - left), dim=1)
        To = T
    else:
        xp, To = x, T - p + 1
    best = xp[:, 0:To]
    for j in range(1, p):
        cand = xp[:, j:j + To]
        best = torch.where(best >= cand, best, cand)
    return Tensor(best)


def dropout(inputs, rate=0.5, noise_shape=None, seed=None, training=False, name=None):
    """core.py Dropout: identity unless training=True — and the reference never passes it (modules.py:24)."""
    if training:
        raise NotImplementedError("the reference never enables dropout")
    return identity(inputs)


# ------------------------------------------------------------------------------------------------------------------
# RNN cells (python/ops/rnn_cell_impl.py, contrib/rnn/python/ops/core_rnn_cell.py)
# ------------------------------------------------------------------------------------------------------------------
def _zero_state_tensors(state_size, batch_size, dtype):
    def one(s):
        dims = _ints(s) if isinstance(s, (TensorShape, list, tuple)) else [_scalar_int(s)]
        return Tensor(torch.zeros([_scalar_int(batch_size)] + dims, dtype=dtype.torch))
    if isinstance(state_size, TensorShape):
        return one(state_size)
    return nest_map_structure(one, state_size)


class RNNCell(Layer):
    def __call__(self, inputs, state, scope=None):
        return Layer.__call__(self, inputs, state, **({"scope": scope} if scope is not None else {}))

    @property
    def state_size(self):
        raise NotImplementedError

    @property
    def output_size(self):
        raise NotImplementedError

    def zero_state(self, batch_size, dtype):
        return _zero_state_tensors(self.state_size, batch_size, dtype)


def _linear(args, output_size, bias, bias_initializer=None, kernel_initializer=None):
    """rnn_cell_impl._linear: concat(args, 1) . kernel[total, out] + bias, variables named "kernel" / "bias"."""
    if not _is_seq(args):
        args = [args]
    x = torch.cat([_c(a) for a in args], dim=1)
    kernel = get_variable("kernel", [x.shape[1], output_size], dtype=float32, initializer=kernel_initializer)
    y = x @ kernel.t
    if bias:
        y = y + get_variable("bias", [output_size], dtype=float32, initializer=bias_initializer or zeros_initializer()).t
    return Tensor(y)


class GRUCell(RNNCell):
    """rnn_cell_impl.GRUCell (r1.4): gates = sigmoid(_linear([x, h]) + 1-initialised bias), r,u = split(gates);
    c = tanh(_linear([x, r*h])); h' = u*h + (1-u)*c.  The reset gate multiplies h BEFORE the candidate matmul."""

    def __init__(self, num_units, activation=None, reuse=None, kernel_initializer=None, bias_initializer=None, name=None):
        super().__init__(name=name)
        self._num_units = num_units
        self._activation = activation or tanh
        self._kernel_initializer, self._bias_initializer = kernel_initializer, bias_initializer

    @property
    def state_size(self):
        return self._num_units

    @property
    def output_size(self):
        return self._num_units

    def call(self, inputs, state):
        with variable_scope("gates"):
            value = sigmoid(_linear([inputs, state], 2 * self._num_units, True,
                                    self._bias_initializer or constant_initializer(1.0), self._kernel_initializer))
            r, u = split(value, 2, axis=1)
        with variable_scope("candidate"):
            c = self._activation(_linear([inputs, r * state], self._num_units, True, self._bias_initializer, self._kernel_initializer))
        new_h = u * state + (1 - u) * c
        return new_h, new_h


class MultiRNNCell(RNNCell):
    def __init__(self, cells, state_is_tuple=True):
        super().__init__()
        assert state_is_tuple
        self._cells = list(cells)

    @property
    def state_size(self):
        return tuple(c.state_size for c in self._cells)

    @property
    def output_size(self):
        return self._cells[-1].output_size

    def zero_state(self, batch_size, dtype):
        return tuple(c.zero_state(batch_size, dtype) for c in self._cells)

    def call(self, inputs, state):
        cur, new_states = inputs, []
        for i, cell in enumerate(self._cells):
            with variable_scope("cell_%d" % i):
                cur, ns = cell(cur, state[i])
                new_states.append(ns)
        return cur, tuple(new_states)


class OutputProjectionWrapper(RNNCell):
    """core_rnn_cell.OutputProjectionWrapper: output of the wrapped cell through _linear(output, output_size, True)."""

    def __init__(self, cell, output_size, activation=None, reuse=None):
        super().__init__()
        self._cell, self._output_size, self._activation = cell, output_size, activation

    @property
    def state_size(self):
        return self._cell.state_size

    @property
    def output_size(self):
        return self._output_size

    def zero_state(self, batch_size, dtype):
        return self._cell.zero_state(batch_size, dtype)

    def call(self, inputs, state):
        output, res_state = self._cell(inputs, state)
        projected = _linear(output, self._output_size, True)
        if self._activation:
            projected = self._activation(projected)
        return projected, res_state


class ResidualWrapper(RNNCell):
    """rnn_cell_impl.ResidualWrapper overrides __call__ (no scope of its own): outputs = inputs + cell(inputs, state)."""

    def __init__(self, cell, residual_fn=None):
        super().__init__()
        self._cell = cell

    @property
    def state_size(self):
        return self._cell.state_size

    @property
    def output_size(self):
        return self._cell.output_size

    def zero_state(self, batch_size, dtype):
        return self._cell.zero_state(batch_size, dtype)

    def __call__(self, inputs, state, scope=None):
        outputs, new_state = self._cell(inputs, state, scope=scope)
        return nest_map_structure(lambda i, o: i + o, inputs, outputs), new_state


# ------------------------------------------------------------------------------------------------------------------
# dynamic_rnn / bidirectional_dynamic_rnn (python/ops/rnn.py)
# ------------------------------------------------------------------------------------------------------------------
def dynamic_rnn(cell, inputs, sequence_length=None, initial_state=None, dtype=None, scope=None, **_kw):
    x = _c(inputs)
    n, T = x.shape[:2]
    with variable_scope(scope or "rnn"):
        state = initial_state if initial_state is not None else cell.zero_state(n, dtype or float32)
        L = _c(sequence_length).long() if sequence_length is not None else None
        outs = []
        for t in range(T):
            out, new_state = cell(Tensor(x[:, t]), state)
            if L is not None:                                          # rnn._rnn_step: zero output, state copied through
                live = (t < L).view(n, 1)
                out = Tensor(torch.where(live, _c(out), torch.zeros_like(_c(out))))
                new_state = nest_map_structure(lambda ns, os_: Tensor(torch.where(live, _c(ns), _c(os_))), new_state, state)
            outs.append(_c(out))
            state = new_state
        return Tensor(torch.stack(outs, 1)), state


def bidirectional_dynamic_rnn(cell_fw, cell_bw, inputs, sequence_length=None, initial_state_fw=None, initial_state_bw=None,
                              dtype=None, scope=None, **_kw):
    with variable_scope(scope or "bidirectional_rnn"):
        with variable_scope("fw") as fw_scope:
            out_fw, st_fw = dynamic_rnn(cell_fw, inputs, sequence_length, initial_state_fw, dtype, scope=fw_scope)
        rev = (lambda z: reverse_sequence(z, sequence_length)) if sequence_length is not None else (lambda z: reverse(z, [1]))
        with variable_scope("bw") as bw_scope:
            tmp, st_bw = dynamic_rnn(cell_bw, rev(inputs), sequence_length, initial_state_bw, dtype, scope=bw_scope)
        return (out_fw, rev(tmp)), (st_fw, st_bw)


# ------------------------------------------------------------------------------------------------------------------
# seq2seq (contrib/seq2seq/python/ops/{helper,basic_decoder,decoder,attention_wrapper}.py)
# ------------------------------------------------------------------------------------------------------------------
class Helper:
    @property
    def batch_size(self):
        raise NotImplementedError

    def initialize(self, name=None):
        raise NotImplementedError

    def sample(self, time, outputs, state, name=None):
        raise NotImplementedError

    def next_inputs(self, time, outputs, state, sample_ids, name=None):
        raise NotImplementedError


BasicDecoderOutput = collections.namedtuple("BasicDecoderOutput", ("rnn_output", "sample_id"))


class BasicDecoder:
    def __init__(self, cell, helper, initial_state, output_layer=None):
        self._cell, self._helper, self._initial_state, self._output_layer = cell, helper, initial_state, output_layer

    @property
    def batch_size(self):
        return self._helper.batch_size

    def initialize(self, name=None):
        return self._helper.initialize() + (self._initial_state,)

    def step(self, time, inputs, state, name=None):
        cell_outputs, cell_state = self._cell(inputs, state)
        if self._output_layer is not None:
            cell_outputs = self._output_layer(cell_outputs)
        sample_ids = self._helper.sample(time=time, outputs=cell_outputs, state=cell_state)
        finished, next_inputs, next_state = self._helper.next_inputs(time=time, outputs=cell_outputs, state=cell_state, sample_ids=sample_ids)
        return BasicDecoderOutput(cell_outputs, sample_ids), next_state, next_inputs, finished


def dynamic_decode(decoder, output_time_major=False, impute_finished=False, maximum_iterations=None, scope=None, **_kw):
    """decoder.py dynamic_decode: loop while not all(finished); finished |= decoder_finished | (time+1 >= maximum_iterations);
    outputs of every executed step are kept (impute_finished=False)."""
    with variable_scope(scope, default_name="decoder"):
        finished, inputs, state = decoder.initialize()
        finished = _c(finished)
        if maximum_iterations is not None:
            maximum_iterations = _scalar_int(maximum_iterations)
            finished = finished | torch.tensor(0 >= maximum_iterations)
        n = finished.shape[0]
        seq_len = torch.zeros(n, dtype=torch.int32)
        time = 0
        steps: List[BasicDecoderOutput] = []
        while not bool(finished.all()):
            outputs, state, inputs, dec_finished = decoder.step(Tensor(torch.tensor(time, dtype=torch.int32)), inputs, state)
            nxt = _c(dec_finished) | finished
            if maximum_iterations is not None:
                nxt = nxt | torch.tensor(time + 1 >= maximum_iterations)
            seq_len = torch.where(~finished & nxt, torch.full_like(seq_len, time + 1), seq_len)
            steps.append(outputs)
            finished = nxt
            time += 1
        stacked = BasicDecoderOutput(*[Tensor(torch.stack([_c(getattr(s, f)) for s in steps], 0 if output_time_major else 1))
                                       for f in BasicDecoderOutput._fields])
        return stacked, state, Tensor(seq_len)


class AttentionWrapperState(collections.namedtuple(
        "AttentionWrapperState", ("cell_state", "attention", "time", "alignments", "alignment_history"))):
    def clone(self, **kwargs):
        return super()._replace(**kwargs)


class AttentionMechanism:
    pass


class _BaseAttentionMechanism(AttentionMechanism):
    """attention_wrapper._BaseAttentionMechanism: values = memory (masked only when memory_sequence_length is given — the
    reference never passes it, tacotron.py:133), keys = memory_layer(values) computed once at construction."""

    def __init__(self, query_layer, memory, probability_fn, memory_sequence_length=None, memory_layer=None,
                 check_inner_dims_defined=True, score_mask_value=float("-inf"), name=None):
        self._query_layer, self._memory_layer, self._probability_fn = query_layer, memory_layer, probability_fn
        self._base_name = name
        self._call_scope: Optional[VariableScope] = None
        with name_scope(name, "BaseAttentionMechanismInit"):
            vals = _c(memory)
            if memory_sequence_length is not None:
                L = _c(memory_sequence_length).long()
                mask = (torch.arange(vals.shape[1]).unsqueeze(0) < L.unsqueeze(1)).to(vals.dtype)
                vals = vals * mask.unsqueeze(-1)
            self._values = Tensor(vals)
            self._keys = self._memory_layer(self._values) if self._memory_layer else self._values
            self._batch_size = vals.shape[0]
            self._alignments_size = vals.shape[1]

    memory_layer = property(lambda self: self._memory_layer)
    query_layer = property(lambda self: self._query_layer)
    values = property(lambda self: self._values)
    keys = property(lambda self: self._keys)
    batch_size = property(lambda self: self._batch_size)
    alignments_size = property(lambda self: self._alignments_size)

    def initial_alignments(self, batch_size, dtype):
        return _zero_state_tensors(self._alignments_size, batch_size, dtype)

    @contextlib.contextmanager
    def _scope(self, default_name):
        """`with variable_scope(None, default_name, [query])` of __call__, captured once (TF traces the loop body once)."""
        with (variable_scope(self._call_scope) if self._call_scope is not None else variable_scope(None, default_name=default_name)) as s:
            if self._call_scope is None:
                self._call_scope = s
            yield s


def _bahdanau_score(processed_query, keys, normalize):
    k = _c(keys)
    num_units = k.shape[2]
    q = _c(processed_query).unsqueeze(1)
    v = get_variable("attention_v", [num_units], dtype=float32)
    if normalize:
        g = get_variable("attention_g", dtype=float32, initializer=math.sqrt(1.0 / num_units))
        b = get_variable("attention_b", [num_units], dtype=float32, initializer=zeros_initializer())
        normed_v = g.t * v.t * torch.rsqrt((v.t * v.t).sum())
        return Tensor((normed_v * torch.tanh(k + q + b.t)).sum(2))
    return Tensor((v.t * torch.tanh(k + q)).sum(2))


class BahdanauAttention(_BaseAttentionMechanism):
    def __init__(self, num_units, memory, memory_sequence_length=None, normalize=False, probability_fn=None,
                 score_mask_value=float("-inf"), name="BahdanauAttention"):
        if probability_fn is None:
            probability_fn = softmax
        super().__init__(query_layer=Dense(num_units, name="query_layer", use_bias=False),
                         memory_layer=Dense(num_units, name="memory_layer", use_bias=False), memory=memory,
                         probability_fn=lambda score, _prev: probability_fn(score),
                         memory_sequence_length=memory_sequence_length, score_mask_value=score_mask_value, name=name)
        self._num_units, self._normalize, self._name = num_units, normalize, name

    def __call__(self, query, previous_alignments):
        with self._scope("bahdanau_attention"):
            processed_query = self.query_layer(query) if self.query_layer else query
            score = _bahdanau_score(processed_query, self._keys, self._normalize)
        return self._probability_fn(score, previous_alignments)


def safe_cumprod(x, axis=0, exclusive=False):
    """attention_wrapper.safe_cumprod: exp(cumsum(log(clip(x, tiny, 1))))."""
    t = _c(x)
    tiny = float(np.finfo(np.float32).tiny)      # x.dtype.as_numpy_dtype of the float32 graph
    return exp(cumsum(log(clip_by_value(t, tiny, 1)), axis=axis, exclusive=exclusive))


def monotonic_attention(p_choose_i, previous_attention, mode):
    p, prev = _c(p_choose_i), _c(previous_attention)
    if mode == "recursive":
        n, T = p.shape
        shifted = torch.cat([torch.zeros(n, 1, dtype=p.dtype), 1 - p[:, :-1]], 1)
        q = torch.zeros(n, dtype=p.dtype)
        outs = []
        for j in range(T):
            q = shifted[:, j] * q + prev[:, j]
            outs.append(q)
        return Tensor(p * torch.stack(outs, 1))
    if mode == "parallel":
        cumprod_1mp = _c(safe_cumprod(1 - p, axis=1, exclusive=True))
        return Tensor(p * cumprod_1mp * torch.cumsum(prev / torch.clamp(cumprod_1mp, 1e-10, 1.0), dim=1))
    if mode == "hard":
        raise NotImplementedError
    raise ValueError("mode must be 'recursive', 'parallel', or 'hard'.")


def _monotonic_probability_fn(score, previous_alignments, sigmoid_noise, mode, seed=None):
    if sigmoid_noise > 0:
        raise NotImplementedError("the reference uses the default sigmoid_noise=0")
    return monotonic_attention(sigmoid(score), previous_alignments, mode)


class _BaseMonotonicAttentionMechanism(_BaseAttentionMechanism):
    def initial_alignments(self, batch_size, dtype):
        n = _scalar_int(batch_size)
        return one_hot(torch.zeros(n, dtype=torch.long), self._alignments_size, dtype=dtype)


class BahdanauMonotonicAttention(_BaseMonotonicAttentionMechanism):
    def __init__(self, num_units, memory, memory_sequence_length=None, normalize=False, score_mask_value=float("-inf"),
                 sigmoid_noise=0.0, sigmoid_noise_seed=None, score_bias_init=0.0, mode="parallel", dtype=None,
                 name="BahdanauMonotonicAttention"):
        fn = lambda score, prev: _monotonic_probability_fn(score, prev, sigmoid_noise, mode, sigmoid_noise_seed)  # noqa: E731
        super().__init__(query_layer=Dense(num_units, name="query_layer", use_bias=False),
                         memory_layer=Dense(num_units, name="memory_layer", use_bias=False), memory=memory,
                         probability_fn=fn, memory_sequence_length=memory_sequence_length,
                         score_mask_value=score_mask_value, name=name)
        self._num_units, self._normalize, self._name, self._score_bias_init = num_units, normalize, name, score_bias_init

    def __call__(self, query, previous_alignments):
        with self._scope("bahdanau_monotonic_attention"):
            processed_query = self.query_layer(query) if self.query_layer else query
            score = _bahdanau_score(processed_query, self._keys, self._normalize)
            score_bias = get_variable("attention_score_bias", dtype=float32, initializer=self._score_bias_init)
            score = score + score_bias
        return self._probability_fn(score, previous_alignments)


def tile_batch(t, multiplier, name=None):
    return nest_map_structure(lambda x: Tensor(_c(x).repeat_interleave(multiplier, dim=0)), t)


# ------------------------------------------------------------------------------------------------------------------
# tf.train
# ------------------------------------------------------------------------------------------------------------------
def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
    p = _c(global_step).to(float32.torch) / decay_steps
    if staircase:
        p = torch.floor(p)
    return Tensor(learning_rate * torch.pow(torch.tensor(decay_rate, dtype=float32.torch), p))


class AdamOptimizer:
    """python/training/adam.py: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; var -= lr_t*m/(sqrt(v)+eps), eps=1e-8.
    As in TF the optimizer's state lives in the graph's variable store, not in the Python object: slots ``<var>/Adam`` and
    ``<var>/Adam_1`` and the accumulators ``beta1_power`` / ``beta2_power`` (created in the scope apply_gradients runs in).
    Re-building the model against the same store therefore continues the optimisation, like a second ``sess.run``."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **_kw):
        self._lr, self._b1, self._b2, self._eps = learning_rate, beta1, beta2, epsilon

    @staticmethod
    def _state_var(full, value):
        if full not in _S.vars:
            _S.vars[full] = Variable(value, name=full, trainable=False)
        return _S.vars[full]

    def compute_gradients(self, loss, var_list=None):
        vs_ = var_list or trainable_variables()
        gs = torch.autograd.grad(_c(loss), [v.t for v in vs_], allow_unused=True, retain_graph=True)
        return [(None if g is None else Tensor(g), v) for g, v in zip(gs, vs_)]

    def get_slot(self, var, name):
        return _S.vars[var.name + {"m": "/Adam", "v": "/Adam_1"}[name]]

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        lr = float(_c(self._lr))
        b1p = self._state_var("/".join(_S.scope + ["beta1_power"]), torch.tensor(self._b1, dtype=torch.float64))
        b2p = self._state_var("/".join(_S.scope + ["beta2_power"]), torch.tensor(self._b2, dtype=torch.float64))
        lr_t = lr * math.sqrt(1 - float(b2p.t)) / (1 - float(b1p.t))
        with torch.no_grad():
            for g, v in grads_and_vars:
                if g is None:
                    continue
                gt = _c(g)
                m = self._state_var(v.name + "/Adam", torch.zeros_like(v.t)).t
                s = self._state_var(v.name + "/Adam_1", torch.zeros_like(v.t)).t
                m.mul_(self._b1).add_(gt * (1 - self._b1))
                s.mul_(self._b2).add_(gt * gt * (1 - self._b2))
                v.t.sub_(lr_t * m / (s.sqrt() + self._eps))
            b1p.t.mul_(self._b1)
            b2p.t.mul_(self._b2)
        if global_step is not None:
            global_step.assign(_c(global_step) + 1)
        return None


# ------------------------------------------------------------------------------------------------------------------
# tf.contrib.training.HParams (only what hparams.py and callers use)
# ------------------------------------------------------------------------------------------------------------------
class HParams:
    def __init__(self, **kwargs):
        self._names = list(kwargs)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def values(self):
        return {k: getattr(self, k) for k in self._names}

    def to_json(self, **kw):
        import json
        return json.dumps(self.values(), **kw)

    def set_hparam(self, name, value):
        if name not in self._names:
            self._names.append(name)
        setattr(self, name, value)

    def add_hparam(self, name, value):
        self.set_hparam(name, value)

    def parse(self, values):
        import ast
        for item in filter(None, (s.strip() for s in values.split(","))):
            k, v = item.split("=", 1)
            cur = getattr(self, k)
            self.set_hparam(k, type(cur)(ast.literal_eval(v)) if not isinstance(cur, (str, list)) else (v if isinstance(cur, str) else ast.literal_eval(v)))
        return self
