"""numpy stand-in for the TensorFlow 1.x graph API -- TEST INFRASTRUCTURE, not product code.

Purpose: let the reference's OWN source files (/root/reference/src/{models,omega}.py, src/tf_smpl/*.py,
src/evaluation/tester.py) be imported and executed unmodified in a container where TensorFlow 1.8 cannot be
installed, so that golden vectors for the hot path come from the reference's code (its op order, index handling,
reshapes, variable scopes and fetch wiring) instead of from a restatement.  See oracle/ref_exec/README.md.

What this is NOT: it is not TensorFlow.  Every `tf.*` symbol below implements the documented semantics of the TF 1.8 op
of that name on numpy float32 arrays (matmul, reshape, concat, tile, pad, scatter_nd, ...).  The layers that live in
tf.contrib (slim resnet_v2, group_norm, conv2d, fully_connected, batch_norm) are restated from their published TF 1.8
definitions in tensorflow/contrib/* of this stand-in and are marked [TF-ext] there.

Model: a lazy graph.  A Tensor is (fn, inputs); building the graph evaluates every op once on *probe* values
(placeholders = zeros) so that static shapes (`x.shape[0].value`, `x.shape.as_list()`) exist at build time like in TF1;
`Session.run(fetches, feed_dict)` re-evaluates the needed sub-graph with the fed values.  Probes of large tensors are
replaced by zero-stride views to keep memory flat.
"""
import contextlib

import numpy as np

__version__ = '1.8.0-numpy-standin'

float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64
bool = np.bool_          # noqa: A001  (tf.bool)
AUTO_REUSE = 'AUTO_REUSE'

_PROBE_KEEP = 4096       # probes with more elements than this are stored as zero-stride views


# ------------------------------------------------------------------------------------------------------------
# shapes
# ------------------------------------------------------------------------------------------------------------
class Dimension(object):
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

    def __mul__(self, other):
        return Dimension(self.value * int(other))

    __rmul__ = __mul__

    def __add__(self, other):
        return Dimension(self.value + int(other))

    __radd__ = __add__

    def __repr__(self):
        return 'Dimension(%r)' % self.value


class TensorShape(object):
    def __init__(self, dims):
        self._dims = [d if isinstance(d, Dimension) else Dimension(d) for d in dims]

    @property
    def dims(self):
        return self._dims

    @property
    def ndims(self):
        return len(self._dims)

    def as_list(self):
        return [d.value for d in self._dims]

    def __len__(self):
        return len(self._dims)

    def __iter__(self):
        return iter(self._dims)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return TensorShape(self._dims[i])
        return self._dims[i]

    def __eq__(self, other):
        try:
            other = list(other.as_list()) if isinstance(other, TensorShape) else [
                (d.value if isinstance(d, Dimension) else d) for d in other]
        except TypeError:
            return False
        return self.as_list() == other

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return 'TensorShape(%r)' % (self.as_list(),)


def _ints(shape):
    """python ints from a shape given as list / tuple / TensorShape of ints / Dimensions."""
    if isinstance(shape, (int, np.integer, Dimension)):
        return [int(shape)]
    return [int(s) for s in shape]


# ------------------------------------------------------------------------------------------------------------
# graph
# ------------------------------------------------------------------------------------------------------------
class _Graph(object):
    def __init__(self):
        self.counter = 0
        self.variables = []           # creation order == tf.GraphKeys.GLOBAL_VARIABLES
        self.var_by_name = {}
        self.scope_stack = []         # variable_scope names
        self.reuse_stack = []


_G = _Graph()


def reset_default_graph():
    global _G
    _G = _Graph()


def _shrink(a):
    a = np.asarray(a)
    if a.size > _PROBE_KEEP:
        return np.broadcast_to(np.zeros((), a.dtype), a.shape)
    return a


class Tensor(object):
    __array_priority__ = 1000        # numpy scalars / arrays defer to Tensor.__radd__ etc.
    __array_ufunc__ = None

    def __init__(self, fn, inputs, name=None):
        self.fn = fn
        self.inputs = list(inputs)
        _G.counter += 1
        self.id = _G.counter
        self.name = '%s_%d:0' % (name or getattr(fn, '__name__', 'op'), self.id)
        with np.errstate(all='ignore'):
            self.probe = _shrink(fn(*[i.probe for i in self.inputs]))

    # --- static info -------------------------------------------------------------------------------------
    @property
    def shape(self):
        return TensorShape(self.probe.shape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return self.probe.dtype.type

    def __repr__(self):
        return '<Tensor %s shape=%s dtype=%s>' % (self.name, self.probe.shape, self.probe.dtype)

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def __bool__(self):
        raise TypeError('a graph Tensor has no truth value (same as TF1)')

    def __iter__(self):
        for i in _builtin_range(self.probe.shape[0]):
            yield self[i]

    def __len__(self):
        raise TypeError('len() of a graph Tensor (same as TF1)')

    # --- operators ---------------------------------------------------------------------------------------
    def __getitem__(self, key):
        return _strided_slice(self, key)

    def __add__(self, o):
        return add(self, o)

    def __radd__(self, o):
        return add(o, self)

    def __sub__(self, o):
        return subtract(self, o)

    def __rsub__(self, o):
        return subtract(o, self)

    def __mul__(self, o):
        return multiply(self, o)

    def __rmul__(self, o):
        return multiply(o, self)

    def __truediv__(self, o):
        return div(self, o)

    def __rtruediv__(self, o):
        return div(o, self)

    __div__ = __truediv__

    def __neg__(self):
        return _op(lambda a: -a, [self], 'neg')

    def __lt__(self, o):
        return _op(lambda a, b: a < b, [self, _conv_like(o, self)], 'less')

    def __gt__(self, o):
        return _op(lambda a, b: a > b, [self, _conv_like(o, self)], 'greater')

    def __le__(self, o):
        return _op(lambda a, b: a <= b, [self, _conv_like(o, self)], 'less_equal')

    def __ge__(self, o):
        return _op(lambda a, b: a >= b, [self, _conv_like(o, self)], 'greater_equal')


def _op(fn, inputs, name=None):
    return Tensor(fn, inputs, name)


def _const_tensor(value, name='Const'):
    value = np.asarray(value)
    return Tensor(lambda: value, [], name)


def convert_to_tensor(x, dtype=None, name=None):
    if isinstance(x, Tensor):
        return x
    if isinstance(x, Dimension):
        x = x.value
    a = np.asarray(x)
    if dtype is not None:
        a = a.astype(dtype)
    elif a.dtype == np.float64:
        a = a.astype(np.float32)            # python floats become float32 constants, like tf.convert_to_tensor
    elif a.dtype == np.int64:
        a = a.astype(np.int32)              # python ints become int32 constants
    return _const_tensor(a, name or 'Const')


def _conv_like(x, ref):
    """Second operand of a binary op: python scalars / numpy arrays take the dtype of the Tensor operand (TF behaviour)."""
    if isinstance(x, Tensor):
        return x
    if isinstance(x, Dimension):
        x = x.value
    return _const_tensor(np.asarray(x).astype(ref.probe.dtype))


def _binary(fn, name):
    def op(x, y, name_=None, **kw):
        if isinstance(x, Tensor):
            y = _conv_like(y, x)
        elif isinstance(y, Tensor):
            x = _conv_like(x, y)
        else:
            x = convert_to_tensor(x)
            y = _conv_like(y, x)
        return _op(fn, [x, y], name)
    op.__name__ = name
    return op


add = _binary(lambda a, b: a + b, 'add')
subtract = _binary(lambda a, b: a - b, 'sub')
multiply = _binary(lambda a, b: a * b, 'mul')


def _div_fn(a, b):
    if a.dtype.kind in 'iu' and b.dtype.kind in 'iu':
        return a // b
    return a / b


div = _binary(_div_fn, 'div')
divide = div
maximum = _binary(np.maximum, 'maximum')
minimum = _binary(np.minimum, 'minimum')


def _unary(fn, name):
    def op(x, name_=None, **kw):
        return _op(fn, [convert_to_tensor(x)], name)
    op.__name__ = name
    return op


cos = _unary(np.cos, 'cos')
sin = _unary(np.sin, 'sin')
acos = _unary(np.arccos, 'acos')
sqrt = _unary(np.sqrt, 'sqrt')
rsqrt = _unary(lambda a: (1.0 / np.sqrt(a)).astype(a.dtype), 'rsqrt')
abs = _unary(np.abs, 'abs')          # noqa: A001
square = _unary(np.square, 'square')
identity = _unary(lambda a: a, 'identity')
stop_gradient = identity


def cast(x, dtype, name=None):
    return _op(lambda a: a.astype(dtype), [convert_to_tensor(x)], 'cast')


def to_float(x, name=None):
    return cast(x, np.float32)


def clip_by_value(t, lo, hi, name=None):
    t = convert_to_tensor(t)
    return _op(lambda a: np.clip(a, np.asarray(lo, a.dtype), np.asarray(hi, a.dtype)), [t], 'clip_by_value')


def where(cond, x=None, y=None, name=None):
    x = convert_to_tensor(x)
    return _op(lambda c, a, b: np.where(c, a, b), [convert_to_tensor(cond), x, _conv_like(y, x)], 'where')


# --- slicing ---------------------------------------------------------------------------------------------------
def _strided_slice(t, key):
    if not isinstance(key, tuple):
        key = (key,)

    def norm(k):
        if isinstance(k, Dimension):
            return k.value
        if isinstance(k, slice):
            return slice(norm(k.start), norm(k.stop), norm(k.step))
        if isinstance(k, np.integer):
            return int(k)
        return k
    key = tuple(norm(k) for k in key)
    for k in key:
        if isinstance(k, Tensor):
            raise NotImplementedError('tensor-valued slice index')
    return _op(lambda a: a[key], [t], 'strided_slice')


# --- creation --------------------------------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name='Const'):
    a = np.asarray(value)
    if dtype is None:
        dtype = np.float32 if a.dtype.kind == 'f' or a.size == 0 else (np.int32 if a.dtype.kind in 'iu' else a.dtype)
    a = a.astype(dtype)
    if shape is not None:
        shape = _ints(shape)
        if a.size == int(np.prod(shape)):
            a = a.reshape(shape)
        elif a.size == 1:
            a = np.full(shape, a.reshape(()), dtype)
        else:
            raise ValueError('constant: %d values for shape %r' % (a.size, shape))
    return _const_tensor(a, name)


def ones(shape, dtype=np.float32, name=None):
    return _const_tensor(np.ones(_ints(shape), dtype), 'ones')


def zeros(shape, dtype=np.float32, name=None):
    return _const_tensor(np.zeros(_ints(shape), dtype), 'zeros')


def eye(n, dtype=np.float32, name=None):
    return _const_tensor(np.eye(int(n), dtype=dtype), 'eye')


def range(start, limit=None, delta=1, dtype=None, name=None):     # noqa: A001
    if limit is None:
        start, limit = 0, start
    a = np.arange(int(start), int(limit), int(delta), dtype=dtype or np.int32)
    return _const_tensor(a, 'range')


def placeholder(dtype, shape=None, name=None):
    shape = _ints(shape)
    t = Tensor(lambda: np.broadcast_to(np.zeros((), dtype), shape), [], name or 'Placeholder')
    t.is_placeholder = True
    return t


class Variable(Tensor):
    """tf.Variable: a named, assignable leaf.  Evaluating an uninitialised variable raises (FailedPrecondition in TF)."""

    def __init__(self, initial_value=None, trainable=True, name=None, dtype=None, _shape=None, _full_name=None):
        if initial_value is not None:
            val = np.asarray(initial_value)
            if isinstance(initial_value, Tensor):
                raise NotImplementedError
            val = np.asarray(val, dtype=dtype or (np.float32 if val.dtype.kind == 'f' else val.dtype))
            self.initialized = True
        else:
            val = np.zeros(_ints(_shape), dtype or np.float32)
            self.initialized = False
        self.value = val
        self.trainable = trainable
        base = _full_name if _full_name is not None else '/'.join(_G.scope_stack + [name or 'Variable'])
        full, k = base, 0
        while full in _G.var_by_name:        # tf.Variable uniquifies; get_variable (below) never reaches this
            k += 1
            full = '%s_%d' % (base, k)
        Tensor.__init__(self, self._read, [], base)
        self.name = full + ':0'
        self.op_name = full
        _G.variables.append(self)
        _G.var_by_name[full] = self

    def _read(self):
        return self.value

    def load(self, value):
        value = np.asarray(value)
        if tuple(value.shape) != tuple(self.value.shape):
            raise ValueError('shape mismatch restoring %s: checkpoint %r vs variable %r' % (self.name, value.shape, self.value.shape))
        self.value = value.astype(self.value.dtype)
        self.initialized = True


def global_variables():
    return list(_G.variables)


def trainable_variables():
    return [v for v in _G.variables if v.trainable]


# --- scopes ----------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def name_scope(name=None, default_name=None, values=None):
    """Name scopes only affect op names (never variable names created through get_variable); nothing here depends on them."""
    yield name if isinstance(name, str) else default_name


class VariableScope(object):
    def __init__(self, name, reuse):
        self.name = name
        self.reuse = reuse

    def __str__(self):
        return self.name


@contextlib.contextmanager
def variable_scope(name_or_scope=None, default_name=None, values=None, reuse=None, **kw):
    if isinstance(name_or_scope, VariableScope):
        saved = (_G.scope_stack, _G.reuse_stack)
        _G.scope_stack = name_or_scope.name.split('/') if name_or_scope.name else []
        _G.reuse_stack = [reuse if reuse is not None else name_or_scope.reuse]
        try:
            yield name_or_scope
        finally:
            _G.scope_stack, _G.reuse_stack = saved
        return
    name = name_or_scope if name_or_scope is not None else default_name
    if name is None:
        raise ValueError('variable_scope needs a name or default_name')
    if name_or_scope is None:
        # default_name is uniquified within the enclosing scope (bottleneck_v2, bottleneck_v2_1, ...): each unit has its own
        # enclosing `unit_k` scope in resnet_v2, so the first name is always free there
        prefix = '/'.join(_G.scope_stack + [name])
        taken = [n for n in _G.var_by_name if n.startswith(prefix + '/')]
        if taken:
            raise NotImplementedError('default_name scope %r would need uniquifying' % prefix)
    inherited = _G.reuse_stack[-1] if _G.reuse_stack else None
    eff = reuse if reuse is not None else inherited          # reuse=None / False inherit; True and AUTO_REUSE propagate down
    if reuse is False:
        eff = inherited
    _G.scope_stack.append(name)
    _G.reuse_stack.append(eff)
    try:
        yield VariableScope('/'.join(_G.scope_stack), eff)
    finally:
        _G.scope_stack.pop()
        _G.reuse_stack.pop()


def get_variable_scope():
    return VariableScope('/'.join(_G.scope_stack), _G.reuse_stack[-1] if _G.reuse_stack else None)


def get_variable(name, shape=None, dtype=np.float32, initializer=None, trainable=True, **kw):
    full = '/'.join(_G.scope_stack + [name])
    reuse = _G.reuse_stack[-1] if _G.reuse_stack else None
    if full in _G.var_by_name:
        if reuse in (True, AUTO_REUSE):
            v = _G.var_by_name[full]
            if shape is not None and _ints(shape) != list(v.value.shape):
                raise ValueError('get_variable(%s): shape %r vs existing %r' % (full, shape, v.value.shape))
            return v
        raise ValueError('Variable %s already exists, disallowed. Did you mean to set reuse=True or reuse=tf.AUTO_REUSE?' % full)
    if reuse is True:
        raise ValueError('Variable %s does not exist, or was not created with tf.get_variable()' % full)
    return Variable(None, trainable=trainable, dtype=dtype, _shape=shape, _full_name=full)


# --- shape ops -------------------------------------------------------------------------------------------------
def _shape_arg(shape):
    """A target shape: python list / tuple (possibly holding Dimensions or scalar int Tensors) or an int Tensor."""
    if isinstance(shape, Tensor):
        return shape, None
    parts = list(shape) if not isinstance(shape, (int, np.integer, Dimension)) else [shape]
    if builtins_any(isinstance(p, Tensor) for p in parts):
        return stack([convert_to_tensor(p, np.int32) if not isinstance(p, Tensor) else p for p in parts]), None
    return None, [int(p) for p in parts]


import builtins as _b          # noqa: E402
builtins_any = _b.any
_builtin_range = _b.range
_builtin_abs = _b.abs


def reshape(tensor, shape, name=None):
    tensor = convert_to_tensor(tensor)
    st, sl = _shape_arg(shape)
    if st is not None:
        return _op(lambda a, s: a.reshape([int(v) for v in s]), [tensor, st], 'reshape')
    return _op(lambda a: a.reshape(sl), [tensor], 'reshape')


def shape(x, name=None, out_type=np.int32):
    return _op(lambda a: np.asarray(a.shape, out_type), [convert_to_tensor(x)], 'shape')


def expand_dims(x, axis=None, name=None, dim=None):
    axis = dim if axis is None else axis
    return _op(lambda a: np.expand_dims(a, axis), [convert_to_tensor(x)], 'expand_dims')


def squeeze(x, axis=None, name=None, squeeze_dims=None):
    axis = squeeze_dims if axis is None else axis
    ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
    return _op(lambda a: np.squeeze(a, ax), [convert_to_tensor(x)], 'squeeze')


def _all_tensors(values):
    ts = [v for v in values if isinstance(v, Tensor)]
    ref = ts[0] if ts else None
    return [v if isinstance(v, Tensor) else (_conv_like(v, ref) if ref is not None else convert_to_tensor(v)) for v in values]


def concat(values, axis, name='concat'):
    values = _all_tensors(list(values))
    return _op(lambda *a: np.concatenate(a, axis=axis), values, 'concat')


def stack(values, axis=0, name='stack'):
    values = _all_tensors(list(values))
    return _op(lambda *a: np.stack(a, axis=axis), values, 'stack')


def tile(x, multiples, name=None):
    m = _ints(multiples)
    return _op(lambda a: np.tile(a, m), [convert_to_tensor(x)], 'tile')


def pad(tensor, paddings, mode='CONSTANT', name=None, constant_values=0):
    if mode != 'CONSTANT':
        raise NotImplementedError(mode)
    p = [(int(a), int(b)) for a, b in paddings]
    return _op(lambda a: np.pad(a, p, mode='constant', constant_values=constant_values), [convert_to_tensor(tensor)], 'pad')


def transpose(a, perm=None, name=None):
    return _op(lambda x: np.transpose(x, perm), [convert_to_tensor(a)], 'transpose')


def gather(params, indices, validate_indices=None, name=None, axis=0):
    return _op(lambda p, i: np.take(p, i, axis=axis), [convert_to_tensor(params), convert_to_tensor(indices)], 'gather')


def scatter_nd(indices, updates, shape, name=None):
    """Sums `updates` into zeros(shape) at `indices` (duplicates accumulate), indices [..., 1..rank]."""
    shp = _ints(shape)

    def fn(idx, upd):
        out = np.zeros(shp, upd.dtype)
        np.add.at(out, tuple(idx[..., k] for k in _builtin_range(idx.shape[-1])), upd)
        return out
    return _op(fn, [convert_to_tensor(indices), convert_to_tensor(updates)], 'scatter_nd')


# --- math ------------------------------------------------------------------------------------------------------
def matmul(a, b, transpose_a=False, transpose_b=False, name=None, **kw):
    a, b = convert_to_tensor(a), convert_to_tensor(b)

    def fn(x, y):
        if transpose_a:
            x = np.swapaxes(x, -1, -2)
        if transpose_b:
            y = np.swapaxes(y, -1, -2)
        return np.matmul(x, y)
    return _op(fn, [a, b], 'matmul')


def _reduce(npfn, name):
    def op(x, axis=None, keepdims=False, name_=None, keep_dims=None, reduction_indices=None, **kw):
        if keep_dims is not None:
            keepdims = keep_dims
        if axis is None:
            axis = reduction_indices
        ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
        return _op(lambda a: np.asarray(npfn(a, axis=ax, keepdims=keepdims)).astype(a.dtype), [convert_to_tensor(x)], name)
    op.__name__ = name
    return op


reduce_sum = _reduce(np.sum, 'reduce_sum')
reduce_mean = _reduce(np.mean, 'reduce_mean')
reduce_max = _reduce(np.max, 'reduce_max')


def norm(tensor, ord='euclidean', axis=None, keepdims=None, name=None, keep_dims=None):    # noqa: A002
    if ord not in ('euclidean', 2):
        raise NotImplementedError(ord)
    kd = _b.bool(keepdims if keepdims is not None else keep_dims)
    # tf.norm: sqrt(reduce_sum(x * conj(x), axis))
    return _op(lambda a: np.sqrt(np.sum(a * a, axis=axis, keepdims=kd)).astype(a.dtype), [convert_to_tensor(tensor)], 'norm')


def trace(x, name=None):
    return _op(lambda a: np.trace(a, axis1=-2, axis2=-1).astype(a.dtype), [convert_to_tensor(x)], 'trace')


def matrix_inverse(x, adjoint=False, name=None):
    return _op(lambda a: np.linalg.inv(a).astype(a.dtype), [convert_to_tensor(x)], 'matrix_inverse')


# --- session ---------------------------------------------------------------------------------------------------
class GPUOptions(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


class ConfigProto(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _flatten_fetches(f):
    if isinstance(f, dict):
        return [t for v in f.values() for t in _flatten_fetches(v)]
    if isinstance(f, (list, tuple)):
        return [t for v in f for t in _flatten_fetches(v)]
    return [f]


def _rebuild(f, values):
    if isinstance(f, dict):
        return {k: _rebuild(v, values) for k, v in f.items()}
    if isinstance(f, (list, tuple)):
        return type(f)(_rebuild(v, values) for v in f)
    return values[id(f)]


class Session(object):
    def __init__(self, target='', graph=None, config=None):
        self.config = config

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass

    def run(self, fetches, feed_dict=None):
        feed = {}
        for k, v in (feed_dict or {}).items():
            v = np.asarray(v, dtype=k.probe.dtype)
            if tuple(v.shape) != tuple(k.probe.shape):
                raise ValueError('Cannot feed value of shape %r for Tensor %s, which has shape %r' % (v.shape, k.name, k.probe.shape))
            feed[id(k)] = v
        flat = _flatten_fetches(fetches)
        # needed sub-graph; creation ids are a topological order (inputs are always created before their consumers)
        need, todo = {}, list(flat)
        while todo:
            t = todo.pop()
            if id(t) in need:
                continue
            need[id(t)] = t
            if id(t) not in feed:
                todo.extend(t.inputs)
        order = sorted(need.values(), key=lambda t: t.id)
        uses = {}
        for t in order:
            if id(t) in feed:
                continue
            for i in t.inputs:
                uses[id(i)] = uses.get(id(i), 0) + 1
        keep = set(id(t) for t in flat)
        val = {}
        for t in order:
            if id(t) in feed:
                val[id(t)] = feed[id(t)]
                continue
            if getattr(t, 'is_placeholder', False):
                raise ValueError('You must feed a value for placeholder tensor %s' % t.name)
            if isinstance(t, Variable) and not t.initialized:
                raise RuntimeError('Attempting to use uninitialized value %s' % t.name)
            val[id(t)] = np.asarray(t.fn(*[val[id(i)] for i in t.inputs]))
            for i in t.inputs:
                uses[id(i)] -= 1
                if uses[id(i)] == 0 and id(i) not in keep:
                    del val[id(i)]           # free intermediates as soon as their last consumer ran
        return _rebuild(fetches, val)


from . import train, nn, contrib      # noqa: E402,F401
