"""A NumPy stand-in for the slice of TensorFlow 1.x that the reference's dense-head path uses.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): it exists so that /root/reference/{dataset,utils,losses,
augmentation}.py can be imported and executed UNMODIFIED as a second oracle (oracle/ref_exec.py).  TensorFlow
itself is not installable in this image.  Nothing in the product imports this package.

Semantics kept from TF 1.x (each one matters for parity):
  * eager evaluation of every op, each result rounded to its dtype (float32 stays float32, no FMA);
  * binary operators convert the non-Tensor operand to the Tensor's dtype (`ops.convert_to_tensor(y,
    dtype=x.dtype)`): an ndarray is cast with astype (float64 -> int32 TRUNCATES, dataset.py:53), a Python
    scalar becomes a constant of that dtype;
  * `/` on two integer tensors is true division computed in float64 (`_truediv_python3`);
  * `tf.linspace`: step = (stop - start) / (num - 1); v[i] = start + step * i in the input dtype;
  * `tf.argmax`: int64, first maximum;  `tf.one_hot`: out-of-range index -> all-zero row;
  * `tf.image.non_max_suppression`: the CPU kernel's greedy algorithm (descending score, stable; suppress a
    candidate whose IoU with a kept box is > threshold; corners min/max-normalised; area <= 0 -> IoU 0);
  * `tf.losses.huber_loss` / `compute_weighted_loss` with SUM_BY_NONZERO_WEIGHTS;
  * `tf.nn.sigmoid_cross_entropy_with_logits`: relu(x) - x z + log1p(exp(-|x|)) built from where/exp/log1p;
  * `tf.gradients`: reverse mode with TF's registered gradient formulas (SigmoidGrad y(1-y)dy, Log dy/x,
    Pow dy*y*x^(y-1), Select, Sum/Mean broadcasts, RealDiv, ...), evaluated in the tensors' own dtype.
Deliberate deviations: reductions over float32 accumulate in float64 and round once (Eigen's tree order is
unknowable; this is closer to the exact sum than any float32 order); exp/log are NumPy's float32 ufuncs,
not Eigen's (both ~1 ulp).
"""
import contextlib
import unittest

import numpy as np

__version__ = '1.x-numpy-stub'


# ---- dtypes -------------------------------------------------------------------------------------------------
class DType:
    def __init__(self, name, np_dtype):
        self.name, self.np = name, np.dtype(np_dtype)

    @property
    def base_dtype(self):
        return self

    @property
    def as_numpy_dtype(self):
        return self.np.type

    @property
    def is_integer(self):
        return self.np.kind in 'iu'

    @property
    def is_floating(self):
        return self.np.kind == 'f'

    def __repr__(self):
        return 'tf.' + self.name

    def __eq__(self, other):
        return isinstance(other, DType) and other.np == self.np

    def __hash__(self):
        return hash(self.np)


float32 = DType('float32', np.float32)
float64 = DType('float64', np.float64)
int32 = DType('int32', np.int32)
int64 = DType('int64', np.int64)
uint8 = DType('uint8', np.uint8)
bool = DType('bool', np.bool_)   # noqa: A001  (tf.bool)
string = DType('string', np.object_)
_BY_NP = {d.np: d for d in (float32, float64, int32, int64, uint8, bool)}


def _as_np_dtype(dtype):
    if dtype is None:
        return None
    return dtype.np if isinstance(dtype, DType) else np.dtype(dtype)


# ---- the tensor -----------------------------------------------------------------------------------------------
class Tensor:
    """An eagerly evaluated value + the tape entry that produced it."""
    __array_priority__ = 1000.0   # ndarray (op) Tensor defers to Tensor.__r<op>__
    __array_ufunc__ = None

    def __init__(self, value, parents=(), vjp=None):
        self._v = value if isinstance(value, np.ndarray) else np.asarray(value)
        self._parents = tuple(parents)
        self._vjp = vjp   # grad_out (ndarray) -> tuple of ndarray-or-None, one per parent

    # -- introspection
    @property
    def dtype(self):
        return _BY_NP.get(self._v.dtype, DType(str(self._v.dtype), self._v.dtype))

    @property
    def shape(self):
        return tuple(int(d) for d in self._v.shape)

    def get_shape(self):
        return self.shape

    def numpy(self):
        return self._v

    def __repr__(self):
        return 'Tensor(%r)' % (self._v,)

    def __len__(self):
        return len(self._v)

    def __iter__(self):
        for i in _range(self._v.shape[0]):
            yield self[i]

    def __bool__(self):
        if self._v.shape == ():
            return builtins_bool(self._v)
        raise TypeError('Using a tf.Tensor as a Python bool is not allowed')

    def __index__(self):
        if self._v.shape == () and self._v.dtype.kind in 'iu':
            return int(self._v)
        raise TypeError('only integer scalar tensors can be used as an index')

    def __int__(self):
        return int(self._v)

    def __float__(self):
        return float(self._v)

    __hash__ = object.__hash__

    # -- indexing (strided_slice)
    def __getitem__(self, idx):
        idx = _static_index(idx)
        src = self

        def vjp(g):
            out = np.zeros_like(src._v)
            out[idx] = g   # basic indexing only on the gradient path
            return (out,)
        return Tensor(self._v[idx], (self,), vjp)

    # -- operators
    def __neg__(self):
        return negative(self)

    def __abs__(self):
        return abs(self)   # noqa: F821 (module-level abs below)

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
        return truediv(self, o)

    def __rtruediv__(self, o):
        return truediv(o, self)

    def __pow__(self, o):
        return pow(self, o)   # noqa: F821

    def __rpow__(self, o):
        return pow(o, self)   # noqa: F821

    def __lt__(self, o):
        return less(self, o)

    def __le__(self, o):
        return less_equal(self, o)

    def __gt__(self, o):
        return greater(self, o)

    def __ge__(self, o):
        return greater_equal(self, o)

    def __invert__(self):
        return logical_not(self)

    def __and__(self, o):
        return logical_and(self, o)

    def __or__(self, o):
        return logical_or(self, o)


import builtins as _builtins   # noqa: E402

builtins_bool = _builtins.bool
_range = _builtins.range


def _static_index(idx):
    def one(i):
        if isinstance(i, Tensor):
            return i._v if i._v.dtype == np.bool_ else int(i._v)
        if isinstance(i, slice):
            return slice(*[None if v is None else int(v) for v in (i.start, i.stop, i.step)])
        return i
    return tuple(one(i) for i in idx) if isinstance(idx, tuple) else one(idx)


def convert_to_tensor(value, dtype=None, name=None, preferred_dtype=None):
    """ops.convert_to_tensor: Python floats -> float32, Python ints -> int32, ndarrays keep their dtype; with an
    explicit dtype an ndarray is cast with astype (make_tensor_proto), i.e. float -> int truncates."""
    npd = _as_np_dtype(dtype)
    if isinstance(value, Tensor):
        if npd is not None and value._v.dtype != npd:
            raise TypeError('Tensor conversion requested dtype %s for Tensor with dtype %s' % (npd, value._v.dtype))
        return value
    if isinstance(value, (list, tuple)) and any(isinstance(v, Tensor) for v in _flatten(value)):
        return stack([convert_to_tensor(v, dtype) for v in value], 0)   # auto-packing
    arr = np.asarray(value)
    if npd is not None:
        return Tensor(arr.astype(npd))
    if isinstance(value, np.ndarray) or isinstance(value, np.generic):
        return Tensor(arr)
    if arr.dtype.kind == 'f':
        return Tensor(arr.astype(np.float32))
    if arr.dtype.kind in 'iu':
        return Tensor(arr.astype(np.int32))
    return Tensor(arr)


constant = convert_to_tensor


def _flatten(x):
    for v in x:
        if isinstance(v, (list, tuple)):
            yield from _flatten(v)
        else:
            yield v


def _pair(x, y):
    """Operands of a binary op, converted the way TF's operator overloads / op_def_library do."""
    if isinstance(x, Tensor) and isinstance(y, Tensor):
        if x._v.dtype != y._v.dtype:
            raise TypeError('binary op on tensors of different dtypes: %s vs %s' % (x._v.dtype, y._v.dtype))
        return x, y
    if isinstance(x, Tensor):
        return x, convert_to_tensor(y, dtype=x._v.dtype)
    if isinstance(y, Tensor):
        return convert_to_tensor(x, dtype=y._v.dtype), y
    x = convert_to_tensor(x)
    return x, convert_to_tensor(y, dtype=x._v.dtype)


def _unbroadcast(g, shape):
    """Sum a broadcast gradient back to `shape` (TF's BroadcastGradientArgs + Sum + Reshape)."""
    if g.shape == tuple(shape):
        return g
    extra = g.ndim - len(shape)
    if extra > 0:
        g = g.sum(axis=tuple(_range(extra)), dtype=_acc(g.dtype)).astype(g.dtype)
    axes = tuple(i for i, d in enumerate(shape) if d == 1 and g.shape[i] != 1)
    if axes:
        g = g.sum(axis=axes, keepdims=True, dtype=_acc(g.dtype)).astype(g.dtype)
    return g.reshape(shape)


def _acc(dt):
    return np.float64 if np.dtype(dt).kind == 'f' else None


def _binary(fn, grad_fn):
    def op(x, y, name=None):
        x, y = _pair(x, y)
        with np.errstate(all='ignore'):
            out = fn(x._v, y._v)

        def vjp(g):
            with np.errstate(all='ignore'):
                gx, gy = grad_fn(g, x._v, y._v, out)
            return (None if gx is None else _unbroadcast(gx, x._v.shape),
                    None if gy is None else _unbroadcast(gy, y._v.shape))
        return Tensor(out, (x, y), vjp)
    return op


add = _binary(np.add, lambda g, x, y, o: (g, g))
subtract = _binary(np.subtract, lambda g, x, y, o: (g, -g))
multiply = _binary(np.multiply, lambda g, x, y, o: (g * y, g * x))
# RealDivGrad: dx = dy / y;  dy = dy * (-x / y / y)
realdiv = _binary(np.divide, lambda g, x, y, o: (g / y, g * (-x / y / y)))
# MaximumGrad / MinimumGrad: the gradient goes to x where x >= y (resp. <=), else to y
maximum = _binary(np.maximum, lambda g, x, y, o: (np.where(x >= y, g, 0).astype(g.dtype), np.where(x >= y, 0, g).astype(g.dtype)))
minimum = _binary(np.minimum, lambda g, x, y, o: (np.where(x <= y, g, 0).astype(g.dtype), np.where(x <= y, 0, g).astype(g.dtype)))


def _pow_grad(g, x, y, o):
    # PowGrad: dx = dy * y * x^(y-1);  dy = dy * z * log(x) where x > 0 else 0
    gx = g * y * np.power(x, y - 1)
    gy = g * o * np.where(x > 0, np.log(np.where(x > 0, x, 1)), 0).astype(x.dtype)
    return gx, gy


pow = _binary(np.power, _pow_grad)   # noqa: A001


def truediv(x, y, name=None):
    x, y = _pair(x, y)
    if x._v.dtype.kind in 'iu':   # _truediv_python3: int8/16 -> float32, int32/64 -> float64
        to = np.float32 if x._v.dtype.itemsize < 4 else np.float64
        x, y = cast(x, to), cast(y, to)
    return realdiv(x, y)


div = truediv
divide = truediv


def _compare(fn):
    def op(x, y, name=None):
        x, y = _pair(x, y)
        with np.errstate(all='ignore'):
            return Tensor(fn(x._v, y._v))
    return op


less, less_equal = _compare(np.less), _compare(np.less_equal)
greater, greater_equal = _compare(np.greater), _compare(np.greater_equal)
equal, not_equal = _compare(np.equal), _compare(np.not_equal)
logical_or, logical_and = _compare(np.logical_or), _compare(np.logical_and)


def logical_not(x, name=None):
    return Tensor(np.logical_not(convert_to_tensor(x)._v))


def _unary(fn, grad_fn):
    def op(x, name=None):
        x = convert_to_tensor(x)
        with np.errstate(all='ignore'):
            out = fn(x._v)

        def vjp(g):
            with np.errstate(all='ignore'):
                return (grad_fn(g, x._v, out),)
        return Tensor(out, (x,), vjp)
    return op


negative = _unary(np.negative, lambda g, x, o: -g)
exp = _unary(np.exp, lambda g, x, o: g * o)
log = _unary(np.log, lambda g, x, o: g / x)                      # LogGrad: dy * reciprocal(x)
log1p = _unary(np.log1p, lambda g, x, o: g / (1 + x))
abs = _unary(np.abs, lambda g, x, o: g * np.sign(x))           # noqa: A001
sqrt = _unary(np.sqrt, lambda g, x, o: g * 0.5 / o)
square = _unary(np.square, lambda g, x, o: g * (2 * x))
ceil = _unary(np.ceil, lambda g, x, o: None)
floor = _unary(np.floor, lambda g, x, o: None)
round = _unary(np.rint, lambda g, x, o: None)                   # noqa: A001  (half to even, like TF)
identity = _unary(lambda v: v, lambda g, x, o: g)
stop_gradient = _unary(lambda v: v, lambda g, x, o: None)


def _sigmoid_np(v):
    one = v.dtype.type(1)
    return one / (one + np.exp(-v))


sigmoid = _unary(_sigmoid_np, lambda g, x, o: g * o * (1 - o))  # SigmoidGrad: dy * y * (1 - y)


def cast(x, dtype, name=None):
    x = convert_to_tensor(x)
    npd = _as_np_dtype(dtype)
    src = x._v.dtype
    grad_ok = src.kind == 'f' and npd.kind == 'f'
    return Tensor(x._v.astype(npd), (x,), (lambda g: (g.astype(src),)) if grad_ok else (lambda g: (None,)))


def to_float(x, name=None):
    return cast(x, np.float32)


def to_double(x, name=None):
    return cast(x, np.float64)


def to_int32(x, name=None):
    return cast(x, np.int32)   # float -> int truncates toward zero, as TF's Cast


def to_int64(x, name=None):
    return cast(x, np.int64)


# ---- shapes, structure ------------------------------------------------------------------------------------------
def shape(x, name=None, out_type=int32):
    return Tensor(np.array(convert_to_tensor(x)._v.shape, dtype=_as_np_dtype(out_type)))


def size(x, name=None):
    return Tensor(np.int32(convert_to_tensor(x)._v.size))


def _dims(d):
    if isinstance(d, Tensor):
        return tuple(int(v) for v in np.atleast_1d(d._v))
    return tuple(int(v) for v in d)


def reshape(x, shape, name=None):   # noqa: A002
    x = convert_to_tensor(x)
    src = x._v.shape
    return Tensor(x._v.reshape(_dims(shape)), (x,), lambda g: (g.reshape(src),))


def expand_dims(x, axis, name=None):
    x = convert_to_tensor(x)
    src = x._v.shape
    return Tensor(np.expand_dims(x._v, int(axis)), (x,), lambda g: (g.reshape(src),))


def squeeze(x, axis=None, name=None):
    x = convert_to_tensor(x)
    src = x._v.shape
    return Tensor(np.squeeze(x._v, axis=axis), (x,), lambda g: (g.reshape(src),))


def tile(x, multiples, name=None):
    x = convert_to_tensor(x)
    m = _dims(multiples)
    src = x._v.shape

    def vjp(g):
        split = g.reshape([v for pair in zip(m, src) for v in pair])
        return (split.sum(axis=tuple(_range(0, 2 * len(m), 2)), dtype=_acc(g.dtype)).astype(g.dtype),)
    return Tensor(np.tile(x._v, m), (x,), vjp)


def concat(values, axis, name=None):
    ts = [convert_to_tensor(v) for v in values]
    axis = int(axis)
    sizes = [t._v.shape[axis] if t._v.ndim else 1 for t in ts]

    def vjp(g):
        return tuple(np.split(g, np.cumsum(sizes)[:-1], axis=axis))
    return Tensor(np.concatenate([t._v for t in ts], axis=axis), ts, vjp)


def stack(values, axis=0, name=None):
    ts = [convert_to_tensor(v) for v in values]
    if len({t._v.dtype for t in ts}) > 1:
        raise TypeError('tf.stack: mixed dtypes')
    axis = int(axis)

    def vjp(g):
        return tuple(np.moveaxis(g, axis if axis >= 0 else g.ndim + axis, 0))
    return Tensor(np.stack([t._v for t in ts], axis=axis), ts, vjp)


def unstack(value, num=None, axis=0, name=None):
    value = convert_to_tensor(value)
    n = value._v.shape[axis]
    idx = [builtins_slice_all(value._v.ndim, axis, i) for i in _range(n)]
    return [value[i] for i in idx]


def builtins_slice_all(ndim, axis, i):
    axis = axis if axis >= 0 else ndim + axis
    return tuple(i if d == axis else slice(None) for d in _range(ndim))


def split(value, num_or_size_splits, axis=0, name=None):
    value = convert_to_tensor(value)
    axis = int(axis)
    axis = axis if axis >= 0 else value._v.ndim + axis
    if isinstance(num_or_size_splits, int):
        n = value._v.shape[axis]
        sizes = [n // num_or_size_splits] * num_or_size_splits
    else:
        sizes = list(num_or_size_splits)
    out, start = [], 0
    for s in sizes:
        out.append(value[tuple(slice(start, start + s) if d == axis else slice(None) for d in _range(value._v.ndim))])
        start += s
    return out


def reverse(x, axis, name=None):
    x = convert_to_tensor(x)
    ax = tuple(int(a) for a in _dims(axis))
    return Tensor(np.flip(x._v, ax), (x,), lambda g: (np.flip(g, ax),))


def transpose(x, perm=None, name=None):
    x = convert_to_tensor(x)
    p = None if perm is None else _dims(perm)
    inv = None if p is None else tuple(np.argsort(p))
    return Tensor(np.transpose(x._v, p), (x,), lambda g: (np.transpose(g, inv),))


def zeros_like(x, dtype=None, name=None):
    x = convert_to_tensor(x)
    return Tensor(np.zeros(x._v.shape, _as_np_dtype(dtype) or x._v.dtype))


def ones_like(x, dtype=None, name=None):
    x = convert_to_tensor(x)
    return Tensor(np.ones(x._v.shape, _as_np_dtype(dtype) or x._v.dtype))


def zeros(shape, dtype=float32, name=None):   # noqa: A002
    return Tensor(np.zeros(_dims(shape), _as_np_dtype(dtype)))


def ones(shape, dtype=float32, name=None):   # noqa: A002
    return Tensor(np.ones(_dims(shape), _as_np_dtype(dtype)))


def fill(dims, value, name=None):
    v = convert_to_tensor(value)
    return Tensor(np.full(_dims(dims), v._v, dtype=v._v.dtype))


def range(*args, dtype=None, name=None):   # noqa: A001
    vals = [int(a) if not isinstance(a, float) else a for a in (a._v if isinstance(a, Tensor) else a for a in args)]
    return Tensor(np.arange(*vals).astype(_as_np_dtype(dtype) or (np.int32 if all(isinstance(v, int) for v in vals) else np.float32)))


def linspace(start, stop, num, name=None):
    """LinSpaceOp (tensorflow/core/kernels/sequence_ops.cc): flat(0) = start; step = (stop - start) / (num - 1);
    flat(i) = start + step * i, all in T."""
    start, stop = _pair(start, stop)
    n = int(num._v) if isinstance(num, Tensor) else int(num)
    T = start._v.dtype.type
    if n == 1:
        return Tensor(np.array([start._v], T))
    step = T((stop._v - start._v) / T(n - 1))
    i = np.arange(n).astype(T)
    return Tensor((start._v + step * i).astype(T))


def meshgrid(*args, **kwargs):
    indexing = kwargs.get('indexing', 'xy')
    ts = [convert_to_tensor(a) for a in args]
    return [Tensor(np.ascontiguousarray(m)) for m in np.meshgrid(*[t._v for t in ts], indexing=indexing)]


def where(condition, x=None, y=None, name=None):
    c = convert_to_tensor(condition)._v
    if x is None and y is None:
        return Tensor(np.argwhere(c).astype(np.int64))
    x, y = _pair(x, y)
    if c.shape != x._v.shape and not (c.ndim == 1 and c.shape[0] == x._v.shape[0]):
        raise ValueError('tf.where: condition must have the shape of x (or be a vector over its first axis)')
    cb = c.reshape(c.shape + (1,) * (x._v.ndim - c.ndim))

    def vjp(g):   # SelectGrad
        z = np.zeros_like(g)
        return (np.where(cb, g, z), np.where(cb, z, g))
    return Tensor(np.where(cb, x._v, y._v), (x, y), vjp)


def gather(params, indices, axis=0, name=None):
    p, i = convert_to_tensor(params), convert_to_tensor(indices)

    def vjp(g):
        out = np.zeros_like(p._v)
        np.add.at(out, i._v, g)
        return (out, None)
    return Tensor(np.take(p._v, i._v, axis=axis), (p, i), vjp if axis == 0 else None)


def boolean_mask(tensor, mask, name='boolean_mask', axis=None):
    t, m = convert_to_tensor(tensor), convert_to_tensor(mask)._v.astype(np.bool_)
    if m.shape != t._v.shape[:m.ndim]:
        raise ValueError('boolean_mask: mask shape %s does not match tensor shape %s' % (m.shape, t._v.shape))

    def vjp(g):
        out = np.zeros_like(t._v)
        out[m] = g
        return (out,)
    return Tensor(t._v[m], (t,), vjp)


def one_hot(indices, depth, on_value=None, off_value=None, axis=None, dtype=None, name=None):
    idx = convert_to_tensor(indices)._v
    depth = int(depth._v) if isinstance(depth, Tensor) else int(depth)
    npd = _as_np_dtype(dtype) or np.dtype(np.float32)
    on = npd.type(1 if on_value is None else on_value)
    off = npd.type(0 if off_value is None else off_value)
    hot = np.equal(np.expand_dims(idx, -1), np.arange(depth))   # an out-of-range index matches nothing
    out = np.where(hot, on, off).astype(npd)
    if axis is not None and axis != -1:
        out = np.moveaxis(out, -1, int(axis))
    return Tensor(np.ascontiguousarray(out))


def argmax(x, axis=None, name=None, dimension=None, output_type=int64):
    ax = axis if axis is not None else (dimension if dimension is not None else 0)
    return Tensor(np.argmax(convert_to_tensor(x)._v, axis=int(ax)).astype(_as_np_dtype(output_type)))


def argmin(x, axis=None, name=None, dimension=None, output_type=int64):
    ax = axis if axis is not None else (dimension if dimension is not None else 0)
    return Tensor(np.argmin(convert_to_tensor(x)._v, axis=int(ax)).astype(_as_np_dtype(output_type)))


# ---- reductions -----------------------------------------------------------------------------------------------------
def _axes(axis, ndim):
    if axis is None:
        return tuple(_builtins.range(ndim))
    if isinstance(axis, Tensor):
        axis = axis._v.tolist()
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) % ndim for a in axis)
    return (int(axis) % ndim,)


def _keep_shape(shape, axes):   # noqa: A002
    return tuple(1 if i in axes else d for i, d in enumerate(shape))


def reduce_sum(x, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    x = convert_to_tensor(x)
    axes = _axes(axis if axis is not None else reduction_indices, x._v.ndim)
    keep = keepdims or builtins_bool(keep_dims)
    out = x._v.sum(axis=axes, keepdims=keep, dtype=_acc(x._v.dtype)).astype(x._v.dtype)
    src = x._v.shape

    def vjp(g):   # SumGrad: reshape to the kept shape and tile
        return (np.broadcast_to(g.reshape(_keep_shape(src, axes)), src).astype(g.dtype),)
    return Tensor(out, (x,), vjp)


def reduce_mean(x, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    x = convert_to_tensor(x)
    axes = _axes(axis if axis is not None else reduction_indices, x._v.ndim)
    keep = keepdims or builtins_bool(keep_dims)
    n = 1
    for a in axes:
        n *= x._v.shape[a]
    T = x._v.dtype.type
    with np.errstate(all='ignore'):
        out = (x._v.sum(axis=axes, keepdims=keep, dtype=_acc(x._v.dtype)) / max(n, 1) if n else
               np.full(_keep_shape(x._v.shape, axes) if keep else tuple(d for i, d in enumerate(x._v.shape) if i not in axes),
                       np.nan)).astype(x._v.dtype)
    src = x._v.shape

    def vjp(g):   # MeanGrad: SumGrad / factor
        return ((np.broadcast_to(g.reshape(_keep_shape(src, axes)), src) / T(n)).astype(g.dtype),)
    return Tensor(out, (x,), vjp)


def reduce_max(x, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    x = convert_to_tensor(x)
    axes = _axes(axis if axis is not None else reduction_indices, x._v.ndim)
    return Tensor(x._v.max(axis=axes, keepdims=keepdims or builtins_bool(keep_dims)))


def reduce_min(x, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    x = convert_to_tensor(x)
    axes = _axes(axis if axis is not None else reduction_indices, x._v.ndim)
    return Tensor(x._v.min(axis=axes, keepdims=keepdims or builtins_bool(keep_dims)))


def reduce_any(x, axis=None, keepdims=False, name=None):
    x = convert_to_tensor(x)
    return Tensor(x._v.any(axis=_axes(axis, x._v.ndim), keepdims=keepdims))


def reduce_all(x, axis=None, keepdims=False, name=None):
    x = convert_to_tensor(x)
    return Tensor(x._v.all(axis=_axes(axis, x._v.ndim), keepdims=keepdims))


# ---- graph-mode conveniences that are no-ops when every op is eager --------------------------------------------------
@contextlib.contextmanager
def name_scope(name=None, default_name=None, values=None):
    yield name


@contextlib.contextmanager
def control_dependencies(control_inputs):
    yield


class errors:   # noqa: N801
    class InvalidArgumentError(Exception):
        pass

    class OutOfRangeError(Exception):
        pass


def _assert(fn, what):
    def op(x, y, data=None, summarize=None, message=None, name=None):
        x, y = _pair(x, y)
        with np.errstate(all='ignore'):
            ok = fn(x._v, y._v)
        if not np.all(ok):
            raise errors.InvalidArgumentError('assertion failed: [Condition x %s y did not hold element-wise]' % what)
        return None
    return op


assert_less_equal = _assert(np.less_equal, '<=')
assert_less = _assert(np.less, '<')
assert_greater_equal = _assert(np.greater_equal, '>=')
assert_greater = _assert(np.greater, '>')
assert_equal = _assert(np.equal, '==')


class summary:   # noqa: N801
    @staticmethod
    def histogram(name, values, collections=None, family=None):
        return None

    @staticmethod
    def scalar(name, tensor, collections=None, family=None):
        return None

    @staticmethod
    def image(name, tensor, max_outputs=3, collections=None, family=None):
        return None


# ---- tf.nn ---------------------------------------------------------------------------------------------------------------
class nn:   # noqa: N801
    sigmoid = staticmethod(sigmoid)

    @staticmethod
    def relu(x, name=None):
        x = convert_to_tensor(x)
        return maximum(x, zeros_like(x))

    @staticmethod
    def softmax(logits, axis=None, name=None, dim=None):
        x = convert_to_tensor(logits)
        ax = -1 if axis is None and dim is None else int(axis if axis is not None else dim)
        e = np.exp(x._v - x._v.max(axis=ax, keepdims=True))
        out = (e / e.sum(axis=ax, keepdims=True, dtype=_acc(e.dtype)).astype(e.dtype)).astype(x._v.dtype)

        def vjp(g):   # SoftmaxGrad: (dy - sum(dy * y)) * y
            s = (g * out).sum(axis=ax, keepdims=True, dtype=_acc(g.dtype)).astype(g.dtype)
            return ((g - s) * out,)
        return Tensor(out, (x,), vjp)

    @staticmethod
    def sigmoid_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
        """nn_impl.sigmoid_cross_entropy_with_logits: relu(x) - x * z + log1p(exp(-|x|)), composed of the
        same primitive ops (so tf.gradients differentiates it the way TF does)."""
        logits, labels = _pair(logits, labels)
        zeros_ = zeros_like(logits)
        cond = greater_equal(logits, zeros_)
        relu_logits = where(cond, logits, zeros_)
        neg_abs_logits = where(cond, negative(logits), logits)
        return add(subtract(relu_logits, multiply(logits, labels)), log1p(exp(neg_abs_logits)))


# ---- tf.losses --------------------------------------------------------------------------------------------------------------
class losses:   # noqa: N801
    class Reduction:
        NONE = 'none'
        SUM = 'weighted_sum'
        MEAN = 'weighted_mean'
        SUM_OVER_BATCH_SIZE = 'weighted_sum_over_batch_size'
        SUM_BY_NONZERO_WEIGHTS = 'weighted_sum_by_nonzero_weights'

    @staticmethod
    def compute_weighted_loss(losses_, weights=1.0, reduction='weighted_sum_by_nonzero_weights'):
        losses_ = to_float(losses_)
        weights = to_float(convert_to_tensor(weights))
        weighted = multiply(losses_, weights)
        if reduction == 'none':
            return weighted
        loss = reduce_sum(weighted)
        if reduction == 'weighted_sum':
            return loss
        if reduction == 'weighted_mean':
            return _safe_div(loss, reduce_sum(multiply(ones_like(losses_), weights)))
        if reduction == 'weighted_sum_over_batch_size':
            return _safe_div(loss, to_float(size(losses_)))
        # SUM_BY_NONZERO_WEIGHTS: divide by the number of loss elements whose (broadcast) weight is non-zero
        present = where(equal(weights, 0.0), zeros_like(weights), ones_like(weights))
        num_present = reduce_sum(multiply(ones_like(losses_), present))
        return _safe_div(loss, num_present)

    @staticmethod
    def huber_loss(labels, predictions, weights=1.0, delta=1.0, scope=None, loss_collection=None,
                   reduction='weighted_sum_by_nonzero_weights'):
        predictions = to_float(predictions)
        labels = to_float(labels)
        error = subtract(predictions, labels)
        abs_error = abs(error)
        quadratic = minimum(abs_error, delta)
        linear = subtract(abs_error, quadratic)   # == max(|e| - delta, 0), but with a usable gradient
        l = add(multiply(0.5, multiply(quadratic, quadratic)), multiply(delta, linear))
        return losses.compute_weighted_loss(l, weights, reduction)


def _safe_div(num, den):
    """losses_impl._safe_div: 0 where the denominator is 0."""
    return where(greater(den, 0.0), truediv(num, where(equal(den, 0.0), ones_like(den), den)), zeros_like(num))


# ---- tf.image ---------------------------------------------------------------------------------------------------------------
class image:   # noqa: N801
    @staticmethod
    def non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5, name=None):
        """NonMaxSuppressionOp (CPU): visit candidates in descending score (stable: equal scores in index order);
        keep one unless its IoU with an already kept box is > iou_threshold; stop at max_output_size."""
        b = convert_to_tensor(boxes)._v.astype(np.float32)
        s = convert_to_tensor(scores)._v.astype(np.float32)
        limit = int(max_output_size._v) if isinstance(max_output_size, Tensor) else int(max_output_size)
        thr = np.float32(iou_threshold)
        order = np.argsort(-s, kind='stable')
        ymin = np.minimum(b[:, 0], b[:, 2])
        xmin = np.minimum(b[:, 1], b[:, 3])
        ymax = np.maximum(b[:, 0], b[:, 2])
        xmax = np.maximum(b[:, 1], b[:, 3])
        area = (ymax - ymin) * (xmax - xmin)
        keep = []
        for i in order:
            if len(keep) >= limit:
                break
            ok = True
            for j in reversed(keep):   # the kernel scans the kept boxes from the most recent one
                if area[i] <= 0 or area[j] <= 0:
                    continue
                iy = max(np.float32(0), min(ymax[i], ymax[j]) - max(ymin[i], ymin[j]))
                ix = max(np.float32(0), min(xmax[i], xmax[j]) - max(xmin[i], xmin[j]))
                inter = np.float32(iy * ix)
                if inter / np.float32(np.float32(area[i] + area[j]) - inter) > thr:
                    ok = False
                    break
            if ok:
                keep.append(int(i))
        return Tensor(np.array(keep, np.int32))


# ---- tf.gradients -------------------------------------------------------------------------------------------------------------
def gradients(ys, xs, grad_ys=None, name=None):
    """Reverse-mode over the tape; d(sum of ys)/d(x) for every x (None when unconnected), like tf.gradients."""
    ys = ys if isinstance(ys, (list, tuple)) else [ys]
    single = not isinstance(xs, (list, tuple))
    xs_l = [xs] if single else list(xs)
    order, seen = [], set()

    def visit(t):
        stack_ = [(t, False)]
        while stack_:
            node, done = stack_.pop()
            if done:
                order.append(node)
                continue
            if id(node) in seen:
                continue
            seen.add(id(node))
            stack_.append((node, True))
            for p in node._parents:
                stack_.append((p, False))
    for y in ys:
        visit(y)
    grads = {}
    for k, y in enumerate(ys):
        g0 = np.ones_like(y._v) if grad_ys is None or grad_ys[k] is None else convert_to_tensor(grad_ys[k])._v
        grads[id(y)] = grads.get(id(y), 0) + g0
    for node in reversed(order):
        g = grads.get(id(node))
        if g is None or node._vjp is None or not node._parents:
            continue
        for p, gp in zip(node._parents, node._vjp(np.asarray(g, dtype=node._v.dtype))):
            if gp is None:
                continue
            gp = np.asarray(gp, dtype=p._v.dtype)
            grads[id(p)] = gp if id(p) not in grads else grads[id(p)] + gp
    out = [None if id(x) not in grads else Tensor(np.asarray(grads[id(x)])) for x in xs_l]
    return out


# ---- tf.test ---------------------------------------------------------------------------------------------------------------------
def _evaluate(x):
    if isinstance(x, Tensor):
        return x._v
    if isinstance(x, dict):
        return {k: _evaluate(v) for k, v in x.items()}
    if isinstance(x, tuple) and hasattr(x, '_fields'):
        return type(x)(*[_evaluate(v) for v in x])
    if isinstance(x, (list, tuple)):
        return type(x)(_evaluate(v) for v in x)
    return x


class test:   # noqa: N801
    class TestCase(unittest.TestCase):
        def evaluate(self, tensors):
            return _evaluate(tensors)

    @staticmethod
    def main():
        unittest.main()


class Session:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        return _evaluate(fetches)
