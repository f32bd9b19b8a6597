"""
tools/tfshim.py -- a numpy stand-in for the TensorFlow/Keras ops that the reference's hot
path calls, so that the reference's OWN python source (imported unmodified from
/root/reference) can be executed in this image, where TensorFlow is not installed.

Used only by tools/gen_golden.py to produce tests/golden/*.npz.  It is not part of the
product, not part of the oracle, and never runs on the GPU box.

Design
  * `Tensor` is an immutable wrapper around an np.ndarray (TF tensors are immutable: the
    reference relies on `x *= y` rebinding, utils.py:1090-1091).
  * every op computes in the array's own dtype and returns a fresh fp32/int32 array, i.e.
    one rounding per op, like eager TF on CPU.
  * anything the hot path does not need (Keras layers, callbacks, pystrum, matplotlib, ...)
    resolves to inert dummies so that `import neurite` succeeds.
"""
import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np


# ---------------------------------------------------------------------------------------
# Tensor
# ---------------------------------------------------------------------------------------
class TensorShape(tuple):
    def as_list(self):
        return list(self)

    def __getitem__(self, k):
        r = tuple.__getitem__(self, k)
        return TensorShape(r) if isinstance(k, slice) else r

    def __add__(self, other):
        return TensorShape(tuple(self) + tuple(other))

    def __radd__(self, other):
        return TensorShape(tuple(other) + tuple(self))


class Dimension(int):
    pass


class DType:
    def __init__(self, np_dtype):
        self.np = np.dtype(np_dtype)

    @property
    def as_numpy_dtype(self):
        return self.np.type

    @property
    def is_floating(self):
        return np.issubdtype(self.np, np.floating)

    @property
    def is_integer(self):
        return np.issubdtype(self.np, np.integer)

    @property
    def base_dtype(self):
        return self

    @property
    def name(self):
        return self.np.name

    def __eq__(self, other):
        try:
            return self.np == _np_dtype(other)
        except TypeError:
            return False

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.np)

    def __repr__(self):
        return 'tf.' + self.np.name


def _np_dtype(d):
    if isinstance(d, DType):
        return d.np
    if d is bool or d == 'bool':
        return np.dtype(bool)
    return np.dtype(d)


def A(x):
    """unwrap to ndarray (python scalars become 0-d arrays of the weak python type)."""
    if isinstance(x, Tensor):
        return x._a
    if isinstance(x, (list, tuple)):
        return np.asarray([A(v) for v in x])
    return x


class Tensor:
    __array_priority__ = 1000

    def __init__(self, a):
        a = np.asarray(a)
        if a.dtype == np.float64 and False:
            a = a.astype(np.float32)
        self._a = a

    # --- introspection
    @property
    def shape(self):
        return TensorShape(self._a.shape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return DType(self._a.dtype)

    @property
    def ndim(self):
        return self._a.ndim

    def numpy(self):
        return self._a

    def __len__(self):
        return len(self._a)

    def __iter__(self):
        return (Tensor(v) for v in self._a)

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __repr__(self):
        return 'shim.Tensor(%r)' % (self._a,)

    def __bool__(self):
        return bool(self._a)

    def __index__(self):
        return int(self._a)

    def __int__(self):
        return int(self._a)

    def __float__(self):
        return float(self._a)

    # --- indexing (the reference indexes with *lists* of slices, layers.py:1186)
    def __getitem__(self, k):
        if isinstance(k, list):
            k = tuple(k)
        if isinstance(k, tuple):
            k = tuple(A(v) for v in k)
        else:
            k = A(k)
        return Tensor(self._a[k])

    # --- arithmetic: same-dtype numpy ops, one rounding per op; no in-place variants
    def _bin(self, other, fn, swap=False):
        o = A(other)
        if isinstance(o, np.ndarray) and o.dtype != self._a.dtype and o.dtype.kind == self._a.dtype.kind == 'f':
            raise TypeError('shim: float dtype mismatch %s vs %s (TF would raise)' % (self._a.dtype, o.dtype))
        return Tensor(fn(o, self._a) if swap else fn(self._a, o))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.true_divide)
    def __rtruediv__(self, o): return self._bin(o, np.true_divide, True)
    def __neg__(self): return Tensor(-self._a)
    def __lt__(self, o): return self._bin(o, np.less)
    def __le__(self, o): return self._bin(o, np.less_equal)
    def __gt__(self, o): return self._bin(o, np.greater)
    def __ge__(self, o): return self._bin(o, np.greater_equal)
    def __eq__(self, o): return self._bin(o, np.equal)
    def __ne__(self, o): return self._bin(o, np.not_equal)
    __hash__ = object.__hash__
    __hash__ = object.__hash__


def T(x):
    return x if isinstance(x, Tensor) else Tensor(np.asarray(A(x)))


# ---------------------------------------------------------------------------------------
# inert dummies for everything off the hot path
# ---------------------------------------------------------------------------------------
class _Inert:
    """Base class / decorator / attribute sink."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and not k and callable(a[0]) and not isinstance(a[0], _Inert):
            return a[0]                       # used as a decorator
        return _Inert()

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Inert()

    def __iter__(self):
        return iter(())


class _InertModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        if name and name[0].isupper():
            cls = type(name, (_InertBase,), {})
            setattr(self, name, cls)
            return cls
        return _Inert()


class _InertBase:
    """usable as a base class (`class Foo(Layer)`) and constructible with any args."""

    def __init__(self, *a, **k):
        pass

    def __init_subclass__(cls, **k):
        pass


_SHIM_ROOTS = ('tensorflow', 'pystrum', 'matplotlib', 'nibabel', 'h5py', 'skimage', 'voxelmorph',
               'keras', 'mpl_toolkits')


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split('.')[0] in _SHIM_ROOTS and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _InertModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        if module.__name__ == 'pystrum':
            module.__version__ = '0.4'


# ---------------------------------------------------------------------------------------
# the ops the hot path uses
# ---------------------------------------------------------------------------------------
def _install_tf():
    tf = _InertModule('tensorflow')
    tf.__path__ = []
    tf.__version__ = '2.shim'
    tf.Tensor = Tensor
    tf.TensorShape = TensorShape
    for n in ('float32', 'float64', 'int32', 'int64', 'bool', 'float16'):
        setattr(tf, n, DType(bool if n == 'bool' else n))

    def cast(x, dtype):
        return Tensor(A(T(x)).astype(_np_dtype(dtype)))     # float->int truncates like TF
    tf.cast = cast
    tf.convert_to_tensor = lambda x, dtype=None, **k: T(x) if dtype is None else cast(T(x), dtype)
    tf.constant = lambda v, dtype=None, **k: Tensor(np.asarray(A(v), dtype=None if dtype is None else _np_dtype(dtype)))
    tf.stack = lambda vals, axis=0, **k: T(vals) if isinstance(vals, Tensor) else Tensor(np.stack([A(T(v)) for v in vals], axis=axis))
    tf.concat = lambda vals, axis, **k: Tensor(np.concatenate([A(T(v)) for v in vals], axis=A(axis)))
    tf.reshape = lambda x, shape, **k: Tensor(np.reshape(A(T(x)), [int(s) for s in np.ravel(A(shape))] if not isinstance(shape, int) else shape))
    tf.floor = lambda x: Tensor(np.floor(A(x)))
    tf.round = lambda x: Tensor(np.rint(A(x)))              # half-to-even, like tf.round
    tf.clip_by_value = lambda x, lo, hi: Tensor(np.clip(A(x), np.asarray(lo).astype(A(x).dtype), np.asarray(hi).astype(A(x).dtype)))
    tf.gather = lambda params, indices, **k: Tensor(A(params)[A(indices)])
    tf.less = lambda a, b: Tensor(np.less(A(a), A(b)))
    tf.greater = lambda a, b: Tensor(np.greater(A(a), A(b)))
    tf.logical_not = lambda a: Tensor(np.logical_not(A(a)))
    tf.reduce_any = lambda x, axis=None, keepdims=False: Tensor(np.any(A(x), axis=axis, keepdims=keepdims))
    tf.range = lambda *a, dtype=None, **k: Tensor(np.arange(*[A(v) for v in a], dtype=np.int32 if dtype is None else _np_dtype(dtype)))
    tf.size = lambda x: Tensor(np.int32(A(T(x)).size))
    tf.tile = lambda x, multiples: Tensor(np.tile(A(x), [int(m) for m in np.ravel(A(multiples))]))
    tf.ones = lambda shape, dtype=None: Tensor(np.ones(shape, dtype=np.float32 if dtype is None else _np_dtype(dtype)))
    tf.zeros = lambda shape, dtype=None: Tensor(np.zeros(shape, dtype=np.float32 if dtype is None else _np_dtype(dtype)))

    def linspace(start, stop, num):
        # tf.linspace (math_ops.linspace_nd, TF >= 2.3) in fp32: endpoints exact,
        # interior start + delta*i.  Third-party arithmetic; see oracle/interp.py.
        start, stop, num = np.float32(start), np.float32(stop), int(num)
        if num == 1:
            return Tensor(np.array([start], np.float32))
        delta = np.float32(stop - start) / np.float32(num - 1)
        i = np.arange(1, num - 1, dtype=np.int64).astype(np.float32)
        return Tensor(np.concatenate([[start], start + delta * i, [stop]]).astype(np.float32))
    tf.linspace = linspace
    tf.map_fn = lambda fn, elems, **k: Tensor(np.stack([A(fn(e)) for e in (elems if isinstance(elems, Tensor) else zip(*elems))], 0))


    # ---- additions for MutualInformation / soft_quantize / gaussian_kernel / separable_conv ----
    tf.newaxis = None
    tf.transpose = lambda x, perm=None, **k: Tensor(np.transpose(A(T(x)), perm))
    tf.minimum = lambda a, b: Tensor(np.minimum(A(T(a)), A(T(b))))
    tf.maximum = lambda a, b: Tensor(np.maximum(A(T(a)), A(T(b))))
    tf.square = lambda x: Tensor(np.square(np.asarray(A(T(x)), dtype=np.float32 if isinstance(x, float) else None)))
    tf.exp = lambda x: Tensor(np.exp(A(x)))
    tf.shape = lambda x: Tensor(np.asarray(A(T(x)).shape, dtype=np.int32))
    tf.expand_dims = lambda x, axis: Tensor(np.expand_dims(A(T(x)), axis))
    tf.is_tensor = lambda x: isinstance(x, Tensor)
    _red = lambda fn: (lambda x, axis=None, keepdims=False: Tensor(
        fn(A(T(x)), axis=axis, keepdims=keepdims, dtype=np.float64).astype(A(T(x)).dtype)
        if A(T(x)).dtype.kind == 'f' else fn(A(T(x)), axis=axis, keepdims=keepdims).astype(A(T(x)).dtype)))
    tf.reduce_sum = _red(np.sum)
    tf.reduce_mean = _red(np.mean)
    tf.reduce_prod = _red(np.prod)
    dtypes = _InertModule('tensorflow.dtypes')
    dtypes.as_dtype = lambda d: d if isinstance(d, DType) else DType(d)
    tf.dtypes = dtypes
    exp_mod = _InertModule('tensorflow.experimental')
    exp_np = _InertModule('tensorflow.experimental.numpy')
    exp_np.diff = lambda x, **k: Tensor(np.diff(A(T(x))))
    exp_mod.numpy = exp_np
    tf.experimental = exp_mod

    def convolution(x, k, padding='VALID', strides=None, dilations=None, **kw):
        """tf.nn.convolution (third party; restated): N-D cross-correlation, one in/out feature,
        zero 'SAME' padding with the extra element at the end; float64 accumulate, one rounding."""
        x, k = A(T(x)), A(T(k))
        nd = x.ndim - 2
        assert x.shape[-1] == 1 and k.shape[-2:] == (1, 1)
        ks = k.shape[:nd]
        st = [1] * nd if strides is None else [int(v) for v in strides]
        dl = [1] * nd if dilations is None else [int(v) for v in dilations]
        if any(a > 1 and b > 1 for a, b in zip(st, dl)):
            raise ValueError('strides > 1 not supported in conjunction with dilation_rate > 1')
        outs, pads = [], []
        for n, kk, s_, d_ in zip(x.shape[1:-1], ks, st, dl):
            eff = (kk - 1) * d_ + 1
            if padding.upper() == 'SAME':
                o = -(-n // s_)
                tot = max((o - 1) * s_ + eff - n, 0)
                pads.append((tot // 2, tot - tot // 2))
            else:
                o = max(-(-(n - eff + 1) // s_), 0)
                pads.append((0, 0))
            outs.append(o)
        xp = np.pad(x[..., 0].astype(np.float64), [(0, 0)] + pads)
        out = np.zeros((x.shape[0],) + tuple(outs), np.float64)
        for tap in np.ndindex(*ks):
            sl = tuple(slice(t * d_, t * d_ + (o - 1) * s_ + 1, s_) for t, d_, o, s_ in zip(tap, dl, outs, st))
            out += np.float64(k[tap + (0, 0)]) * xp[(slice(None),) + sl]
        return Tensor(out.astype(x.dtype)[..., None])
    nn = _InertModule('tensorflow.nn')
    nn.convolution = convolution
    tf.nn = nn

    # tf.math / tf.debugging
    tfmath = _InertModule('tensorflow.math')

    def divide_no_nan(a, b):
        a, b = A(a), A(b)
        out = np.zeros(np.broadcast(a, b).shape, dtype=a.dtype)
        np.divide(a, b, out=out, where=(b != 0))
        return Tensor(out)
    tfmath.divide_no_nan = divide_no_nan
    tf.math = tfmath

    class InvalidArgumentError(Exception):
        pass
    errors = _InertModule('tensorflow.errors')
    errors.InvalidArgumentError = InvalidArgumentError
    tf.errors = errors
    dbg = _InertModule('tensorflow.debugging')

    def _assert(ok, msg):
        if not bool(np.all(ok)):
            raise InvalidArgumentError(msg)
    dbg.assert_greater_equal = lambda x, y, message=None, **k: _assert(A(x) >= A(y), message)
    dbg.assert_less_equal = lambda x, y, message=None, **k: _assert(A(x) <= A(y), message)
    dbg.assert_all_finite = lambda x, message=None, **k: _assert(np.isfinite(A(x)), message)
    dbg.assert_equal = lambda x, y, message=None, **k: _assert(np.array_equal(A(T(x)), A(T(y))), message)
    dbg.assert_non_negative = lambda x, message=None, **k: _assert(A(T(x)) >= 0, message)
    dbg.assert_greater = lambda x, y, message=None, **k: _assert(A(T(x)) > A(T(y)), message)
    tf.debugging = dbg

    compat = _InertModule('tensorflow.compat')
    v1 = _InertModule('tensorflow.compat.v1')
    v1.Dimension = Dimension
    compat.v1 = v1
    tf.compat = compat

    # keras backend
    K = _InertModule('tensorflow.keras.backend')
    K.expand_dims = lambda x, axis=-1: Tensor(np.expand_dims(A(x), axis))
    K.reshape = tf.reshape
    K.shape = lambda x: Tensor(np.asarray(A(x).shape, dtype=np.int32))
    K.int_shape = lambda x: tuple(A(x).shape)
    K.ndim = lambda x: A(x).ndim
    # reductions: float64 accumulate, one rounding -- TF's reduction order is unspecified
    _ax = lambda axis: tuple(axis) if isinstance(axis, list) else axis
    K.sum = lambda x, axis=None, keepdims=False: Tensor(np.sum(A(x), axis=_ax(axis), keepdims=keepdims, dtype=np.float64).astype(A(x).dtype))
    K.mean = lambda x, axis=None, keepdims=False: Tensor(np.mean(A(x), axis=axis, keepdims=keepdims, dtype=np.float64).astype(A(x).dtype))
    K.square = lambda x: Tensor(np.square(A(x)))
    K.min = lambda x, axis=None, keepdims=False: Tensor(np.min(A(x), axis=axis, keepdims=keepdims))
    K.max = lambda x, axis=None, keepdims=False: Tensor(np.max(A(x), axis=axis, keepdims=keepdims))
    K.exp = lambda x: Tensor(np.exp(A(x)))
    K.log = lambda x: Tensor(np.log(A(x)))
    K.epsilon = lambda: 1e-7
    K.flatten = lambda x: Tensor(np.reshape(A(x), -1))
    K.stack = lambda xs, axis=0: Tensor(np.stack([A(T(v)) for v in xs], axis=axis))
    K.argmax = lambda x, axis=-1: Tensor(np.argmax(A(x), axis=axis).astype(np.int64))

    def one_hot(idx, n):
        idx = A(idx).astype(np.int64)
        out = np.zeros(idx.shape + (int(n),), np.float32)
        ok = (idx >= 0) & (idx < n)
        np.put_along_axis(out, np.where(ok, idx, 0)[..., None], ok[..., None].astype(np.float32), axis=-1)
        return Tensor(out)
    K.one_hot = one_hot
    K.concatenate = lambda xs, axis=-1: Tensor(np.concatenate([A(x) for x in xs], axis=axis))
    K.batch_dot = lambda x, y, axes=None: Tensor(np.matmul(A(x), A(y)))     # [P,B,F] x [P,F,C]
    K.permute_dimensions = lambda x, perm: Tensor(np.transpose(A(x), perm))

    def bias_add(x, bias, data_format=None):
        # tf.keras.backend.bias_add with an N-D bias (ndim(x) - 1 dims): channels_last adds reshape(bias, (1,) + shape);
        # channels_first adds reshape(bias, (1, shape[-1]) + shape[:-1]) -- a RAW reshape of the [..., C] weight, not a
        # transpose (keras/backend.py).  Weights trained by Keras carry that layout, so it is the contract.
        b = A(bias)
        if data_format == 'channels_first':
            b = b.reshape((b.shape[-1],) + b.shape[:-1])
        return Tensor(A(x) + b[None])
    K.bias_add = bias_add
    K.image_data_format = lambda: 'channels_last'

    keras = _InertModule('tensorflow.keras')
    keras.__path__ = []
    keras.backend = K
    layers = _InertModule('tensorflow.keras.layers')

    class Layer(_InertBase):
        def get_config(self):
            return {}

        def build(self, input_shape):
            self.built = True

        def add_weight(self, shape=None, **k):
            return None

        def __call__(self, inputs):
            if not getattr(self, 'built', False):
                shp = [tuple(A(T(i)).shape) for i in inputs] if isinstance(inputs, (list, tuple)) \
                    else tuple(A(T(inputs)).shape)
                self.build(shp)
            return self.call(inputs)
    layers.Layer = Layer
    keras.layers = layers

    losses = _InertModule('tensorflow.keras.losses')

    class CategoricalCrossentropy:
        """Keras formula (third party; restated, not executed reference code)."""

        def __init__(self, from_logits=False, label_smoothing=0., axis=-1,
                     reduction='sum_over_batch_size', name='categorical_crossentropy'):
            assert not from_logits and not label_smoothing and axis == -1
            self.reduction = reduction

        def __call__(self, y_true, y_pred, sample_weight=None):
            t, p = A(T(y_true)), A(T(y_pred))
            p = p / np.sum(p, axis=-1, keepdims=True, dtype=np.float64).astype(p.dtype)
            eps = np.float32(1e-7)
            p = np.clip(p, eps, np.float32(1) - eps)
            loss = (-np.sum(t.astype(np.float64) * np.log(p.astype(np.float64)), axis=-1)).astype(np.float32)
            if sample_weight is not None:
                loss = loss * A(sample_weight)
            return Tensor(np.float32(np.sum(loss, dtype=np.float64) / loss.size))
    losses.CategoricalCrossentropy = CategoricalCrossentropy

    class MeanSquaredError(_InertBase):
        pass
    losses.MeanSquaredError = MeanSquaredError
    keras.losses = losses
    tf.keras = keras

    mods = {
        'tensorflow': tf, 'tensorflow.math': tfmath, 'tensorflow.debugging': dbg,
        'tensorflow.errors': errors, 'tensorflow.compat': compat, 'tensorflow.compat.v1': v1,
        'tensorflow.keras': keras, 'tensorflow.keras.backend': K,
        'tensorflow.keras.layers': layers, 'tensorflow.keras.losses': losses,
        'tensorflow.nn': nn, 'tensorflow.dtypes': dtypes, 'tensorflow.experimental': exp_mod,
        'tensorflow.experimental.numpy': exp_np,
    }
    sys.modules.update(mods)

    # keras-internal helpers the layers module imports by name
    cu = _InertModule('tensorflow.python.keras.utils.conv_utils')

    def normalize_tuple(value, n, name):
        return (value,) * n if isinstance(value, int) else tuple(value)
    cu.normalize_tuple = normalize_tuple
    cu.normalize_padding = lambda p: p.lower()
    cu.normalize_data_format = lambda d: 'channels_last' if d is None else d.lower()

    def conv_output_length(input_length, filter_size, padding, stride, dilation=1):
        if input_length is None:
            return None
        out = input_length if padding == 'same' else input_length - filter_size + 1
        return (out + stride - 1) // stride
    cu.conv_output_length = conv_output_length
    tu = _InertModule('tensorflow.python.keras.utils.tf_utils')
    tu.shape_type_conversion = lambda fn: fn
    for name, m in (('tensorflow.python', None), ('tensorflow.python.keras', None),
                    ('tensorflow.python.keras.utils', None),
                    ('tensorflow.python.keras.utils.conv_utils', cu),
                    ('tensorflow.python.keras.utils.tf_utils', tu)):
        if m is None:
            m = _InertModule(name)
            m.__path__ = []
        sys.modules[name] = m
    sys.modules['tensorflow.python.keras.utils'].conv_utils = cu
    sys.modules['tensorflow.python.keras.utils'].tf_utils = tu
    return tf


def install(reference_root='/root/reference'):
    """Install the shim and return the imported reference package `neurite`."""
    if 'neurite' in sys.modules:
        return sys.modules['neurite']
    _install_tf()
    sys.meta_path.insert(0, _Finder())
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    sys.dont_write_bytecode = True            # /root/reference is read-only
    import neurite
    return neurite
