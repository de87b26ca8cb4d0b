"""TEST INFRASTRUCTURE ONLY — minimal pure-Python emulation of the Taichi API surface used by the
reference renderer (see ../README.md).  Semantics emulated:

* `@ti.kernel`, `@ti.func`, `@ti.data_oriented`: identity decorators (bodies run as Python).
* `@ti.dataclass`: a struct class with zero-initialised annotated members, positional/keyword
  construction, methods kept, and `.field(shape=...)` giving a dense container of independent
  instances (element access returns the stored instance: member assignment writes through, which
  is what the reference's python-scope `node.bound = ...` relies on, scene.py:287,301).
* vectors / matrices: NumPy-backed `TiArray` with .x/.y/.z/.w, swizzles, `@`, and NON-mutating
  augmented assignment (Taichi copies on `a = b`, so `color = self.color; color += ...` at
  gaussian.py:199-200 must not modify the struct member).
* fields: `ti.field(dtype, shape)` with from_numpy / to_numpy / struct-for iteration; scalar field
  elements are read through a numeric proxy so that `ti.atomic_add(field[i, j], 1)` can update them.
* `MatrixField.from_numpy` with a same-size but differently shaped array reinterprets the flat
  C-order buffer (this is the "taichi_as_executed" SH layout of SURVEY.md §7 hard part 7; what real
  Taichi does here could not be verified offline).
"""
from __future__ import annotations

import itertools
import math as _pymath
import os

import numpy as np

from . import math  # noqa: F401  (ti.math)
from .math import TiArray, _FP, _VecType, vec2, vec3, vec4  # noqa: F401

f32 = "f32"
i32 = "i32"
f64 = "f64"
gpu = "gpu"
cpu = "cpu"


class _Cfg:
    arch = "python-shim"


cfg = _Cfg()


def init(*args, **kwargs):
    return None


def kernel(fn):
    return fn


def func(fn):
    return fn


def data_oriented(cls):
    return cls


def static(x):
    return x


def sqrt(x):
    return np.sqrt(x)


def random(dtype=None):
    return np.random.random()


def _reduce(fn, args):
    out = args[0]
    for a in args[1:]:
        out = fn(out, a)
    if isinstance(out, np.ndarray) and not isinstance(out, TiArray):
        out = out.view(TiArray)
    return out


def min(*args):  # noqa: A001
    return _reduce(np.minimum, [_unwrap(a) for a in args])


def max(*args):  # noqa: A001
    return _reduce(np.maximum, [_unwrap(a) for a in args])


def ndrange(*dims):
    ranges = [range(d[0], d[1]) if isinstance(d, tuple) else range(int(d)) for d in dims]
    return itertools.product(*ranges)


def grouped(field):
    return iter(field._indices())


# ------------------------------------------------------------------------------ scalar proxy
class _Ref:
    """Numeric proxy for a scalar field element (so atomic_add can write back)."""
    __slots__ = ("_f", "_i")

    def __init__(self, f, i):
        self._f, self._i = f, i

    @property
    def v(self):
        return self._f._data[self._i]

    def __index__(self):
        return int(self.v)

    def __int__(self):
        return int(self.v)

    def __float__(self):
        return float(self.v)

    def __bool__(self):
        return bool(self.v)

    def __repr__(self):
        return repr(self.v)

    def __format__(self, spec):
        return format(self.v, spec)

    def __hash__(self):
        return hash(self.v)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.v, dtype=dtype)


def _unwrap(x):
    return x.v if isinstance(x, _Ref) else x


def _binop(name):
    def op(self, other):
        return getattr(self.v, name)(_unwrap(other))
    return op


for _n in ("add", "sub", "mul", "truediv", "floordiv", "mod", "pow", "radd", "rsub", "rmul", "rtruediv",
           "rfloordiv", "rpow", "lt", "le", "gt", "ge", "eq", "ne"):
    setattr(_Ref, f"__{_n}__", _binop(f"__{_n}__"))
_Ref.__neg__ = lambda self: -self.v


def atomic_add(ref, val):
    if not isinstance(ref, _Ref):
        raise TypeError("taichi_shim.atomic_add needs a scalar field element")
    old = ref.v
    ref._f._data[ref._i] = old + _unwrap(val)
    return old


# ------------------------------------------------------------------------------ fields
def _norm_shape(shape):
    if isinstance(shape, int):
        return (shape,)
    return tuple(int(s) for s in shape)


class _BaseField:
    shape: tuple

    def _indices(self):
        return itertools.product(*[range(s) for s in self.shape])

    def __iter__(self):
        if len(self.shape) == 1:
            return iter(range(self.shape[0]))
        return self._indices()

    @staticmethod
    def _key(idx):
        if isinstance(idx, tuple):
            return tuple(int(_unwrap(i)) for i in idx)
        return (int(_unwrap(idx)),)


class ScalarField(_BaseField):
    def __init__(self, dtype, shape):
        self.shape = _norm_shape(shape)
        self._np = np.int64 if dtype in (i32, int) else _FP()
        self._data = np.zeros(self.shape, dtype=self._np)

    def __getitem__(self, idx):
        return _Ref(self, self._key(idx))

    def __setitem__(self, idx, val):
        self._data[self._key(idx)] = _unwrap(val)

    def from_numpy(self, arr):
        self._data[...] = np.asarray(arr).reshape(self.shape)

    def to_numpy(self):
        out = self._data.copy()
        return out.astype(np.int32) if self._np is np.int64 else out.astype(np.float32)


class MatrixField(_BaseField):
    """Field of vectors (ti.field(ti.math.vec3, shape))."""

    def __init__(self, vtype, shape):
        self.shape = _norm_shape(shape)
        self.n = vtype.n
        self._np = np.int64 if vtype.dtype is np.int64 else _FP()
        self._data = np.zeros(self.shape + (self.n,), dtype=self._np)

    def __getitem__(self, idx):
        return self._data[self._key(idx)].copy().view(TiArray)

    def __setitem__(self, idx, val):
        self._data[self._key(idx)] = np.asarray(_unwrap(val))

    def from_numpy(self, arr):
        arr = np.asarray(arr)
        if arr.size != self._data.size:
            raise ValueError(f"from_numpy: size mismatch {arr.shape} -> {self._data.shape}")
        # same size, possibly different shape: reinterpret the flat C-order buffer (see module docstring)
        self._data[...] = np.ascontiguousarray(arr).reshape(self._data.shape)

    def to_numpy(self):
        return self._data.astype(np.float32 if self._np is not np.int64 else np.int32)


class StructField(_BaseField):
    def __init__(self, cls, shape):
        self.shape = _norm_shape(shape)
        self._cls = cls
        self._data = np.empty(self.shape, dtype=object)
        for i in self._indices():
            self._data[i] = cls()

    def __getitem__(self, idx):
        if self.shape == ():
            return self._data[()]
        return self._data[self._key(idx)]

    def __setitem__(self, idx, val):
        self._data[self._key(idx)] = val


Field = _BaseField


def field(dtype, shape):
    if isinstance(dtype, _VecType):
        return MatrixField(dtype, shape)
    return ScalarField(dtype, shape)


class _Types:
    @staticmethod
    def vector(n, dtype):
        return _VecType(n, np.int64 if dtype in (i32, int) else None)


types = _Types()


# ------------------------------------------------------------------------------ dataclass
def _zero_of(ann):
    if isinstance(ann, _VecType):
        return ann(0)
    if isinstance(ann, type) and getattr(ann, "_ti_struct", False):
        return ann()
    if ann in (i32, int):
        return 0
    return _FP()(0.0)


def dataclass(cls):
    anns = dict(getattr(cls, "__annotations__", {}))
    names = list(anns)
    user_init = cls.__dict__.get("__init__")

    def __init__(self, *args, **kwargs):
        for n in names:
            object.__setattr__(self, n, _zero_of(anns[n]))
        if user_init is not None:
            user_init(self, *args, **kwargs)
            return
        if len(args) > len(names):
            raise TypeError(f"{cls.__name__}: too many arguments")
        for n, v in zip(names, args):
            setattr(self, n, _coerce(anns[n], v))
        for n, v in kwargs.items():
            if n not in anns:
                raise TypeError(f"{cls.__name__}: unknown member {n}")
            setattr(self, n, _coerce(anns[n], v))

    def __repr__(self):
        return f"{cls.__name__}(" + ", ".join(f"{n}={getattr(self, n)!r}" for n in names) + ")"

    cls.__init__ = __init__
    cls.__repr__ = __repr__
    cls.__setattr__ = lambda self, n, v: object.__setattr__(self, n, _unwrap(v))
    cls._ti_struct = True
    cls._ti_members = names
    cls.field = classmethod(lambda c, shape: StructField(c, shape))
    return cls


def _coerce(ann, v):
    v = _unwrap(v)
    if isinstance(ann, _VecType):
        return ann(v)
    if ann in (i32, int):
        return int(v)
    if ann in (f32, f64, float):
        return _FP()(v)
    return v
