"""ti.math subset (TEST INFRASTRUCTURE ONLY — see package docstring)."""
from __future__ import annotations

import os

import numpy as np

inf = float("inf")
pi = float(np.pi)


def _FP():
    return np.float32 if os.environ.get("TAICHI_SHIM_FP", "64") == "32" else np.float64


class TiArray(np.ndarray):
    """Vector / matrix value.  Augmented assignment is NOT in place (Taichi value semantics)."""
    __array_priority__ = 100

    x = property(lambda s: s[0], lambda s, v: s.__setitem__(0, v))
    y = property(lambda s: s[1], lambda s, v: s.__setitem__(1, v))
    z = property(lambda s: s[2], lambda s, v: s.__setitem__(2, v))
    w = property(lambda s: s[3], lambda s, v: s.__setitem__(3, v))
    xyz = property(lambda s: np.asarray(s[:3]).copy().view(TiArray))
    xy = property(lambda s: np.asarray(s[:2]).copy().view(TiArray))

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o

    def __eq__(self, o):
        r = np.ndarray.__eq__(self, o)
        return bool(np.all(r)) if isinstance(r, np.ndarray) else r

    def __ne__(self, o):
        return not self.__eq__(o)

    __hash__ = None

    def dot(self, o):
        return np.dot(np.asarray(self), np.asarray(o))

    def transpose(self):
        return np.asarray(self).T.copy().view(TiArray)

    def __matmul__(self, o):
        return np.matmul(np.asarray(self), np.asarray(o)).view(TiArray)


class _VecType:
    """Callable vector type: usable as constructor, annotation and ti.field dtype."""

    def __init__(self, n, dtype=None):
        self.n = n
        self.dtype = dtype  # None -> working float precision

    def __call__(self, *args):
        dt = self.dtype or _FP()
        parts = []
        for a in args:
            a = getattr(a, "v", a)
            parts.append(np.atleast_1d(np.asarray(a, dtype=dt)))
        a = np.concatenate(parts) if parts else np.zeros(self.n, dtype=dt)
        if a.size == 1 and self.n > 1:
            a = np.full(self.n, a[0], dtype=dt)
        if a.shape != (self.n,):
            raise ValueError(f"vec{self.n}: got shape {a.shape}")
        return a.astype(dt).view(TiArray)


vec2 = _VecType(2)
vec3 = _VecType(3)
vec4 = _VecType(4)


def mat3(rows):
    return np.asarray(rows, dtype=_FP()).reshape(3, 3).view(TiArray)


def mat4(rows):
    return np.asarray(rows, dtype=_FP()).reshape(4, 4).view(TiArray)


def eye(n):
    return np.eye(n, dtype=_FP()).view(TiArray)


def dot(a, b):
    return np.dot(np.asarray(a), np.asarray(b))


def cross(a, b):
    return np.cross(np.asarray(a), np.asarray(b)).view(TiArray)


def length(a):
    a = np.asarray(a)
    return np.sqrt(np.dot(a, a))


def normalize(a):
    a = np.asarray(a)
    return (a / np.sqrt(np.dot(a, a))).view(TiArray)


def inverse(m):
    m = np.asarray(m)
    if m.shape == (3, 3):
        # closed-form adjugate / determinant in the working precision (what a 3x3 inverse compiles to)
        a, b, c, d, e, f, g, h, i = m.reshape(-1)
        det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)
        adj = np.array([[e * i - f * h, c * h - b * i, b * f - c * e],
                        [f * g - d * i, a * i - c * g, c * d - a * f],
                        [d * h - e * g, b * g - a * h, a * e - b * d]], dtype=m.dtype)
        return (adj / det).view(TiArray)
    return np.linalg.inv(m).view(TiArray)


def exp(x):
    return np.exp(x)


def sin(x):
    return np.sin(x)


def cos(x):
    return np.cos(x)


def acos(x):
    return np.arccos(x)


def sqrt(x):
    return np.sqrt(x)


def _red(fn, args):
    out = args[0]
    for a in args[1:]:
        out = fn(out, a)
    return out


def max(*args):  # noqa: A001
    return _red(np.maximum, [getattr(a, "v", a) for a in args])


def min(*args):  # noqa: A001
    return _red(np.minimum, [getattr(a, "v", a) for a in args])
