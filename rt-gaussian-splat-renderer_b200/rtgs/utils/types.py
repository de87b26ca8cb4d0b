"""Small fixed-size vector types — Taichi-free stand-ins for ``ti.math.vec*`` and the
reference's ``rtgs.utils.types.vec2i / vec3i`` (utils/types.py:3-4).

They are NumPy arrays with ``.x/.y/.z/.w`` accessors (and the ``.xyz`` swizzle the reference's
quaternion code uses), constructible from scalars, one broadcast scalar, or any sequence:
``vec3(0)``, ``vec3(1, 2, 3)``, ``vec2i((960, 540))``.
"""
from __future__ import annotations

import numpy as np


class _Vec(np.ndarray):
    _n = 0
    _dtype = np.float32

    def __new__(cls, *args):
        if len(args) == 1:
            a = np.asarray(args[0], dtype=cls._dtype)
            if a.ndim == 0:
                a = np.full(cls._n, a, dtype=cls._dtype)
        elif len(args) == 0:
            a = np.zeros(cls._n, dtype=cls._dtype)
        else:
            a = np.concatenate([np.atleast_1d(np.asarray(v, dtype=cls._dtype)) for v in args])
        if a.shape != (cls._n,):
            raise ValueError(f"{cls.__name__} needs {cls._n} components, got shape {a.shape}")
        return np.array(a, dtype=cls._dtype).view(cls)

    def __array_finalize__(self, obj):
        pass

    x = property(lambda s: s[0].item(), lambda s, v: s.__setitem__(0, v))
    y = property(lambda s: s[1].item(), lambda s, v: s.__setitem__(1, v))
    z = property(lambda s: s[2].item(), lambda s, v: s.__setitem__(2, v))
    w = property(lambda s: s[3].item(), lambda s, v: s.__setitem__(3, v))

    @property
    def xyz(self):
        return np.asarray(self[:3]).view(vec3)

    @property
    def xy(self):
        return np.asarray(self[:2]).view(vec2)

    def __eq__(self, other):  # ti vectors compare by value in the reference's tests
        try:
            return bool(np.array_equal(np.asarray(self), np.asarray(other, dtype=self.dtype)))
        except (TypeError, ValueError):
            return False

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def dot(self, other):
        return float(np.dot(np.asarray(self, dtype=np.float64), np.asarray(other, dtype=np.float64)))

    def to_list(self):
        return np.asarray(self).tolist()


class vec2(_Vec):
    _n = 2


class vec3(_Vec):
    _n = 3


class vec4(_Vec):
    _n = 4


class vec2i(_Vec):
    _n = 2
    _dtype = np.int32


class vec3i(_Vec):
    _n = 3
    _dtype = np.int32


inf = float("inf")
