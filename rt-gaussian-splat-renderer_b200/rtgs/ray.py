"""Ray record — host mirror of the reference's ``rtgs/ray.py``.

The fused render kernel generates rays in registers and never stores them (csrc/render.cu);
this class exists so code written against the reference's ``Ray`` / ``new_ray`` / ``Ray.field``
keeps working (ray.py:4-68) and to feed ``Scene.hit`` with explicit rays.
"""
from __future__ import annotations

import numpy as np

from .utils.types import inf, vec3

#: layout of one ray in the batched arrays passed to the C-ABI: origin, direction, start, end
RAY_DTYPE = np.dtype([("origin", np.float32, 3), ("direction", np.float32, 3),
                      ("start", np.float32), ("end", np.float32)])


class Ray:
    """A ray (ray.py:4-18): origin, direction, start (min t), end (max t)."""

    __slots__ = ("origin", "direction", "start", "end")

    def __init__(self, origin=None, direction=None, start=0.0, end=0.0):
        # Like a zero-initialised Taichi struct when constructed without arguments.
        self.origin = vec3(0) if origin is None else vec3(origin)
        self.direction = vec3(0) if direction is None else vec3(direction)
        self.start = float(start)
        self.end = float(end)

    def init(self, origin=vec3(0, 0, 0), direction=vec3(0, 1, 0), start=0, end=inf):
        """ray.py:20-41 — defaults: origin 0, direction +y, start 0, end inf."""
        self.origin = vec3(origin)
        self.direction = vec3(direction)
        self.start = float(start)
        self.end = float(end)

    def get(self, t):
        """ray.py:43-52 — origin + t * direction (float32 like the reference)."""
        return vec3(np.asarray(self.origin) + np.float32(t) * np.asarray(self.direction))

    def to_record(self):
        r = np.zeros((), dtype=RAY_DTYPE)
        r["origin"], r["direction"], r["start"], r["end"] = self.origin, self.direction, self.start, self.end
        return r

    @staticmethod
    def field(shape):
        """``Ray.field(shape)`` — zero-initialised structured array (ray.py / camera.py:29)."""
        from .fields import StructArrayField
        return StructArrayField(np.zeros(shape, dtype=RAY_DTYPE), Ray._from_record)

    @staticmethod
    def _from_record(rec):
        return Ray(rec["origin"], rec["direction"], rec["start"], rec["end"])

    def __repr__(self):
        return f"Ray(origin={self.origin.to_list()}, direction={self.direction.to_list()}, start={self.start}, end={self.end})"


def new_ray(origin=vec3(0, 0, 0), direction=vec3(0, 1, 0), start=0, end=inf) -> Ray:
    """Python-scope constructor (ray.py:55-68)."""
    return Ray(origin, direction, start, end)
