"""Axis-aligned bound — host mirror of the reference's ``rtgs/bounding_box.py``."""
from __future__ import annotations

import numpy as np

from .utils.types import inf, vec2, vec3


class Bound:
    """Bounding box (bounding_box.py:5-13): p_min, p_max."""

    __slots__ = ("p_min", "p_max")

    def __init__(self, p_min=None, p_max=None):
        self.p_min = vec3(0) if p_min is None else vec3(p_min)
        self.p_max = vec3(0) if p_max is None else vec3(p_max)

    def init(self, p_min=vec3(inf), p_max=vec3(-inf)):
        """bounding_box.py:15-19 — empty box (+inf, -inf)."""
        self.p_min = vec3(p_min)
        self.p_max = vec3(p_max)

    def get_centroid(self):
        return vec3(0.5 * (np.asarray(self.p_min) + np.asarray(self.p_max)))

    def size(self):
        return vec3(np.asarray(self.p_max) - np.asarray(self.p_min))

    def area(self):
        """bounding_box.py:29-34."""
        s = np.asarray(self.size(), dtype=np.float64)
        return float(2 * (s[0] * s[1] + s[1] * s[2] + s[2] * s[0]))

    area_py = area

    def union(self, box: "Bound") -> "Bound":
        """bounding_box.py:42-48."""
        return Bound(np.minimum(self.p_min, box.p_min), np.maximum(self.p_max, box.p_max))

    def hit(self, ray) -> vec2:
        """Slab test (bounding_box.py:50-89): returns (t_min, t_max); hit iff t_min < t_max.
        No clamp to [start, end]; division by a zero direction component is unguarded, as in
        the reference."""
        o = np.asarray(ray.origin, dtype=np.float32)
        d = np.asarray(ray.direction, dtype=np.float32)
        neg = d < 0
        near = np.where(neg, self.p_max, self.p_min).astype(np.float32)
        far = np.where(neg, self.p_min, self.p_max).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            t0 = (near - o) / d
            t1 = (far - o) / d
        return vec2(np.max(t0), np.min(t1))

    def __repr__(self):
        return f"Bound(p_min={self.p_min.to_list()}, p_max={self.p_max.to_list()})"
