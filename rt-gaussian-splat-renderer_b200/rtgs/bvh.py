"""BVH node record — host mirror of the reference's ``rtgs/bvh.py`` (bvh.py:10-44).

The device traversal uses packed 64-byte two-child nodes (csrc/common.cuh); ``Scene.bvh_field``
exposes the LBVH in the reference's record shape: bound, left, right (-1 = none), prim_left,
prim_right (half-open range into the Morton-sorted Gaussian order), depth.
"""
from __future__ import annotations

import numpy as np

from .bounding_box import Bound

BVH_DTYPE = np.dtype([("p_min", np.float32, 3), ("p_max", np.float32, 3), ("left", np.int32),
                      ("right", np.int32), ("prim_left", np.int32), ("prim_right", np.int32),
                      ("depth", np.int32)])


class BVHNode:
    __slots__ = ("bound", "left", "right", "prim_left", "prim_right", "depth")

    def __init__(self, bound=None, left=0, right=0, prim_left=0, prim_right=0, depth=0):
        self.bound = Bound() if bound is None else bound
        self.left, self.right = int(left), int(right)
        self.prim_left, self.prim_right = int(prim_left), int(prim_right)
        self.depth = int(depth)

    def init(self, bound=None, left=-1, right=-1, prim_left=-1, prim_right=-1, depth=-1):
        """bvh.py:19-33 — everything -1, empty bound."""
        if bound is None:
            bound = Bound()
            bound.init()
        self.bound = bound
        self.left, self.right = left, right
        self.prim_left, self.prim_right = prim_left, prim_right
        self.depth = depth

    def hit(self, ray):
        """bvh.py:35-44."""
        return self.bound.hit(ray)

    @staticmethod
    def _from_record(rec):
        return BVHNode(Bound(rec["p_min"], rec["p_max"]), rec["left"], rec["right"], rec["prim_left"],
                       rec["prim_right"], rec["depth"])

    @staticmethod
    def field(shape):
        from .fields import StructArrayField
        a = np.zeros(shape, dtype=BVH_DTYPE)
        return StructArrayField(a, BVHNode._from_record)
