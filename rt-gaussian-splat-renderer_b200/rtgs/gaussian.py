"""Gaussian record — host mirror of the reference's ``rtgs/gaussian.py``.

The device kernels work on packed SoA records (csrc/common.cuh).  This module keeps the
reference's ``Gaussian`` / ``new_gaussian`` surface (gaussian.py:26-84, :233-247) and its
per-Gaussian maths as host-callable float64 helpers with the reference's method names, so code
and tests written against the reference port directly.
"""
from __future__ import annotations

from math import sqrt

import numpy as np

from .bounding_box import Bound
from .utils import quaternion as quat
from .utils.types import inf, vec2, vec3, vec4

BOUNDING_THRESHOLD = 3  # gaussian.py:13

# Spherical-harmonics constants (gaussian.py:17-23)
c_0 = sqrt(3 / np.pi)
c_1 = sqrt(15 / np.pi)
c_2 = sqrt(5 / np.pi)
c_3 = sqrt(35 / (2 * np.pi))
c_4 = sqrt(105 / np.pi)
c_5 = sqrt(21 / (2 * np.pi))
c_6 = sqrt(7 / np.pi)

SH_NAMES = ("sh_10", "sh_11", "sh_12", "sh_20", "sh_21", "sh_22", "sh_23", "sh_24",
            "sh_30", "sh_31", "sh_32", "sh_33", "sh_34", "sh_35", "sh_36")

GAUSSIAN_DTYPE = np.dtype([("position", np.float32, 3), ("rotation", np.float32, 4),
                           ("scale", np.float32, 3), ("color", np.float32, 3), ("opacity", np.float32)]
                          + [(n, np.float32, 3) for n in SH_NAMES])   # 59 f32 = 236 B (gaussian.py:26-55)


class Gaussian:
    """Gaussian (gaussian.py:26-55).  A no-argument construction is zero-initialised like a
    Taichi struct; ``init()`` applies the reference's defaults."""

    __slots__ = ("position", "rotation", "scale", "color", "opacity") + SH_NAMES

    def __init__(self, position=None, rotation=None, scale=None, color=None, opacity=0.0):
        self.position = vec3(0) if position is None else vec3(position)
        self.rotation = vec4(0) if rotation is None else vec4(rotation)
        self.scale = vec3(0) if scale is None else vec3(scale)
        self.color = vec3(0) if color is None else vec3(color)
        self.opacity = float(opacity)
        for n in SH_NAMES:
            setattr(self, n, vec3(0))

    def init(self, position=vec3(0, 0, 0), rotation=vec4(0, 0, 0, 1), scale=vec3(1, 1, 1),
             color=vec3(1, 0, 1), opacity=1):
        """gaussian.py:59-84."""
        self.position, self.rotation = vec3(position), vec4(rotation)
        self.scale, self.color, self.opacity = vec3(scale), vec3(color), float(opacity)

    # ---- maths (float64 on the float32 stored values) ------------------------------------
    def cov(self) -> np.ndarray:
        """gaussian.py:86-102 — R S S^T R^T."""
        R = quat.as_rotation_mat3(self.rotation).astype(np.float64)
        q = np.asarray(self.rotation, dtype=np.float64)
        R = _rotmat64(q)
        S = np.diag(np.asarray(self.scale, dtype=np.float64))
        return R @ S @ S.T @ R.T

    def bounding_box(self) -> Bound:
        """gaussian.py:104-138 — AABB of the six points p +- R S (3 e_k)."""
        R = _rotmat64(np.asarray(self.rotation, dtype=np.float64))
        s = np.asarray(self.scale, dtype=np.float64)
        p = np.asarray(self.position, dtype=np.float64)
        pts = []
        for k in range(3):
            e = np.zeros(3)
            e[k] = BOUNDING_THRESHOLD
            off = R @ (s * e)
            pts += [p + off, p - off]
        pts = np.array(pts)
        return Bound(pts.min(axis=0), pts.max(axis=0))

    def eval_sh(self, dir) -> vec3:
        """gaussian.py:140-181 — including the ``5z^2 - 3z`` term exactly as coded (:160)."""
        x, y, z = (float(v) for v in np.asarray(dir, dtype=np.float64))
        ys = sh_basis(x, y, z)
        out = np.zeros(3)
        for yk, n in zip(ys, SH_NAMES):
            out += yk * np.asarray(getattr(self, n), dtype=np.float64)
        return vec3(out)

    def eval(self, pos, dir) -> vec4:
        """gaussian.py:183-201 — (rgb, alpha) at ``pos`` seen along ``dir``."""
        d = np.asarray(pos, dtype=np.float64) - np.asarray(self.position, dtype=np.float64)
        cov_inv = np.linalg.inv(self.cov())
        rho = np.exp(-d @ (cov_inv @ d))
        alpha = self.opacity * rho
        dn = np.asarray(dir, dtype=np.float64)
        dn = dn / np.linalg.norm(dn)
        color = np.asarray(self.color, dtype=np.float64) + np.asarray(self.eval_sh(dn), dtype=np.float64)
        return vec4(color[0], color[1], color[2], alpha)

    def hit(self, ray) -> vec2:
        """gaussian.py:203-230 — the two roots of the sqrt(3)-sigma quadratic, (inf, inf) on a miss."""
        cov_inv = np.linalg.inv(self.cov())
        d = np.asarray(ray.direction, dtype=np.float64)
        v = np.asarray(ray.origin, dtype=np.float64) - np.asarray(self.position, dtype=np.float64)
        A = d @ (cov_inv @ d)
        B = 2 * d @ (cov_inv @ v)
        C = v @ (cov_inv @ v) - BOUNDING_THRESHOLD
        delta = B ** 2 - 4 * A * C
        if delta > 0:
            return vec2((-B - np.sqrt(delta)) / (2 * A), (-B + np.sqrt(delta)) / (2 * A))
        if delta == 0:
            return vec2(-B / (2 * A), inf)
        return vec2(inf, inf)

    # ---- records -------------------------------------------------------------------------
    def to_record(self):
        r = np.zeros((), dtype=GAUSSIAN_DTYPE)
        for n in GAUSSIAN_DTYPE.names:
            r[n] = getattr(self, n)
        return r

    @staticmethod
    def _from_record(rec):
        g = Gaussian(rec["position"], rec["rotation"], rec["scale"], rec["color"], rec["opacity"])
        for n in SH_NAMES:
            setattr(g, n, vec3(rec[n]))
        return g

    @staticmethod
    def field(shape):
        """``Gaussian.field(shape)`` — zero-initialised structured array (gaussian.py / scene.py:131)."""
        from .fields import StructArrayField
        return StructArrayField(np.zeros(shape, dtype=GAUSSIAN_DTYPE), Gaussian._from_record)

    def __repr__(self):
        return (f"Gaussian(position={self.position.to_list()}, rotation={self.rotation.to_list()}, "
                f"scale={self.scale.to_list()}, color={self.color.to_list()}, opacity={self.opacity})")


def _rotmat64(q):
    m = np.eye(3)
    for k in range(3):
        e = np.zeros(3)
        e[k] = 1.0
        m[:, k] = quat._rot64(q, e)
    return m


def sh_basis(x, y, z):
    """The 15 basis values of gaussian.py:149-163 for a normalised direction."""
    return [
        0.5 * c_0 * y, 0.5 * c_0 * z, 0.5 * c_0 * x,
        0.5 * c_1 * x * y, 0.5 * c_1 * y * z, 0.25 * c_2 * (3 * z ** 2 - 1), 0.5 * c_1 * x * z,
        0.25 * c_1 * (x ** 2 - y ** 2),
        0.25 * c_3 * y * (3 * x ** 2 - y ** 2), 0.5 * c_4 * x * y * z, 0.25 * c_5 * y * (5 * z ** 2 - 1),
        0.25 * c_6 * (5 * z ** 2 - 3 * z), 0.25 * c_5 * x * (5 * z ** 2 - 1),
        0.25 * c_4 * (x ** 2 - y ** 2) * z, 0.25 * c_3 * x * (x ** 2 - 3 * y ** 2),
    ]


def new_gaussian(position=vec3(0, 0, 0), rotation=vec4(0, 0, 0, 1), scale=vec3(1, 1, 1),
                 color=vec3(1, 0, 1), opacity=1) -> Gaussian:
    """Python-scope constructor (gaussian.py:233-247)."""
    return Gaussian(position, rotation, scale, color, opacity)
