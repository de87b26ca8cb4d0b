"""Quaternion utilities — host (NumPy) mirror of the reference's ``rtgs/utils/quaternion.py``.

Scalar-last (x, y, z, w) Hamilton algebra with the same function names and semantics
(utils/quaternion.py:8-147).  The device kernels carry their own float64 implementation
(csrc/gsmath.cuh); these are the host-callable helpers the reference exposes to users/tests.
Arithmetic is float64 internally; results are returned as float32 ``vec3`` / ``vec4`` /
ndarray like the reference's f32 Taichi values.
"""
from __future__ import annotations

import numpy as np

from .types import vec3, vec4


def _q(q):
    return np.asarray(q, dtype=np.float64).reshape(4)


def _v(v):
    return np.asarray(v, dtype=np.float64).reshape(3)


def _mul64(p, q):
    pv, pw = p[:3], p[3]
    qv, qw = q[:3], q[3]
    w = pw * qw - np.dot(pv, qv)                       # utils/quaternion.py:19
    v = pw * qv + qw * pv + np.cross(pv, qv)           # utils/quaternion.py:21
    return np.array([v[0], v[1], v[2], w])


def _conj64(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def mul(p, q) -> vec4:
    """Quaternion multiplication pq (utils/quaternion.py:8-23)."""
    return vec4(_mul64(_q(p), _q(q)))


def conj(q) -> vec4:
    """Complex conjugate (utils/quaternion.py:26-35)."""
    return vec4(_conj64(_q(q)))


def inv(q) -> vec4:
    """Inverse as the reference defines it: conj(q) / |q| (utils/quaternion.py:38-47 — note it
    divides by the length, not the squared length; identical for unit quaternions)."""
    q = _q(q)
    return vec4(_conj64(q) / np.linalg.norm(q))


def from_axis_angle(v) -> vec4:
    """Axis-angle vector (direction = axis, length = angle) -> quaternion
    (utils/quaternion.py:50-64)."""
    v = _v(v)
    theta = np.linalg.norm(v)
    if theta > 0:
        v = v / theta * np.sin(theta / 2)
    w = np.cos(theta / 2)
    return vec4(v[0], v[1], v[2], w)


def as_axis_angle(q) -> vec3:
    """Unit quaternion -> axis-angle vector (utils/quaternion.py:67-81)."""
    q = _q(q)
    theta = np.arccos(np.clip(q[3], -1.0, 1.0)) * 2
    norm = np.linalg.norm(q[:3])
    res = np.zeros(3)
    if norm > 0:
        res = q[:3] / norm * theta
    return vec3(res)


def _rot64(q, v):
    qv = np.array([v[0], v[1], v[2], 0.0])
    return _mul64(q, _mul64(qv, _conj64(q)))[:3]       # utils/quaternion.py:95-96


def rot_vec3(q, v) -> vec3:
    """Rotate v by q as q v q* (utils/quaternion.py:84-96); q is used as given."""
    return vec3(_rot64(_q(q), _v(v)))


def as_rotation_mat3(q) -> np.ndarray:
    """3x3 rotation matrix whose columns are the rotated basis vectors
    (utils/quaternion.py:99-121)."""
    q = _q(q)
    m = np.eye(3)
    m[:, 0] = _rot64(q, np.array([1.0, 0, 0]))
    m[:, 1] = _rot64(q, np.array([0, 1.0, 0]))
    m[:, 2] = _rot64(q, np.array([0, 0, 1.0]))
    return m.astype(np.float32)


def as_rotation_mat4(q) -> np.ndarray:
    """4x4 homogeneous rotation matrix (utils/quaternion.py:124-147)."""
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = as_rotation_mat3(q)
    return m


def from_rotation_matrix(m) -> vec4:
    """Proper rotation matrix -> unit quaternion (x,y,z,w), w >= 0.  Additive helper standing in
    for numpy-quaternion's ``from_rotation_matrix`` that the reference's viewer uses
    (__main__.py:134)."""
    m = np.asarray(m, dtype=np.float64)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s]
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s]
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s, (m[1, 0] - m[0, 1]) / s]
    q = np.asarray(q)
    q = q / np.linalg.norm(q)
    if q[3] < 0:
        q = -q
    return vec4(q)
