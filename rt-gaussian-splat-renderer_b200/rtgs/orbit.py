"""Orbit camera pose and focal length — the headless part of the reference's viewer
(``rtgs/__main__.py:88-92`` focal from vertical FOV, ``:120-142`` orbit pose), used by the benchmark
driver and the headless CLI.  Host-side NumPy float64, results cast to float32 like the reference's
``ti.math.vec3 / vec4`` camera attributes."""
from __future__ import annotations

import numpy as np

from .utils.quaternion import from_rotation_matrix
from .utils.types import vec3, vec4


def focal_from_fov(height: int, fov_deg: float) -> float:
    """__main__.py:91-92."""
    half_angle = (fov_deg * np.pi) / 360
    return float((height / 2) / np.tan(half_angle))


def orbit_pose(theta: float, phi: float, r: float, cursor=(0.0, 0.0, 0.0)):
    """__main__.py:120-142 with the identity global rotation: returns (position vec3, rotation vec4
    (x,y,z,w)).  Camera columns: right, up, -look."""
    pos = np.array([r * np.cos(theta) * np.sin(phi), r * np.sin(theta) * np.sin(phi), r * np.cos(phi)])
    look = -pos / np.linalg.norm(pos)
    cam_right = np.array([-np.sin(theta), np.cos(theta), 0.0])
    cam_up = np.cross(cam_right, look)
    rot = np.array([cam_right, cam_up, -look]).T
    q = from_rotation_matrix(rot)
    return vec3(pos + np.asarray(cursor, dtype=np.float64)), vec4(q)
