"""Synthetic random-Gaussian scenes for the benchmark configurations.

``numpy.random.default_rng(seed)`` draws, in this order, all cast to float32:
pos ~ U[-1,1)^3; quat ~ N(0,1)^4 normalised (x,y,z,w); log_scale ~ N(ln s, 0.5^2) per axis;
opacity_logit ~ N(0, 1.5^2); f_dc ~ N(0,1); f_rest ~ N(0, 0.15^2) shape (N,15,3) (SH degree 3) or none
(degree 0).  Activations exactly as the reference's loader (scene.py:110-114).
s = 2 sqrt(H/(3 pi N)) with H = 16 expected ellipsoid crossings per cube-spanning ray.
"""
from __future__ import annotations

import numpy as np

from .utils.math import sigmoid

CONFIGS = {
    # name: (N, seed, sh_degree, (W, H))
    "100k_deg0_1080p": (100_000, 1001, 0, (1920, 1080)),
    "1m_deg3_1080p": (1_000_000, 1002, 3, (1920, 1080)),
    "1m_deg0_1080p": (1_000_000, 1002, 0, (1920, 1080)),   # same geometry without SH (diagnostic)
    "3m_deg3_2160p": (3_000_000, 1003, 3, (3840, 2160)),
}
# Scenes shaped like a trained capture rather than a uniform cloud (the reference's own target
# workload is truck.ply, docs/source/get-started.md:64-74): see make_surface_scene.
SURFACE_CONFIGS = {
    # name: (N, seed, (W, H), orbit views, orbit phi)
    "surface_1m_1080p": (1_000_000, 4242, (1920, 1080), 16, float(np.pi / 2 - 0.3)),
}
ORBIT_R = 2.2
FOV_DEG = 60.0
H_TARGET = 16.0


def mean_scale(n: int, h_target: float = H_TARGET) -> float:
    return 2.0 * float(np.sqrt(h_target / (3.0 * np.pi * n)))


def make_scene(n: int, seed: int, sh_degree: int = 3, h_target: float = H_TARGET) -> dict:
    """Post-activation parameter arrays: pos, rot, scale, color, opacity, sh (or None)."""
    f32 = np.float32
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-1.0, 1.0, size=(n, 3)).astype(f32)
    quat = rng.normal(0.0, 1.0, size=(n, 4)).astype(f32)
    log_scale = rng.normal(np.log(mean_scale(n, h_target)), 0.5, size=(n, 3)).astype(f32)
    opacity_logit = rng.normal(0.0, 1.5, size=(n,)).astype(f32)
    f_dc = rng.normal(0.0, 1.0, size=(n, 3)).astype(f32)
    sh = rng.normal(0.0, 0.15, size=(n, 15, 3)).astype(f32) if sh_degree > 0 else None
    rot = quat / np.linalg.norm(quat, axis=-1)[:, np.newaxis]     # scene.py:110-111
    scale = np.exp(log_scale) * f32(1.0)                           # scene.py:112
    color = sigmoid(f_dc)                                          # scene.py:113
    opacity = sigmoid(opacity_logit)                               # scene.py:114
    return dict(pos=pos, rot=rot.astype(f32), scale=scale.astype(f32), color=color.astype(f32),
                opacity=opacity.astype(f32), sh=sh)


def export_ply(path, arrays: dict) -> None:
    """Write post-activation arrays (``make_scene``) as a 62-property 3DGS ``.ply`` in the layout of
    tests/data/test.ply, undoing the loader's activations (scene.py:101-114: log scale, logit colour / opacity,
    scalar-first quaternion, channel-major ``f_rest``), so that the reference itself - or ``Scene.load_file`` here -
    can consume the bench scenes."""
    from .ply import write_gs_ply
    f64 = np.float64
    logit = lambda p: np.log(np.asarray(p, f64)) - np.log1p(-np.asarray(p, f64))
    pos, rot = np.asarray(arrays["pos"]), np.asarray(arrays["rot"])
    cols = {"x": pos[:, 0], "y": pos[:, 1], "z": pos[:, 2],
            "rot_0": rot[:, 3], "rot_1": rot[:, 0], "rot_2": rot[:, 1], "rot_3": rot[:, 2],
            "opacity": logit(arrays["opacity"])}
    ls = np.log(np.asarray(arrays["scale"], f64))
    dc = logit(arrays["color"])
    for k in range(3):
        cols[f"scale_{k}"] = ls[:, k]
        cols[f"f_dc_{k}"] = dc[:, k]
    if arrays.get("sh") is not None:
        rest = np.asarray(arrays["sh"]).reshape(-1, 15, 3).transpose(0, 2, 1).reshape(-1, 45)   # f_rest_{15c+k} = sh[k][c]
        for i in range(45):
            cols[f"f_rest_{i}"] = rest[:, i]
    write_gs_ply(path, cols)


def make_surface_scene(n: int, seed: int = 4242) -> dict:
    """Gaussians on SURFACES - six spherical shells and a floor - flat (one axis 10x thinner), with heavy-tailed sizes
    (log-normal, sigma 1), forty huge translucent blobs and six stray points 800 scene radii out: the features of a
    trained 3DGS capture that a uniform cloud lacks (tiles that look along the floor hold thousands of splats; the
    stray points defeat 10-bit-per-axis Morton codes).  SH degree 3.  Post-activation arrays like ``make_scene``."""
    f32 = np.float32
    rng = np.random.default_rng(seed)
    pos = np.empty((n, 3), f32)
    k = rng.integers(0, 7, n)
    c = rng.uniform(-0.6, 0.6, (6, 3))
    r = rng.uniform(0.15, 0.45, 6)
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    shell = k < 6
    pos[shell] = (c[k[shell]] + r[k[shell], None] * u[shell] * (1 + 0.01 * rng.normal(size=(shell.sum(), 1)))).astype(f32)
    pos[~shell] = np.stack([rng.uniform(-1, 1, (~shell).sum()), rng.uniform(-1, 1, (~shell).sum()),
                            -0.8 + 0.005 * rng.normal(size=(~shell).sum())], axis=1).astype(f32)
    quat = rng.normal(size=(n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    base = 2.0 * np.sqrt(16.0 / (3 * np.pi * n)) * 2.0
    ls = rng.normal(np.log(base), 1.0, (n, 3))            # heavy tail (sigma 1.0 instead of 0.5)
    ls[:, 2] -= np.log(10.0)                               # flat splats
    scale = np.exp(ls).astype(f32)
    big = rng.choice(n, 40, replace=False)
    scale[big] = rng.uniform(0.2, 0.6, (40, 3)).astype(f32)   # huge blobs
    stray = rng.choice(n, 6, replace=False)
    pos[stray] = (rng.uniform(-1, 1, (6, 3)) * 800.0).astype(f32)
    opacity = (1 / (1 + np.exp(-rng.normal(0.5, 2.0, n)))).astype(f32)
    opacity[big] = 0.05
    color = (1 / (1 + np.exp(-rng.normal(0, 1, (n, 3))))).astype(f32)
    sh = rng.normal(0, 0.15, (n, 15, 3)).astype(f32)
    return dict(pos=pos, rot=quat.astype(f32), scale=scale, color=color, opacity=opacity, sh=sh)
