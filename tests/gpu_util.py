"""Helpers shared by the GPU parity tests."""
import numpy as np

from oracle import ref_numpy as O


def random_set(n, seed, mean_scale=0.05, sh=True, box=1.0):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return O.GaussianSet(pos=rng.uniform(-box, box, (n, 3)), rot=q,
                         scale=np.exp(rng.normal(np.log(mean_scale), 0.5, (n, 3))),
                         color=1 / (1 + np.exp(-rng.normal(0, 1, (n, 3)))),
                         opacity=1 / (1 + np.exp(-rng.normal(0, 1.5, n))),
                         sh=rng.normal(0, 0.15, (n, 15, 3)) if sh else None)


def make_scene(gs, device=None, **kw):
    from rtgs.scene import Scene
    return Scene(device=device, **kw).from_arrays(gs.pos, gs.rot, gs.scale, gs.color, gs.opacity, gs.sh)


def make_camera(theta, phi, r, W, H, fov=60.0):
    from rtgs.camera import Camera
    from rtgs.orbit import focal_from_fov, orbit_pose
    pos, rot = orbit_pose(theta, phi, r)
    f = focal_from_fov(H, fov)
    cam = Camera(pos, rot, (W, H), (f, f))
    ocam = O.CameraParams(np.asarray(pos), np.asarray(rot), W, H, (f, f))
    return cam, ocam


def compare(img, ref, tol=1e-3):
    """(max-abs, PSNR dB, number of pixels above tol)."""
    d = np.abs(np.asarray(img, np.float64) - ref)
    return float(d.max()), O.psnr(img, ref), int((d.max(axis=-1) > tol).sum())
