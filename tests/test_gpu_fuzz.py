"""A short replay of the round-2 fuzz sweeps (scratch/fuzz*.py ran ~3000 draws; three of them exposed bugs that are now
regression tests in test_gpu_render.py).  Each family draws adversarial scenes / cameras / options and checks every
render route against the float64 oracle (tolerance of the north star: 1e-3; observed <= 1e-5)."""
import numpy as np
import pytest

from oracle import ref_numpy as O

from gpu_util import make_scene

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _camera(pos, rot, W, H, fx, fy):
    from rtgs.camera import Camera
    return Camera(pos, rot, (W, H), (fx, fy)), O.CameraParams(np.asarray(pos), np.asarray(rot), W, H, (fx, fy))


def _adversarial(seed):
    """World scales 1e-3 .. 1e3, needles and pancakes, sheets, dense cores, near-coincident pairs, cameras grazing or
    inside splats, anisotropic focal lengths."""
    from rtgs.orbit import focal_from_fov, orbit_pose
    rng = np.random.default_rng(31000 + seed)
    n = int(rng.integers(1, 3000))
    S = float(10.0 ** rng.uniform(-3, 3)) if rng.random() < 0.5 else 1.0
    ms = float(10.0 ** rng.uniform(-2.5, -0.5))
    aniso = float(rng.choice([0.5, 1.0, 2.0]))
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    pos = rng.uniform(-1, 1, (n, 3))
    kind = int(rng.integers(0, 4))
    if kind == 1: pos[:, 2] *= 0.02
    if kind == 2: pos *= rng.uniform(0.0, 1.0, (n, 1)) ** 3
    if kind == 3: pos[: n // 2] = pos[n // 2: n // 2 + n // 2] + rng.normal(0, 1e-4, (n // 2, 3))
    scale = np.exp(rng.normal(np.log(ms), aniso, (n, 3)))
    gs = O.GaussianSet(pos=pos * S, rot=q, scale=scale * S, color=1 / (1 + np.exp(-rng.normal(0, 1, (n, 3)))),
                       opacity=1 / (1 + np.exp(-rng.normal(0, 1.5, n))),
                       sh=rng.normal(0, 0.15, (n, 15, 3)) if rng.random() < 0.5 else None)
    W, H = int(rng.integers(9, 150)), int(rng.integers(9, 110))
    depth = int(rng.choice([1, 3, 16, 16, 16, 24]))
    r = float(rng.choice([rng.uniform(0.0, 0.3), rng.uniform(0.3, 4.0)]))
    pos_c, rot_c = orbit_pose(float(rng.uniform(0, 6.28)), float(rng.uniform(0.2, 2.9)), r * S)
    f = focal_from_fov(H, float(rng.uniform(15, 130)))
    fy = f * float(rng.choice([1.0, 1.0, 0.7, 1.4]))
    t_cut = float(10.0 ** rng.uniform(-4, -1))
    return gs, _camera(pos_c, rot_c, W, H, f, fy), depth, t_cut


@pytest.mark.parametrize("seed", [25, 57, 91, 126, 128] + list(range(300, 312)))
def test_adversarial_scenes_all_routes(seed):
    from rtgs.ray_tracer import RayTracer
    gs, (cam, ocam), depth, t_cut = _adversarial(seed)
    scene = make_scene(gs)
    ref = np.asarray(O.render(gs, ocam, depth=depth)["rgb"]).reshape(-1, 3)
    rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
    cut = RayTracer(cam.buf_size, scene, cam, t_cut=t_cut)
    imgs = {}
    for mode in (0, 1, 2):
        if depth > 16 and mode == 2:
            continue
        scene.set_option("render_mode", mode)
        err = float(np.abs(rt.render(depth).reshape(-1, 3) - ref).max())
        assert err <= TOL, (seed, mode, err)
        if mode < 2:
            imgs[mode] = cut.render(min(depth, 16)).copy()
    # with a transmittance cut the routes follow the same rule (stop when T < t_cut): they must agree
    assert np.abs(imgs[0] - imgs[1]).max() <= 1e-5, seed
    # the opt-in depth-slab lists render the same pixels as the default route, bit for bit
    scene.set_option("render_mode", 0)
    base = rt.render(min(depth, 16)).copy()
    scene.set_option("heavy_limit", 24); scene.set_option("heavy_lists", 0)
    forced = rt.render(min(depth, 16)).copy()
    scene.set_option("heavy_lists", 2)
    assert np.array_equal(rt.render(min(depth, 16)), forced), seed
    assert np.abs(forced - base).max() <= 1e-5, seed


@pytest.mark.parametrize("seed", [59] + list(range(500, 512)))
def test_degenerate_images_and_cameras(seed):
    """1x1 and sliver images, 1 .. 170 degree fields of view, camera quaternions that are not unit, duplicated centres,
    one to a few Gaussians."""
    from rtgs.orbit import focal_from_fov, orbit_pose
    from rtgs.ray_tracer import RayTracer
    rng = np.random.default_rng(61000 + seed)
    n = int(rng.choice([1, 2, 3, 7, 33, 300, 3000]))
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    pos = rng.uniform(-1, 1, (n, 3)); scale = np.exp(rng.normal(np.log(10 ** rng.uniform(-2, -0.3)), 0.7, (n, 3)))
    if n > 4 and rng.random() < 0.3:
        pos[n // 2:] = pos[: n - n // 2]
    gs = O.GaussianSet(pos=pos, rot=q, scale=scale, color=rng.uniform(0, 1, (n, 3)), opacity=rng.uniform(0.01, 0.99, n),
                       sh=rng.normal(0, 0.15, (n, 15, 3)) if rng.random() < 0.5 else None)
    shape = int(rng.integers(0, 4))
    W, H = [(1, 1), (int(rng.integers(1, 4)), int(rng.integers(50, 300))), (int(rng.integers(50, 400)), int(rng.integers(1, 4))),
            (int(rng.integers(100, 700)), int(rng.integers(100, 400)))][shape]
    pos_c, rot_c = orbit_pose(float(rng.uniform(0, 6.28)), float(rng.uniform(0.05, 3.09)), float(10 ** rng.uniform(-1, 0.7)))
    rot_c = np.asarray(rot_c, np.float64) * (float(rng.uniform(0.5, 2.0)) if rng.random() < 0.3 else 1.0)
    fov = float(rng.choice([rng.uniform(1, 10), rng.uniform(10, 120), rng.uniform(120, 170)]))
    f = focal_from_fov(max(H, 2), fov)
    fy = f * float(rng.choice([1.0, 0.5, 2.0]))
    cam, ocam = _camera(pos_c, rot_c, W, H, f, fy)
    depth = int(rng.choice([1, 16, 16, 32]))
    scene = make_scene(gs)
    ref = np.asarray(O.render(gs, ocam, depth=depth)["rgb"]).reshape(-1, 3)
    rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
    for mode in (0, 2, 1):
        scene.set_option("render_mode", mode)
        err = float(np.abs(rt.render(depth).reshape(-1, 3) - ref).max())
        assert err <= TOL, (seed, mode, err)
    scene.set_option("render_mode", 0)
