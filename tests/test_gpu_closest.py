"""Scene.hit (closest entry, scene.py:406-450) and Camera.generate_ray_field (camera.py:57-71)
against the oracle."""
import numpy as np
import pytest

from oracle import ref_numpy as O

from gpu_util import make_camera, make_scene, random_set

pytestmark = pytest.mark.gpu


def test_generate_ray_field():
    cam, ocam = make_camera(0.7, 1.2, 2.2, 96, 64)
    f = cam.cam_ray_field.to_numpy()
    o, d = O.camera_rays(ocam)
    assert f.shape == (96, 64, 8)
    assert np.allclose(f[..., :3], o, atol=1e-7)
    assert np.abs(f[..., 3:6].reshape(-1, 3) - d).max() <= 1e-7
    assert (f[..., 6] == 0).all() and np.isinf(f[..., 7]).all()


def test_closest_hit_camera_rays():
    gs = random_set(5000, seed=3, mean_scale=0.03, sh=False)
    scene = make_scene(gs)
    cam, ocam = make_camera(0.3, 1.0, 2.5, 96, 64)
    rays = cam.cam_ray_field.to_numpy().reshape(-1, 8)
    hit = scene.hit(rays)
    idx, t12 = O.closest_hit(gs, rays[:, :3].astype(np.float64), rays[:, 3:6].astype(np.float64))
    assert np.array_equal(hit.gaussian_idx, idx)
    m = idx >= 0
    assert m.any() and np.allclose(hit.intersections[m], t12[m], rtol=1e-6)
    assert np.isinf(hit.intersections[~m]).all() and (hit.depth[~m] == -1).all() and (hit.depth[m] > 0).all()


def test_closest_hit_arbitrary_rays_with_interval():
    rng = np.random.default_rng(8)
    gs = random_set(3000, seed=4, mean_scale=0.05, sh=False)
    scene = make_scene(gs)
    n = 4000
    o = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[:50, 0] = 0.0                                    # axis-parallel components (unguarded division)
    start = rng.uniform(0, 1.0, n).astype(np.float32)
    end = (start + rng.uniform(0.2, 3.0, n)).astype(np.float32)
    rays = np.concatenate([o, d, start[:, None], end[:, None]], axis=1)
    hit = scene.hit(rays)
    idx, t12 = O.closest_hit(gs, o.astype(np.float64), d.astype(np.float64), start.astype(np.float64),
                             end.astype(np.float64))
    assert np.array_equal(hit.gaussian_idx, idx)
    m = idx >= 0
    assert m.sum() > 100 and np.allclose(hit.intersections[m], t12[m], rtol=1e-6, atol=1e-7)


def test_single_ray_object():
    from rtgs.ray import new_ray
    from rtgs.utils.types import vec3
    gs = O.GaussianSet(pos=[[0, 0, 0], [0, 3, 0]], rot=[[0, 0, 0, 1]] * 2, scale=[[1, 1, 1]] * 2,
                       color=[[1, 0, 1]] * 2, opacity=[1, 1])
    scene = make_scene(gs)
    h = scene.hit(new_ray(vec3(0, -5, 0), vec3(0, 1, 0)))
    assert h.gaussian_idx[0] == 0 and np.allclose(h.intersections[0], [5 - np.sqrt(3), 5 + np.sqrt(3)], atol=1e-5)
    h = scene.hit(new_ray(vec3(0, -5, 0), vec3(0, 1, 0), start=4.0))
    assert h.gaussian_idx[0] == 1
