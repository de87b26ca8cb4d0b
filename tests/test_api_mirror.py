"""Ports of the reference's host-side tests (tests/test_ray.py, tests/test_gaussian.py) against the
Taichi-free API mirror: same constructor defaults, field zero-initialisation and Ray.get."""
import numpy as np

from rtgs.bounding_box import Bound
from rtgs.bvh import BVHNode
from rtgs.gaussian import Gaussian, new_gaussian
from rtgs.ray import Ray, new_ray
from rtgs.utils.math import sigmoid
from rtgs.utils.types import inf, vec2i, vec3, vec3i, vec4


def test_new_ray_defaults():
    # tests/test_ray.py:15-28
    ray = new_ray()
    assert ray.origin == vec3(0) and ray.direction == vec3(0, 1, 0)
    assert ray.start == 0 and ray.end == inf
    ray = new_ray(vec3(1, 2, 3), vec3(0, 0, 1), 1, 10)
    assert ray.origin == vec3(1, 2, 3) and ray.direction == vec3(0, 0, 1) and ray.start == 1 and ray.end == 10


def test_ray_field_zero_init_then_init():
    # tests/test_ray.py:31-50
    f = Ray.field(shape=(4, 4))
    assert f.shape == (4, 4)
    assert f[0, 0].origin == vec3(0) and f[0, 0].direction == vec3(0) and f[0, 0].start == 0 and f[0, 0].end == 0
    r = f[0, 0]
    r.init()
    f[0, 0] = r
    assert f[0, 0].direction == vec3(0, 1, 0) and f[0, 0].end == inf


def test_ray_get():
    # tests/test_ray.py:53-116 — o + t d at 1e-6
    rng = np.random.default_rng(42)
    for _ in range(16):
        o, d, t = rng.normal(size=3), rng.normal(size=3), rng.uniform(0, 10)
        ray = new_ray(vec3(o), vec3(d))
        assert np.allclose(np.asarray(ray.get(t)), o + t * d, atol=1e-5, rtol=1e-6)


def test_new_gaussian_defaults():
    # tests/test_gaussian.py:16-39
    g = new_gaussian()
    assert g.position == vec3(0) and g.rotation == vec4(0, 0, 0, 1) and g.scale == vec3(1, 1, 1)
    assert g.color == vec3(1, 0, 1) and g.opacity == 1
    g = new_gaussian(vec3(1, 2, 3), vec4(0, 0, 0.3826834, 0.9238795), vec3(2, 3, 4), vec3(1, 0, 0), 0.75)
    assert g.position == vec3(1, 2, 3) and g.scale == vec3(2, 3, 4) and g.color == vec3(1, 0, 0) and g.opacity == 0.75


def test_gaussian_field():
    # tests/test_gaussian.py:42-63
    f = Gaussian.field(shape=(32,))
    assert f.shape == (32,)
    g = f[0]
    assert g.position == vec3(0) and g.rotation == vec4(0) and g.scale == vec3(0) and g.color == vec3(0) and g.opacity == 0
    g.init()
    f[0] = g
    assert f[0].rotation == vec4(0, 0, 0, 1) and f[0].scale == vec3(1, 1, 1) and f[0].color == vec3(1, 0, 1)
    assert f.to_numpy().dtype.itemsize == 236          # 59 f32 (gaussian.py:26-55)


def test_gaussian_hit_matches_current_reference_code():
    # gaussian.py:215-230 for the unit Gaussian from its centre along +y: A=1, B=0, C=-3 -> (-sqrt3, +sqrt3).
    # (The reference's own test_gaussian_hit expects (0, inf) and is stale — SURVEY.md §4.)
    t = new_gaussian().hit(new_ray())
    assert np.allclose(np.asarray(t), [-np.sqrt(3), np.sqrt(3)], atol=1e-6)
    miss = new_gaussian().hit(new_ray(vec3(5, 0, 0), vec3(0, 1, 0)))
    assert miss.x == inf and miss.y == inf


def test_gaussian_eval_center():
    g = new_gaussian(opacity=0.5)
    e = g.eval(vec3(0, 0, 0), vec3(0, 0, 1))
    assert np.allclose(np.asarray(e), [1, 0, 1, 0.5], atol=1e-6)     # rho = 1 at the centre, SH all zero


def test_bound_and_bvhnode():
    b = Bound()
    b.init()
    assert b.p_min == vec3(inf) and b.p_max == vec3(-inf)
    b = Bound(vec3(-1, -1, -1), vec3(1, 2, 3))
    assert abs(b.area() - 2 * (2 * 3 + 3 * 4 + 4 * 2)) < 1e-6
    u = b.union(Bound(vec3(0, 0, 0), vec3(5, 1, 1)))
    assert u.p_max == vec3(5, 2, 3) and u.p_min == vec3(-1, -1, -1)
    t = b.hit(new_ray(vec3(0, -5, 0), vec3(0, 1, 0)))
    assert np.allclose(np.asarray(t), [4, 7])
    n = BVHNode()
    n.init()
    assert n.left == -1 and n.right == -1 and n.prim_left == -1 and n.prim_right == -1 and n.depth == -1


def test_vec_types():
    assert vec2i((960, 540)).x == 960 and vec2i((960, 540)).y == 540
    assert vec3i(1, 2, 3).z == 3
    assert np.asarray(vec3(0)).dtype == np.float32 and np.asarray(vec2i(1, 2)).dtype == np.int32
    assert vec4(1, 2, 3, 4).xyz == vec3(1, 2, 3)


def test_sigmoid():
    x = np.array([-2.0, 0.0, 3.0], dtype=np.float32)
    assert sigmoid(x).dtype == np.float32 and np.allclose(sigmoid(x), 1 / (1 + np.exp(-x.astype(np.float64))), atol=1e-6)


def test_synthetic_scene_round_trips_through_a_3dgs_ply(tmp_path):
    """rtgs.synthetic.export_ply writes what the reference's loader (scene.py:95-114) reads: raw columns whose
    activations give the bench scene back."""
    from oracle import ref_numpy as O
    from rtgs.ply import GS_PROPERTIES, read_ply
    from rtgs.synthetic import export_ply, make_scene
    a = make_scene(500, seed=11, sh_degree=3)
    export_ply(tmp_path / "s.ply", a)
    cols = read_ply(tmp_path / "s.ply")
    assert list(cols) == GS_PROPERTIES and len(cols["x"]) == 500
    act = O.activate(cols, 1.0)
    for k, tol in (("pos", 0), ("rot", 1e-6), ("scale", 1e-6), ("color", 1e-6), ("opacity", 1e-6), ("sh", 0)):
        got, want = np.asarray(act[k], np.float64), np.asarray(a[k], np.float64)
        assert got.shape == want.shape, k
        assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max()), k
