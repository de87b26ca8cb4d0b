"""Self-consistency and analytic checks of the float64 image oracle (oracle/ref_numpy.py):
analytic single-Gaussian cases, literal per-pair evaluation of the reference formulas, loader
activations against the 16-Gaussian PLY, orbit pose / focal formulas."""
import numpy as np
import pytest

from oracle import ref_numpy as O
from rtgs.gaussian import Gaussian
from rtgs.ply import read_ply
from rtgs.ray import new_ray
from rtgs.utils.types import vec3


def _unit_set(opacity=0.5, sh=None):
    return O.GaussianSet(pos=[[0, 0, 0]], rot=[[0, 0, 0, 1]], scale=[[1, 1, 1]], color=[[0.2, 0.4, 0.6]],
                         opacity=[opacity], sh=sh)


def test_unit_gaussian_on_axis():
    # ray from (0,-5,0) along +y through the centre: t1 = 5 - sqrt3, t2 = 5 + sqrt3 (gaussian.py:215-226)
    gs = _unit_set()
    t1, t2 = O.intersect_all(gs, [[0, -5, 0]], [[0, 1, 0]])
    assert np.allclose([t1[0, 0], t2[0, 0]], [5 - np.sqrt(3), 5 + np.sqrt(3)], atol=1e-12)
    # off-axis by x: q_min = x^2 -> hit iff x^2 < 3
    t1, _ = O.intersect_all(gs, [[1.7, -5, 0], [1.74, -5, 0]], [[0, 1, 0], [0, 1, 0]])
    assert np.isfinite(t1[0, 0]) and np.isinf(t1[1, 0])


def test_single_gaussian_pixel_value():
    # camera on the -y axis looking at the origin: centre pixel sees alpha = opacity * exp(-q_min)
    gs = _unit_set(opacity=0.5)
    pos, rot = O.orbit_pose(-np.pi / 2, np.pi / 2, 5.0)      # position (0,-5,0)
    cam = O.CameraParams(pos, rot, 3, 3, (100.0, 100.0))
    out = O.render(gs, cam, depth=4)
    assert np.allclose(out["rgb"][1, 1], 0.5 * np.array([0.2, 0.4, 0.6]), atol=1e-6)
    assert np.allclose(out["T"][1, 1], 0.5, atol=1e-6)
    assert out["nhit"].reshape(3, 3)[1, 1] == 1


def test_matches_literal_reference_formulas():
    # compare the vectorised oracle with a literal per-pair evaluation through the API mirror's
    # Gaussian.hit / Gaussian.eval (which transcribe gaussian.py:183-230 one-to-one)
    rng = np.random.default_rng(3)
    n = 12
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    gs = O.GaussianSet(pos=rng.uniform(-0.5, 0.5, (n, 3)), rot=q, scale=np.exp(rng.normal(-2.0, 0.4, (n, 3))),
                       color=rng.uniform(0, 1, (n, 3)), opacity=rng.uniform(0.2, 1, n),
                       sh=rng.normal(0, 0.15, (n, 15, 3)))
    pos, rot = O.orbit_pose(0.3, 1.2, 2.0)
    cam = O.CameraParams(pos, rot, 8, 6, (6.0, 6.0))
    out = O.render(gs, cam, depth=5, return_layers=True)
    o, dirs = O.camera_rays(cam)
    gl = []
    for k in range(n):
        g = Gaussian(gs.pos[k], gs.rot[k], gs.scale[k], gs.color[k], gs.opacity[k])
        for j, name in enumerate(g.__slots__[5:]):
            setattr(g, name, vec3(gs.sh[k, j]))
        gl.append(g)
    img = np.zeros((8 * 6, 3))
    for r in range(8 * 6):
        hits = []
        for k, g in enumerate(gl):
            # float64 literal: reuse the mirror's maths but on float64 ray data
            cov_inv = np.linalg.inv(g.cov())
            v = o - gs.pos[k].astype(np.float64)
            d = dirs[r]
            A, B, C = d @ cov_inv @ d, 2 * d @ cov_inv @ v, v @ cov_inv @ v - 3
            delta = B * B - 4 * A * C
            if delta > 0:
                t1, t2 = (-B - np.sqrt(delta)) / (2 * A), (-B + np.sqrt(delta)) / (2 * A)
                if t1 > 0:
                    hits.append((t1, t2, k))
        hits.sort()
        T, acc = 1.0, np.zeros(3)
        for t1, t2, k in hits[:5]:
            p = o + 0.5 * (t1 + t2) * dirs[r]
            e = np.asarray(gl[k].eval(p, dirs[r]), dtype=np.float64)
            # Gaussian.eval returns float32; redo alpha/colour in float64 for a tight comparison
            dv = p - gs.pos[k].astype(np.float64)
            alpha = float(gs.opacity[k]) * np.exp(-dv @ np.linalg.inv(gl[k].cov()) @ dv)
            col = gs.color[k].astype(np.float64) + O.sh_basis(dirs[r] / np.linalg.norm(dirs[r])) @ gs.sh[k].astype(np.float64)
            assert np.allclose(e[:3], col, atol=1e-5) and abs(e[3] - alpha) < 1e-6
            acc += T * alpha * col
            T *= 1 - alpha
        img[r] = acc
    assert np.allclose(out["rgb"].reshape(-1, 3), img, atol=1e-10)


def test_closest_hit_is_first_layer():
    rng = np.random.default_rng(5)
    n = 40
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    gs = O.GaussianSet(pos=rng.uniform(-1, 1, (n, 3)), rot=q, scale=np.exp(rng.normal(-1.6, 0.4, (n, 3))),
                       color=rng.uniform(0, 1, (n, 3)), opacity=rng.uniform(0.2, 1, n))
    pos, rot = O.orbit_pose(1.0, 1.0, 3.0)
    cam = O.CameraParams(pos, rot, 16, 12, (14.0, 14.0))
    out = O.render(gs, cam, depth=3, return_layers=True)
    o, dirs = O.camera_rays(cam)
    idx, t12 = O.closest_hit(gs, o[None], dirs)
    assert np.array_equal(idx, out["layers"]["idx"][:, 0])
    # restarting from the previous entry point reproduces the second layer (ray_tracer.py:100-102)
    hit = idx >= 0
    idx2, _ = O.closest_hit(gs, o[None], dirs[hit], start=t12[hit, 0])
    assert np.array_equal(idx2, out["layers"]["idx"][hit, 1])
    # generic-origin code path agrees with the common-origin path
    oo = np.broadcast_to(o, dirs.shape).copy()
    oo[0] += 1e-9
    idx3, t3 = O.closest_hit(gs, oo, dirs)
    assert np.array_equal(idx3, idx) and np.allclose(t3[hit], t12[hit], rtol=1e-7)


def test_reference_aabb_contains_tight_aabb():
    # the reference's six-endpoint box is a superset of the sqrt(3)-sigma ellipsoid box
    rng = np.random.default_rng(7)
    q = rng.normal(size=(2000, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    s = np.exp(rng.normal(-3, 0.7, (2000, 3)))
    p = rng.uniform(-1, 1, (2000, 3))
    lo_r, hi_r = O.bounding_box_reference(p, q, s)
    lo_t, hi_t = O.bounding_box_tight(p, q, s)
    assert (lo_r <= lo_t + 1e-12).all() and (hi_r >= hi_t - 1e-12).all()


def test_activation_of_test_ply(test_ply):
    cols = read_ply(test_ply)
    assert len(cols["x"]) == 16 and len(cols) == 62
    a = O.activate(cols, scale=30.0)
    assert a["pos"].dtype == np.float32 and a["sh"].shape == (16, 15, 3)
    assert np.allclose(np.linalg.norm(a["rot"], axis=1), 1, atol=1e-6)
    assert np.allclose(a["scale"], np.exp(np.stack([cols[f"scale_{i}"] for i in range(3)], -1)) * 30, rtol=1e-6)
    assert np.array_equal(a["sh"][:, 2, 1], cols["f_rest_17"])          # channel-major: sh_k[c] = f_rest_{15c+k}
    b = O.activate(cols, scale=30.0, sh_layout="interleaved")
    assert np.array_equal(b["sh"][:, 2, 1], cols["f_rest_7"])           # flat[3k + c]


def test_orbit_pose_and_focal():
    # __main__.py:91-92 and :120-142: theta=0, phi=pi/2, r=1 -> position (1,0,0), looking down -x, up +z
    pos, rot = O.orbit_pose(0.0, np.pi / 2, 1.0)
    assert np.allclose(pos, [1, 0, 0], atol=1e-7)
    assert np.allclose(O.rot_vec3(rot.astype(np.float64), [0, 0, -1]), [-1, 0, 0], atol=1e-6)   # forward
    assert np.allclose(O.rot_vec3(rot.astype(np.float64), [1, 0, 0]), [0, 1, 0], atol=1e-6)    # right
    assert np.allclose(O.rot_vec3(rot.astype(np.float64), [0, 1, 0]), [0, 0, 1], atol=1e-6)    # up
    assert abs(O.focal_from_fov(256, 90.0) - 128.0) < 1e-9
    from rtgs.orbit import focal_from_fov, orbit_pose
    p2, r2 = orbit_pose(0.7, 1.1, 2.2)
    p1, r1 = O.orbit_pose(0.7, 1.1, 2.2)
    assert np.allclose(np.asarray(p2), p1) and np.allclose(np.asarray(r2), r1, atol=1e-7)
    assert focal_from_fov(1080, 60.0) == pytest.approx(O.focal_from_fov(1080, 60.0))


def test_camera_rays_centre_and_layout():
    pos, rot = O.orbit_pose(0.0, np.pi / 2, 2.0)
    cam = O.CameraParams(pos, rot, 5, 3, (4.0, 4.0))
    o, d = O.camera_rays(cam)
    d = d.reshape(5, 3, 3)
    assert np.allclose(d[2, 1], [-1, 0, 0], atol=1e-6)             # centre pixel looks at the origin
    assert d[4, 1, 1] > d[0, 1, 1]                                  # i grows to the right (+y here)
    assert d[2, 2, 2] > d[2, 0, 2]                                  # j grows upward (+z here)


# ---- size-independent properties of the image oracle ---------------------------------------------------------------
def _rand_set(n, seed, sh=False):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return O.GaussianSet(pos=rng.uniform(-1, 1, (n, 3)), rot=q, scale=np.exp(rng.normal(np.log(0.12), 0.5, (n, 3))),
                         color=rng.uniform(0, 1, (n, 3)), opacity=rng.uniform(0.1, 0.95, n),
                         sh=rng.normal(0, 0.15, (n, 15, 3)) if sh else None)


def test_image_is_invariant_under_a_rigid_motion_of_scene_and_camera():
    """Without SH (whose colour depends on the world direction) moving every Gaussian and the camera by the same
    rotation + translation must not change a pixel: pins the quaternion conventions of Sigma = R S S^T R^T
    (gaussian.py:86-102) against those of the camera rays (camera.py:31-55)."""
    gs = _rand_set(300, 5)
    pos, rot = O.orbit_pose(0.4, 1.1, 2.6)
    cam = O.CameraParams(np.asarray(pos), np.asarray(rot), 48, 32, (40.0, 40.0))
    a = O.render(gs, cam, depth=16)
    g = np.array([0.3, -0.5, 0.2, 0.79])
    g /= np.linalg.norm(g)
    t = np.array([0.7, -1.3, 2.1])
    gs2 = O.GaussianSet(pos=O.rot_vec3(g, gs.pos) + t, rot=O.quat_mul(g, gs.rot), scale=gs.scale, color=gs.color,
                        opacity=gs.opacity, sh=None)
    cam2 = O.CameraParams(O.rot_vec3(g, np.asarray(pos)) + t, O.quat_mul(g, np.asarray(rot)), 48, 32, (40.0, 40.0))
    b = O.render(gs2, cam2, depth=16)
    # (the moved parameters are rounded to float32 again when the set is built: invariance up to that rounding)
    assert np.abs(a["rgb"] - b["rgb"]).max() < 2e-6 and (a["nhit"] != b["nhit"]).mean() < 1e-3


def test_depth_prefix_permutation_and_colour_linearity():
    gs = _rand_set(400, 9, sh=True)
    pos, rot = O.orbit_pose(1.0, 1.3, 2.4)
    cam = O.CameraParams(np.asarray(pos), np.asarray(rot), 40, 30, (36.0, 36.0))
    full = O.render(gs, cam, depth=16)
    # the order of the Gaussians in memory does not matter
    perm = np.random.default_rng(1).permutation(gs.n)
    gp = O.GaussianSet(pos=gs.pos[perm], rot=gs.rot[perm], scale=gs.scale[perm], color=gs.color[perm],
                       opacity=gs.opacity[perm], sh=gs.sh[perm])
    assert np.abs(O.render(gp, cam, depth=16)["rgb"] - full["rgb"]).max() < 1e-12
    # front-to-back compositing: T never increases with depth, and every partial sum is a prefix of the next
    prev = O.render(gs, cam, depth=1)
    for d in (2, 4, 8, 16):
        cur = O.render(gs, cam, depth=d)
        assert (cur["T"] <= prev["T"] + 1e-15).all()
        same = np.asarray(cur["nhit"]).reshape(-1) <= d // 2                        # rays that gained no layer
        assert same.any() and (~same).any()
        assert np.abs(cur["rgb"].reshape(-1, 3)[same] - prev["rgb"].reshape(-1, 3)[same]).max() < 1e-12
        assert (np.abs(cur["rgb"].reshape(-1, 3)[~same] - prev["rgb"].reshape(-1, 3)[~same]).max(axis=1) > 0).mean() > 0.9
        prev = cur
    # the image is linear in the colours: scaling DC colour and SH by s scales every pixel by s
    gl = O.GaussianSet(pos=gs.pos, rot=gs.rot, scale=gs.scale, color=0.5 * gs.color, opacity=gs.opacity, sh=0.5 * gs.sh)
    assert np.abs(O.render(gl, cam, depth=16)["rgb"] - 0.5 * full["rgb"]).max() < 1e-12
