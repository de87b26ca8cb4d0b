"""Deep LBVHs: 63-bit Morton codes with many repeated codes give one tree level per differing bit of the sorted
positions of equal codes (csrc/lbvh.cu: k_karras<DUP>), on top of the 63 code levels.  The traversal stacks of
lists_group / k_render / k_trace_closest are sized for the bound RTGS_MAX_TREE_DEPTH = 96 and the batch threshold
of lists_group follows the depth the build measured."""
import numpy as np
import pytest

from oracle import ref_numpy as O

from gpu_util import compare, make_camera, random_set

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _coincident_scene(n_dup, n_free, seed, dup_scale=0.06):
    """n_dup Gaussians on ONE centre (different shapes), n_free ordinary ones around it and 8 outliers a few
    thousand scene radii away (which is what makes 10 bits per axis useless and 21 bits necessary)."""
    rng = np.random.default_rng(seed)
    gs = random_set(n_dup + n_free + 8, seed=seed, mean_scale=0.03)
    gs.pos[:n_dup] = np.array([0.1, -0.05, 0.2])
    gs.scale[:n_dup] = np.exp(rng.normal(np.log(dup_scale), 0.4, (n_dup, 3)))
    gs.opacity[:n_dup] = rng.uniform(0.01, 0.2, n_dup)
    gs.pos[-8:] = rng.uniform(-1, 1, (8, 3)) * 3000.0
    return gs


def test_sixteen_thousand_coincident_centres_render_and_hit():
    from rtgs.ray_tracer import RayTracer
    from rtgs.scene import Scene
    n_dup = 1 << 14
    gs = _coincident_scene(n_dup, 1500, seed=5)
    scene = Scene(morton_bits=63).from_arrays(gs.pos, gs.rot, gs.scale, gs.color, gs.opacity, gs.sh)
    assert scene.morton_bits == 63
    depth = scene.get_option("tree_depth")
    print("tree depth", depth)
    assert 14 <= depth <= 96                      # >= log2(n_dup) levels come from the duplicate rule alone
    cam, ocam = make_camera(0.4, 1.2, 2.4, 48, 32)
    ref = O.render(gs, ocam, depth=16)
    assert np.asarray(ref["nhit"]).max() > 1000   # rays through the pile cross thousands of ellipsoids
    rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
    for mode in (0, 2, 1):
        scene.set_option("render_mode", mode)
        img = rt.render(16).copy()
        mx, ps, bad = compare(img, ref["rgb"], TOL)
        print(f"mode {mode}: max-abs {mx:.2e} psnr {ps:.1f}")
        assert mx <= TOL and ps >= 60.0, mode
    scene.set_option("render_mode", 0)
    # the stack bounds, checked on the device (compute-sanitizer is not available on the GPU pool): high-water marks of
    # the list traversal (256-entry stack, 960-entry group list) and of the fused kernel (512-entry stack)
    rt.render_device(16, collect_stats=True)
    st = rt.last_stats
    print("high water: lists stack", st["max_lists_stack"], "group list", st["max_group_list"], "fused stack", st["max_fused_stack"])
    assert 0 < st["max_lists_stack"] <= 256 and st["max_group_list"] <= 960 and st["max_fused_stack"] <= 512
    assert st["fallback_tiles"] > 0                      # the pile overflows the group lists: the fused kernel ran
    scene.set_option("render_mode", 1)
    rt.render_device(16, collect_stats=True)
    assert 0 < rt.last_stats["max_fused_stack"] <= 512
    scene.set_option("render_mode", 0)
    # closest hit through the same deep tree (k_trace_closest: per-ray stack of depth + 2 entries)
    rays = cam.cam_ray_field.to_numpy().reshape(-1, 8)[::7]
    hit = scene.hit(rays)
    idx, t12 = O.closest_hit(gs, rays[:, :3].astype(np.float64), rays[:, 3:6].astype(np.float64))
    m = idx >= 0
    assert m.sum() > 20
    # coincident centres produce exact ties only by accident; compare distances, and indices where they are unique
    assert np.array_equal(hit.gaussian_idx >= 0, m)
    assert np.allclose(hit.intersections[m, 0], t12[m, 0], rtol=1e-6)


def test_a_million_coincident_centres_stay_within_the_stack_bound():
    """2^20 Gaussians on one centre: ~20 duplicate levels on top of the code levels.  The build reports the depth,
    every kernel that walks the tree completes, and the list path agrees with the fused kernel."""
    from rtgs.ray_tracer import RayTracer
    from rtgs.scene import Scene
    n_dup = 1 << 20
    gs = _coincident_scene(n_dup, 64, seed=6, dup_scale=0.01)
    scene = Scene(morton_bits="auto").from_arrays(gs.pos, gs.rot, gs.scale, gs.color, gs.opacity, gs.sh)
    assert scene.morton_bits == 63                # chosen automatically: nearly every 30-bit code repeats
    depth = scene.get_option("tree_depth")
    print("tree depth", depth)
    assert 20 <= depth <= 96
    cam, _ = make_camera(0.4, 1.2, 2.4, 8, 8)
    rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
    imgs = {}
    for mode in (2, 0, 1):
        scene.set_option("render_mode", mode)
        imgs[mode] = rt.render(16).copy()
        assert np.isfinite(imgs[mode]).all()
    assert np.array_equal(imgs[0], imgs[2])
    assert np.abs(imgs[1] - imgs[0]).max() <= 1e-5
    for mode in (0, 1):
        scene.set_option("render_mode", mode)
        rt.render_device(16, collect_stats=True)
        st = rt.last_stats
        print(f"mode {mode} high water: lists stack {st['max_lists_stack']} group list {st['max_group_list']} "
              f"fused stack {st['max_fused_stack']}")
        assert st["max_lists_stack"] <= 256 and st["max_group_list"] <= 960 and 0 < st["max_fused_stack"] <= 512
    scene.set_option("render_mode", 0)
    rays = cam.cam_ray_field.to_numpy().reshape(-1, 8)
    hit = scene.hit(rays)
    assert (hit.gaussian_idx >= 0).any()
    m = hit.gaussian_idx >= 0
    assert np.isfinite(hit.intersections[m]).all() and (hit.depth[m] >= 20).any()


def test_non_finite_geometry_is_rejected():
    """NaN / Inf positions, rotations or scales have no place in Morton codes or boxes: rtgs_scene_create refuses
    them (RTGS_ERR_INVALID) instead of building an undefined tree."""
    from rtgs import _native
    from rtgs.scene import Scene
    gs = random_set(100, seed=9)
    for field, bad in (("pos", np.nan), ("pos", np.inf), ("scale", np.nan), ("rot", -np.inf)):
        a = {k: np.array(getattr(gs, k), np.float32, copy=True) for k in ("pos", "rot", "scale", "color", "opacity", "sh")}
        a[field][17, 1] = bad
        with pytest.raises(_native.RtgsError) as e:
            Scene().from_arrays(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
        assert e.value.status == -1 and "finite" in str(e.value)
