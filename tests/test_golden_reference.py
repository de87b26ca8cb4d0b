"""Pins the oracle to the REFERENCE'S OWN CODE: tests/golden/reference_golden_fp64.npz was produced by
importing /root/reference/src/rtgs unmodified and executing it through oracle/taichi_shim (a pure-Python
Taichi stand-in; tests/golden/make_golden.py).  It holds known-answer vectors for every function on the
path and the output of the full reference pipeline (Scene.load_file incl. the reference's SAH BVH build,
Camera, RayTracer.sample x 16, generate_disp_buffer) on tests/data/test.ply.

CPU tests: oracle == golden to float64 round-off.  GPU test: product == golden within 1e-3 / 60 dB."""
import numpy as np
import pytest

from oracle import ref_numpy as O

from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN / "reference_golden_fp64.npz")


@pytest.fixture(scope="module")
def g32():
    return np.load(GOLDEN / "reference_golden_fp32.npz")


def _kat_set(g):
    return O.GaussianSet(g["kat_g_pos"], g["kat_g_rot"], g["kat_g_scale"], g["kat_g_color"], g["kat_g_opacity"],
                         g["kat_g_sh"])


def _pipe_set(g):
    return O.GaussianSet(g["pipe_pos"], g["pipe_rot"], g["pipe_scale"], g["pipe_color"], g["pipe_opacity"],
                         g["pipe_sh"])


def _pipe_cam(g):
    res = int(g["pipe_res"])
    return O.CameraParams(g["pipe_cam_pos"], g["pipe_cam_rot"], res, res, (float(g["pipe_focal"]),) * 2)


def test_quaternion_kats(g):
    assert np.abs(O.quat_mul(g["kat_p"], g["kat_q"]) - g["kat_mul"]).max() < 1e-12
    assert np.abs(O.quat_conj(g["kat_q"]) - g["kat_conj"]).max() == 0
    assert np.abs(O.rot_vec3(g["kat_q"], g["kat_v"]) - g["kat_rot"]).max() < 1e-12
    assert np.abs(O.as_rotation_mat3(g["kat_q"]) - g["kat_mat3"]).max() < 1e-12


def test_gaussian_kats(g):
    gs = _kat_set(g)
    assert np.abs(O.covariance(gs.rot, gs.scale) - g["kat_cov"]).max() < 1e-12
    lo, hi = O.bounding_box_reference(gs.pos, gs.rot, gs.scale)
    assert np.abs(lo - g["kat_bmin"]).max() < 1e-12 and np.abs(hi - g["kat_bmax"]).max() < 1e-12
    n = gs.n
    for i in range(n):
        sub = O.GaussianSet(gs.pos[i:i + 1], gs.rot[i:i + 1], gs.scale[i:i + 1], gs.color[i:i + 1],
                            gs.opacity[i:i + 1], gs.sh[i:i + 1])
        t1, t2 = O.intersect_all(sub, g["kat_ray_o"][i:i + 1], g["kat_ray_d"][i:i + 1])
        want = g["kat_hit"][i]
        assert np.isfinite(want[0]) == np.isfinite(t1[0, 0])
        if np.isfinite(want[0]):
            assert np.allclose([t1[0, 0], t2[0, 0]], want, rtol=1e-9)
            # Gaussian.eval at the mid point: rgb + alpha (gaussian.py:183-201)
            o = g["kat_ray_o"][i].astype(np.float64)
            d = g["kat_ray_d"][i].astype(np.float64)
            pos = o + 0.5 * (t1[0, 0] + t2[0, 0]) * d
            dv = pos - gs.pos[i].astype(np.float64)
            Minv = O.cov_inverse(gs.rot[i], gs.scale[i])
            alpha = float(gs.opacity[i]) * np.exp(-dv @ Minv @ dv)
            col = gs.color[i].astype(np.float64) + O.sh_basis(d / np.linalg.norm(d)) @ gs.sh[i].astype(np.float64)
            assert np.allclose(np.r_[col, alpha], g["kat_eval"][i], rtol=1e-8, atol=1e-12)
    assert np.isfinite(g["kat_hit"][:, 0]).sum() >= 10
    dn = g["kat_ray_d"].astype(np.float64)
    dn /= np.linalg.norm(dn, axis=1, keepdims=True)
    esh = np.einsum("nk,nkc->nc", O.sh_basis(dn), gs.sh.astype(np.float64))
    assert np.abs(esh - g["kat_eval_sh"]).max() < 1e-12


def test_bound_hit_and_camera_kats(g):
    t0, t1 = O.bound_hit(g["kat_box_lo"], g["kat_box_hi"], g["kat_ray_o"], g["kat_ray_d"])
    assert np.abs(np.stack([t0, t1], -1) - g["kat_box_hit"]).max() < 1e-12
    cam = O.CameraParams(g["kat_cam_pos"], g["kat_cam_rot"], 7, 5, (6.0, 6.5))
    o, d = O.camera_rays(cam)
    rays = g["kat_cam_rays"]
    assert np.abs(d.reshape(7, 5, 3) - rays[..., 3:6]).max() < 1e-12
    assert np.abs(o - rays[..., :3]).max() == 0 and (rays[..., 6] == 0).all() and np.isinf(rays[..., 7]).all()


def test_loader_matches_reference_load_file(g, test_ply):
    """Scene.load_file of the reference (PLY -> activations, scene.py:95-128) vs the oracle's restatement.
    The reference's BVH build permutes gaussian_field, so compare as sets (sorted by position)."""
    from rtgs.ply import read_ply
    a = O.activate(read_ply(test_ply), float(g["pipe_scale_arg"]), sh_layout="interleaved")
    ia, ib = np.lexsort(a["pos"].T), np.lexsort(g["pipe_pos"].T)
    for k in ("pos", "rot", "scale", "color", "opacity", "sh"):
        assert np.array_equal(a[k][ia], g["pipe_" + k][ib]), k


def test_pipeline_image_fp64(g):
    """The reference's RayTracer output (K closest-hit restarts through ITS OWN SAH BVH, float64 shim) ==
    the brute-force K-nearest oracle, to round-off: pins the compositing semantics of SURVEY.md §3.3."""
    out = O.render(_pipe_set(g), _pipe_cam(g), depth=int(g["pipe_depth"]))
    assert (g["pipe_sample_buf"].max(axis=-1) > 0).sum() > 100
    assert np.abs(out["rgb"] - g["pipe_sample_buf"]).max() < 1e-9
    assert np.abs(out["T"] - g["pipe_attenuation"]).max() < 1e-9
    assert np.abs(out["rgb"] - g["pipe_disp"]).max() < 1e-9           # one sample: disp == sample


def test_pipeline_image_fp32_shows_reference_noise(g, g32):
    """The same pipeline evaluated in float32 (Taichi's default precision) differs from the float64
    evaluation by isolated silhouette flips of up to ~0.05 per layer — why parity is defined on float64
    (SURVEY.md §7 hard part 1)."""
    d = np.abs(g32["pipe_sample_buf"] - g["pipe_sample_buf"]).max(axis=-1)
    assert np.median(d) < 1e-5
    assert d.max() < 0.5


def test_cpp_port_matches_golden(g):
    from oracle import ref_cpu
    gs, cam = _pipe_set(g), _pipe_cam(g)
    cs = ref_cpu.CpuScene(gs.pos, gs.rot, gs.scale, gs.color, gs.opacity, gs.sh)
    out = cs.render(cam, int(g["pipe_depth"]), precision="double")
    res = int(g["pipe_res"])
    assert np.abs(out["rgb"].reshape(res, res, 3) - g["pipe_sample_buf"]).max() < 1e-9


@pytest.mark.gpu
def test_product_matches_reference_golden(g, test_ply):
    """The CUDA path, through Scene.load_file (same SH layout the reference executed) / Camera / RayTracer,
    against the reference's own render: max-abs <= 1e-3, PSNR >= 60 dB."""
    from rtgs.camera import Camera
    from rtgs.ray_tracer import RayTracer
    from rtgs.scene import Scene
    res = int(g["pipe_res"])
    scene = Scene(128, 1, 8).load_file(test_ply, float(g["pipe_scale_arg"]), sh_layout="interleaved")
    cam = Camera(g["pipe_cam_pos"], g["pipe_cam_rot"], (res, res), (float(g["pipe_focal"]),) * 2)
    rt = RayTracer((res, res), scene, cam, t_cut=0.0)
    for _ in range(int(g["pipe_depth"])):
        rt.sample(int(g["pipe_depth"]))
    rt.generate_disp_buffer(rt.num_samples, rt.num_steps, int(g["pipe_depth"]))
    img = rt.disp_buf.to_numpy()
    err = np.abs(img - g["pipe_disp"]).max()
    assert err <= 1e-3 and O.psnr(img, g["pipe_disp"]) >= 60.0
    assert np.abs(rt.attenuation_buf.to_numpy() - g["pipe_attenuation"]).max() <= 1e-3


# ---- second reference-run golden: saturated k-buffer, multi-level SAH tree (tests/golden/make_golden_ksat.py) ----
@pytest.fixture(scope="module")
def k64():
    return np.load(GOLDEN / "reference_ksat_fp64.npz")


@pytest.fixture(scope="module")
def k32():
    return np.load(GOLDEN / "reference_ksat_fp32.npz")


def _ksat(g):
    gs = O.GaussianSet(g["pos"], g["rot"], g["scale"], g["color"], g["opacity"], g["sh"])
    res = int(g["res"])
    return gs, O.CameraParams(g["cam_pos"], g["cam_rot"], res, res, (float(g["focal"]),) * 2), res


def test_ksat_golden_exercises_truncation_and_a_deep_reference_tree(k64):
    """The fixture is only worth something if the reference really had to truncate: a quarter of the rays cross more
    than 16 ellipsoids (up to 29), and the reference's own SAH builder (Scene(1024, 4, 16), __main__.py:97)
    produced a tree of depth >= 4."""
    gs, cam, res = _ksat(k64)
    nhit = np.asarray(O.render(gs, cam, depth=int(k64["depth"]))["nhit"]).reshape(-1)
    assert nhit.max() > 24 and (nhit > 16).mean() > 0.2
    assert k64["bvh_int"][:, 4].max() >= 4 and k64["bvh_int"].shape[0] >= 30


def test_ksat_oracle_matches_reference_run(k64):
    """K closest-hit restarts with ray.start = t1 + 1e-8 (ray_tracer.py:96-104) and the strict start < t1 rule
    (scene.py:433) through the reference's own multi-level tree == the brute-force 16-nearest oracle, and == the
    C++ restart port with either restart epsilon."""
    from oracle import ref_cpu
    gs, cam, res = _ksat(k64)
    out = O.render(gs, cam, depth=int(k64["depth"]))
    assert np.abs(out["rgb"] - k64["sample_buf"]).max() < 1e-9
    assert np.abs(out["T"] - k64["attenuation"]).max() < 1e-9
    cs = ref_cpu.CpuScene(gs.pos, gs.rot, gs.scale, gs.color, gs.opacity, gs.sh)
    for eps in (0.0, 1e-8):
        o = cs.render(cam, int(k64["depth"]), precision="double", restart_eps=eps)
        assert np.abs(o["rgb"].reshape(res, res, 3) - k64["sample_buf"]).max() < 1e-9, eps


def test_ksat_fp32_reference_agrees_with_fp64_here(k64, k32):
    """On this scene the reference's float32 evaluation (Taichi's default) stays within 1e-5 of its float64
    evaluation at every pixel: the 1e-3 tolerance of the CUDA path holds against either."""
    d = np.abs(k32["sample_buf"] - k64["sample_buf"]).max(axis=-1)
    assert d.max() < 1e-5


@pytest.mark.gpu
def test_product_matches_ksat_golden(k64, k32):
    """The CUDA path against reference-produced output where the k-buffer saturates (16 of up to 29 crossings kept):
    Scene.load_file of the same PLY, every render mode; reports how many pixels differ from the float32 reference
    run by more than 1e-3 (none)."""
    from conftest import DATA
    from rtgs.camera import Camera
    from rtgs.ray_tracer import RayTracer
    from rtgs.scene import Scene
    res, depth = int(k64["res"]), int(k64["depth"])
    scene = Scene(1024, 4, 16).load_file(DATA / "ksat.ply", 1.0, sh_layout="interleaved")
    g = scene.read_gaussians()
    ia, ib = np.lexsort(g["pos"].T), np.lexsort(k64["pos"].T)       # the reference's build permutes its field
    for k in ("pos", "rot", "scale", "color", "opacity", "sh"):
        assert np.array_equal(g[k][ia], k64[k][ib]), k
    cam = Camera(k64["cam_pos"], k64["cam_rot"], (res, res), (float(k64["focal"]),) * 2)
    rt = RayTracer((res, res), scene, cam, t_cut=0.0)
    for mode in (0, 2, 1):
        scene.set_option("render_mode", mode)
        rt.clear_sample()
        rt.num_steps = rt.num_samples = 0
        for _ in range(depth):
            rt.sample(depth)
        rt.generate_disp_buffer(rt.num_samples, rt.num_steps, depth)
        img = rt.disp_buf.to_numpy()
        err64 = np.abs(img - k64["disp"]).max()
        bad32 = int((np.abs(img - k32["disp"]).max(axis=-1) > 1e-3).sum())
        print(f"ksat mode {mode}: max-abs vs fp64 reference run {err64:.2e}, PSNR {O.psnr(img, k64['disp']):.1f} dB, "
              f"pixels off by > 1e-3 vs the fp32 reference run: {bad32}")
        assert err64 <= 1e-3 and O.psnr(img, k64["disp"]) >= 60.0 and bad32 == 0, mode
        assert np.abs(rt.attenuation_buf.to_numpy() - k64["attenuation"]).max() <= 1e-3
