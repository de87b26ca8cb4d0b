"""One-off fuzz 4: degenerate image shapes (1x1, slivers), large sparse images, non-unit camera quaternions, extreme fov,
single / few Gaussians, duplicates - all routes against the float64 oracle."""
import sys, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import ref_numpy as O
from gpu_util import make_scene
from rtgs.camera import Camera
from rtgs.orbit import focal_from_fov, orbit_pose
from rtgs.ray_tracer import RayTracer
lo, hi = int(sys.argv[1]), int(sys.argv[2]); worst = 0.0; nfail = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(61000 + seed)
    n = int(rng.choice([1, 2, 3, 7, 33, 300, 3000]))
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    pos = rng.uniform(-1, 1, (n, 3)); scale = np.exp(rng.normal(np.log(10 ** rng.uniform(-2, -0.3)), 0.7, (n, 3)))
    if n > 4 and rng.random() < 0.3:              # exact duplicates of position (different shapes)
        pos[n // 2:] = pos[: n - n // 2]
    gs = O.GaussianSet(pos=pos, rot=q, scale=scale, color=rng.uniform(0, 1, (n, 3)), opacity=rng.uniform(0.01, 0.99, n),
                       sh=rng.normal(0, 0.15, (n, 15, 3)) if rng.random() < 0.5 else None)
    scene = make_scene(gs)
    shape = rng.integers(0, 4)
    W, H = [(1, 1), (int(rng.integers(1, 4)), int(rng.integers(50, 300))), (int(rng.integers(50, 400)), int(rng.integers(1, 4))),
            (int(rng.integers(100, 700)), int(rng.integers(100, 400)))][shape]
    pos_c, rot_c = orbit_pose(float(rng.uniform(0, 6.28)), float(rng.uniform(0.05, 3.09)), float(10 ** rng.uniform(-1, 0.7)))
    rot_c = np.asarray(rot_c, np.float64) * (float(rng.uniform(0.5, 2.0)) if rng.random() < 0.3 else 1.0)   # rotation used as given
    fov = float(rng.choice([rng.uniform(1, 10), rng.uniform(10, 120), rng.uniform(120, 170)]))
    f = focal_from_fov(max(H, 2), fov)
    fy = f * float(rng.choice([1.0, 0.5, 2.0]))
    cam = Camera(pos_c, rot_c, (W, H), (f, fy))
    ocam = O.CameraParams(np.asarray(pos_c), np.asarray(rot_c), W, H, (f, fy))
    depth = int(rng.choice([1, 16, 16, 32]))
    ref = O.render(gs, ocam, depth=depth)["rgb"].reshape(W, H, 3)
    rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
    errs = []
    for mode in (0, 2, 1):
        scene.set_option("render_mode", mode)
        errs.append(float(np.abs(rt.render(depth) - ref).max()))
    scene.set_option("render_mode", 0)
    mx = max(errs); worst = max(worst, mx); ok = mx <= 1e-3; nfail += not ok
    print(f"seed {seed}: n={n} {W}x{H} depth={depth} fov={fov:.0f} |q|={np.linalg.norm(rot_c):.2f} errs {[f'{e:.1e}' for e in errs]}{'' if ok else '  <<<<<< FAIL'}", flush=True)
print("worst", worst, "failures", nfail)
