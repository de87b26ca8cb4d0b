"""One-off fuzz sweep 3 (round 2): larger scenes (up to 40 k Gaussians, clusters of clusters), region renders, stripes,
the opt-in depth-slab lists, the blocking and the pipelined host delivery - all against the float64 oracle."""
import sys, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import ref_numpy as O
from gpu_util import compare, make_scene
from rtgs.camera import Camera
from rtgs.orbit import focal_from_fov, orbit_pose
from rtgs.ray_tracer import RayTracer
lo, hi = int(sys.argv[1]), int(sys.argv[2])
worst = 0.0; nfail = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(47000 + seed)
    n = int(10 ** rng.uniform(2, 4.6))
    ms = float(10.0 ** rng.uniform(-2.7, -1.0))
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    nc = int(rng.integers(1, 12))
    centres = rng.uniform(-0.8, 0.8, (nc, 3)); radii = 10 ** rng.uniform(-2.5, -0.3, nc)
    k = rng.integers(0, nc, n)
    pos = centres[k] + rng.normal(0, 1, (n, 3)) * radii[k, None]
    if rng.random() < 0.3: pos[: n // 3] = rng.uniform(-1, 1, (n // 3, 3))
    scale = np.exp(rng.normal(np.log(ms), float(rng.choice([0.3, 1.0])), (n, 3)))
    gs = O.GaussianSet(pos=pos, rot=q, scale=scale, color=1 / (1 + np.exp(-rng.normal(0, 1, (n, 3)))),
                       opacity=1 / (1 + np.exp(-rng.normal(0, 1.5, n))), sh=rng.normal(0, 0.15, (n, 15, 3)) if rng.random() < 0.5 else None)
    scene = make_scene(gs)
    W, H = int(rng.integers(16, 120)), int(rng.integers(16, 90))
    depth = int(rng.choice([2, 16, 16, 16]))
    target = centres[int(rng.integers(0, nc))]
    r = float(10 ** rng.uniform(-1.5, 0.5))
    pos_c, rot_c = orbit_pose(float(rng.uniform(0, 6.28)), float(rng.uniform(0.3, 2.8)), r)
    pos_c = np.asarray(pos_c) + target          # orbit one of the clusters
    f = focal_from_fov(H, float(rng.uniform(25, 110)))
    cam = Camera(pos_c, rot_c, (W, H), (f, f))
    ocam = O.CameraParams(np.asarray(pos_c), np.asarray(rot_c), W, H, (f, f))
    ref = O.render(gs, ocam, depth=depth)["rgb"].reshape(W, H, 3)
    rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
    errs = {}
    for mode in (0, 2, 1):
        scene.set_option("render_mode", mode)
        errs[f"m{mode}"] = compare(rt.render(depth), ref.reshape(-1, 3), 1e-3)[0] if False else float(np.abs(rt.render(depth) - ref).max())
    scene.set_option("render_mode", 0)
    base = rt.render(depth).copy()
    # depth-slab lists, with a random heavy limit
    scene.set_option("heavy_limit", int(rng.choice([0, 16, 200, -1]))); scene.set_option("heavy_lists", 2)
    img = rt.render(depth).copy(); errs["slab"] = float(np.abs(img - ref).max()); same = np.array_equal(img, base)
    scene.set_option("heavy_lists", 0); scene.set_option("heavy_limit", -1)
    # a random region
    x0, y0 = int(rng.integers(0, W - 4)), int(rng.integers(0, H - 4)); w, h = int(rng.integers(1, W - x0 + 1)), int(rng.integers(1, H - y0 + 1))
    reg = rt.render(depth, tile=(x0, y0, w, h))
    reg = np.asarray(reg); reg = reg[x0:x0 + w, y0:y0 + h] if reg.shape[0] == W else reg
    errs["region"] = float(np.abs(reg - ref[x0:x0 + w, y0:y0 + h]).max())
    # pipelined delivery
    fut = rt.render_async(depth); errs["async"] = float(np.abs(fut.result() - ref).max())
    mx = max(errs.values()); worst = max(worst, mx)
    ok = mx <= 1e-3 and same
    nfail += not ok
    print(f"seed {seed}: n={n} ms={ms:.4f} clusters={nc} {W}x{H} depth={depth} r={r:.3f} " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()) +
          f" slab identical {same}{'' if ok else '  <<<<<< FAIL'}", flush=True)
print("worst", worst, "failures", nfail)
