"""One-off fuzz 6: random render OPTIONS (list pool size incl. none, forced 30- / 63-bit Morton codes, heavy limit, slab
lists, render mode, t_cut) on clustered scenes against the float64 oracle."""
import sys, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import ref_numpy as O
from rtgs.camera import Camera
from rtgs.orbit import focal_from_fov, orbit_pose
from rtgs.ray_tracer import RayTracer
from rtgs.scene import Scene
lo, hi = int(sys.argv[1]), int(sys.argv[2]); worst = 0.0; nfail = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(83000 + seed)
    n = int(10 ** rng.uniform(1, 4.0))
    nc = int(rng.integers(1, 8)); centres = rng.uniform(-0.8, 0.8, (nc, 3)); radii = 10 ** rng.uniform(-2.5, -0.2, nc)
    k = rng.integers(0, nc, n)
    pos = centres[k] + rng.normal(0, 1, (n, 3)) * radii[k, None]
    if rng.random() < 0.3: pos[:3] = rng.uniform(-1, 1, (3, 3)) * 500      # far outliers
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    gs = O.GaussianSet(pos=pos, rot=q, scale=np.exp(rng.normal(np.log(10 ** rng.uniform(-2.5, -0.8)), 0.6, (n, 3))),
                       color=rng.uniform(0, 1, (n, 3)), opacity=rng.uniform(0.05, 0.95, n), sh=rng.normal(0, 0.15, (n, 15, 3)))
    bits = int(rng.choice([0, 30, 63]))
    scene = Scene(morton_bits=bits if bits else "auto")
    scene.from_arrays(gs.pos, gs.rot, gs.scale, gs.color, gs.opacity, gs.sh)
    W, H = int(rng.integers(16, 160)), int(rng.integers(16, 110))
    pos_c, rot_c = orbit_pose(float(rng.uniform(0, 6.28)), float(rng.uniform(0.3, 2.8)), float(10 ** rng.uniform(-1.3, 0.5)))
    pos_c = np.asarray(pos_c) + centres[int(rng.integers(0, nc))]
    f = focal_from_fov(H, float(rng.uniform(25, 110)))
    cam = Camera(pos_c, rot_c, (W, H), (f, f))
    ocam = O.CameraParams(np.asarray(pos_c), np.asarray(rot_c), W, H, (f, f))
    depth = int(rng.choice([4, 16, 16]))
    ref = O.render(gs, ocam, depth=depth)["rgb"].reshape(W, H, 3)
    rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
    errs = []
    for trial in range(3):
        opts = dict(render_mode=int(rng.choice([0, 0, 2, 1])), list_pool_chunks=int(rng.choice([-1, 0, 8, 64, 512])),
                    heavy_limit=int(rng.choice([-1, 0, 40, 300])), heavy_lists=int(rng.choice([0, 1, 2])))
        for kk, vv in opts.items(): scene.set_option(kk, vv)
        e1 = float(np.abs(rt.render(depth) - ref).max())
        e2 = float(np.abs(rt.render(depth) - ref).max())          # second frame: pool may have grown, slab lists switched on
        errs.append((opts, max(e1, e2)))
    for kk, vv in dict(render_mode=0, list_pool_chunks=-1, heavy_limit=-1, heavy_lists=0).items(): scene.set_option(kk, vv)
    mx = max(e for _, e in errs); worst = max(worst, mx); ok = mx <= 1e-3; nfail += not ok
    print(f"seed {seed}: n={n} bits={bits}->{scene.get_option('morton_bits')} {W}x{H} depth={depth} max err {mx:.1e}" +
          ("" if ok else "  <<<<<< FAIL " + str([o for o, e in errs if e > 1e-3])), flush=True)
print("worst", worst, "failures", nfail)
