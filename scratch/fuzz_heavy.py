"""One-off fuzz sweep of the opt-in depth-slab lists: random scenes / cameras with a random (low) heavy limit, the frame
with heavy_lists = 2 must equal the default route bit for bit and the oracle within tolerance."""
import sys, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import ref_numpy as O
from gpu_util import compare, make_camera, make_scene, random_set
from rtgs.ray_tracer import RayTracer
lo, hi = int(sys.argv[1]), int(sys.argv[2])
worst = 0.0; nfail = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.integers(1, 4000))
    dense = rng.random() < 0.6
    ms = float(rng.uniform(0.05, 0.3)) if dense else float(rng.uniform(0.005, 0.08))
    gs = random_set(n, seed=11000 + seed, mean_scale=ms, sh=bool(rng.integers(0, 2)))
    if rng.random() < 0.3: gs.pos[:, 2] *= 0.02          # a sheet: grazing views
    scene = make_scene(gs)
    W, H = int(rng.integers(9, 200)), int(rng.integers(9, 140))
    depth = int(rng.choice([1, 2, 5, 16, 16, 16]))
    cam, ocam = make_camera(float(rng.uniform(0, 6.28)), float(rng.uniform(0.2, 2.9)), float(rng.uniform(0.1, 4.0)), W, H,
                            fov=float(rng.uniform(20, 120)))
    ref = O.render(gs, ocam, depth=depth)
    rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
    scene.set_option("heavy_limit", int(rng.choice([0, 8, 32, 128, -1])))
    scene.set_option("heavy_lists", 0)
    base = rt.render(depth).copy()
    scene.set_option("heavy_lists", 2)
    img = rt.render(depth).copy()
    rt.render_device(depth, collect_stats=True)
    st = rt.last_stats
    mx, ps, bad = compare(img, ref["rgb"], 1e-3)
    same = np.array_equal(img, base)
    worst = max(worst, mx)
    ok = same and mx <= 1e-3
    nfail += not ok
    print(f"seed {seed}: n={n} ms={ms:.3f} {W}x{H} depth={depth} kbar={np.minimum(ref['nhit'], depth).mean():.2f} heavy tiles {st['heavy_groups']} "
          f"failed {st['heavy_failed']} fallback {st['fallback_tiles']} err {mx:.2e} identical {same}{'' if ok else '  <<<<<< FAIL'}", flush=True)
print("worst", worst, "failures", nfail)
