import sys, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import ref_numpy as O
from gpu_util import compare, make_camera, make_scene, random_set
from rtgs.ray_tracer import RayTracer
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(9000 + seed)
n = int(rng.integers(1, 4000))
dense = rng.random() < 0.6
ms = float(rng.uniform(0.05, 0.3)) if dense else float(rng.uniform(0.005, 0.08))
gs = random_set(n, seed=11000 + seed, mean_scale=ms, sh=bool(rng.integers(0, 2)))
if rng.random() < 0.3: gs.pos[:, 2] *= 0.02
scene = make_scene(gs)
W, H = int(rng.integers(9, 200)), int(rng.integers(9, 140))
depth = int(rng.choice([1, 2, 5, 16, 16, 16]))
cam, ocam = make_camera(float(rng.uniform(0, 6.28)), float(rng.uniform(0.2, 2.9)), float(rng.uniform(0.1, 4.0)), W, H,
                        fov=float(rng.uniform(20, 120)))
hl = int(rng.choice([0, 8, 32, 128, -1]))
print("n", n, "ms", ms, W, H, "depth", depth, "heavy_limit", hl)
for t_cut in (0.0,):
    for d in (depth, 16, 3, 8):
        ref = O.render(gs, ocam, depth=d)
        rt = RayTracer(cam.buf_size, scene, cam, t_cut=t_cut)
        for mode, lim in ((0, -1), (0, hl), (1, -1), (2, -1)):
            scene.set_option("render_mode", mode); scene.set_option("heavy_limit", lim); scene.set_option("heavy_lists", 0)
            img = rt.render(d).copy()
            rt.render_device(d, collect_stats=True)
            st = rt.last_stats
            df = np.abs(img.astype(np.float64) - ref["rgb"]).max(axis=-1)
            bad = np.argwhere(df > 1e-3)
            print(f"depth {d} mode {mode} limit {lim}: max {df.max():.2e} bad px {len(bad)} fallback tiles {st['fallback_tiles']}", [tuple(b) + (float(df[tuple(b)]), int(np.asarray(ref['nhit']).reshape(W, H)[tuple(b)])) for b in bad[:6]])
