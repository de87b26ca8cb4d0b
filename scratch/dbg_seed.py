import sys, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import ref_numpy as O
from gpu_util import compare, make_camera, make_scene, random_set
from rtgs.ray_tracer import RayTracer
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rng = np.random.default_rng(1000 + seed)
n = int(rng.integers(1, 2500))
gs = random_set(n, seed=2000 + seed, mean_scale=float(rng.uniform(0.01, 0.15)), sh=bool(rng.integers(0, 2)))
scene = make_scene(gs)
W, H = int(rng.integers(17, 150)), int(rng.integers(9, 110))
depth = int(rng.choice([1, 3, 8, 16, 16, 16, 24]))
th, ph, r, fov = float(rng.uniform(0, 6.28)), float(rng.uniform(0.3, 2.8)), float(rng.uniform(0.2, 3.5)), float(rng.uniform(30, 110))
cam, ocam = make_camera(th, ph, r, W, H, fov=fov)
print("n", n, "WH", W, H, "depth", depth, "pose", th, ph, r, fov)
ref = O.render(gs, ocam, depth=depth)
rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
for mode in (0, 1):
    scene.set_option("render_mode", mode)
    img = rt.render(depth).copy()
    d = np.abs(img - ref["rgb"]).max(axis=-1)
    bad = np.argwhere(d > 1e-4)
    print("mode", mode, "max", d.max(), "bad px", len(bad), bad[:8].tolist(), "nhit at bad", [int(np.asarray(ref["nhit"]).reshape(W, H)[i, j]) for i, j in bad[:8]])
scene.set_option("render_mode", 0)
rt.render_device(depth, collect_stats=True)
print(rt.last_stats)
