import sys, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import ref_numpy as O
from gpu_util import compare, make_camera, make_scene, random_set
from rtgs.ray_tracer import RayTracer
seed = 20
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
print("camera pos", cam.position, "sh", gs.sh is not None)
rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
scene.set_option("render_mode", 1)
for (pi, pj) in ((140, 51), (138, 68)):
    pix = np.array([[pi, pj]])
    o, d = O.camera_rays(ocam, pix)
    t1, t2 = O.intersect_all(gs, o[None, :] if o.ndim == 1 else o, d)
    t = t1[0]; ok = np.isfinite(t) & (t > 0)
    idx = np.argsort(np.where(ok, t, np.inf))[:ok.sum()]
    print("pixel", pi, pj, "hits", len(idx), "inside (t1<0<t2):", int((np.isfinite(t1[0]) & (t1[0] <= 0) & (t2[0] > 0)).sum()))
    for k in range(8):
        i = idx[k]; print("   layer", k, "id", i, "t1", repr(t[i]), "t2", t2[0][i], "scale", gs.scale[i], "opacity", gs.opacity[i])
    near0 = np.argsort(np.abs(t1[0]))[:5]
    print("   |t1| smallest:", [(int(i), float(t1[0][i]), float(t2[0][i])) for i in near0])
    for dd in (1, 2, 3, 4):
        ref = O.render(gs, ocam, depth=dd, pixels=pix)
        img = rt.render(dd)
        print("   depth", dd, "cuda", img[pi, pj], "oracle", ref["rgb"][0], "diff", np.abs(img[pi, pj] - ref["rgb"][0]).max())
