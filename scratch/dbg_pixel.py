import sys, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import ref_numpy as O
from gpu_util import random_set
from rtgs.orbit import focal_from_fov, orbit_pose
seed = 2
rng = np.random.default_rng(1000 + seed)
n = int(rng.integers(1, 2500))
gs = random_set(n, seed=2000 + seed, mean_scale=float(rng.uniform(0.01, 0.15)), sh=bool(rng.integers(0, 2)))
W, H = int(rng.integers(17, 150)), int(rng.integers(9, 110))
depth = int(rng.choice([1, 3, 8, 16, 16, 16, 24]))
th, ph, r, fov = float(rng.uniform(0, 6.28)), float(rng.uniform(0.3, 2.8)), float(rng.uniform(0.2, 3.5)), float(rng.uniform(30, 110))
pos, rot = orbit_pose(th, ph, r); f = focal_from_fov(H, fov)
ocam = O.CameraParams(np.asarray(pos), np.asarray(rot), W, H, (f, f))
pix = np.array([[3, 63]])
o, d = O.camera_rays(ocam, pix)
t1, t2 = O.intersect_all(gs, o if o.ndim == 2 else o[None, :], d)
t = t1[0]; ok = np.isfinite(t) & (t > 0)
idx = np.argsort(np.where(ok, t, np.inf))[:ok.sum()]
ts = t[idx]
print("hits", len(ts))
for k, (i, tt) in enumerate(zip(idx, ts)):
    f32 = np.float32(tt)
    print(k, i, repr(tt), f32.view(np.uint32) if hasattr(f32, 'view') else None, hex(np.float32(tt).view(np.uint32)))
