import sys, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.')
from oracle import ref_cpu, ref_numpy as O
from rtgs.camera import Camera
from rtgs.orbit import focal_from_fov, orbit_pose
from rtgs.ray_tracer import RayTracer
from rtgs.scene import Scene
from rtgs.synthetic import CONFIGS, FOV_DEG, ORBIT_R, make_scene
n, seed, deg, (W, H) = CONFIGS["3m_deg3_2160p"]
a = make_scene(n, seed, deg)
scene = Scene().from_arrays(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
cs = ref_cpu.CpuScene(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
gs = O.GaussianSet(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
f = focal_from_fov(H, FOV_DEG)
view = 5
pts = [(1884, 871), (1994, 1520)]
pos, rot = orbit_pose(2 * np.pi * view / 64, np.pi / 2, ORBIT_R)
cam = Camera(pos, rot, (W, H), (f, f))
ocam = O.CameraParams(np.asarray(pos), np.asarray(rot), W, H, (f, f))
rt = RayTracer((W, H), scene, cam, t_cut=0.0)
imgs = {}
for mode in (0, 1):
    scene.set_option("render_mode", mode)
    imgs[mode] = rt.render(16).copy()
pix = np.array(pts)
cpp = cs.render(ocam, 16, pixels=pix, precision="double")
brute = cs.render(ocam, 16, pixels=pix, precision="double", brute=True)
o, d = O.camera_rays(ocam, pix)
for k, (i, j) in enumerate(pts):
    t1, t2 = O.intersect_all(gs, o[k:k+1] if o.ndim == 2 else o[None, :], d[k:k+1])
    t = t1[0]; ok = np.isfinite(t) & (t > 0)
    ts = np.sort(t[ok])
    print(f"px ({i},{j}): lists {imgs[0][i,j]} fused {imgs[1][i,j]} cpp64 {cpp['rgb'][k]} brute64 {brute['rgb'][k]} nlayers {cpp['nlayers'][k]} nhit {brute['nhit'][k]}")
    print("   sorted t (first 20):", [float(x) for x in ts[:20]])
    print("   gaps:", [float(x) for x in np.diff(ts[:20])])
