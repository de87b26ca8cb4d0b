import sys, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.')
from oracle import ref_cpu, ref_numpy as O
from rtgs.camera import Camera
from rtgs.orbit import focal_from_fov, orbit_pose
from rtgs.ray_tracer import RayTracer
from rtgs.scene import Scene
from rtgs.synthetic import CONFIGS, FOV_DEG, ORBIT_R, make_scene
n, seed, deg, (W, H) = CONFIGS["1m_deg3_1080p"]
a = make_scene(n, seed, deg)
scene = Scene().from_arrays(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
cs = ref_cpu.CpuScene(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
gs = O.GaussianSet(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
f = focal_from_fov(H, FOV_DEG)
cases = {0: [(1346, 258), (357, 567), (1111, 707)], 13: [(918, 121)], 40: [(1215, 252), (451, 56)]}
for view, pts in cases.items():
    pos, rot = orbit_pose(2 * np.pi * view / 64, np.pi / 2, ORBIT_R)
    cam = Camera(pos, rot, (W, H), (f, f))
    ocam = O.CameraParams(np.asarray(pos), np.asarray(rot), W, H, (f, f))
    rt = RayTracer((W, H), scene, cam, t_cut=0.0)
    imgs = {}
    for mode in (0, 1):
        scene.set_option("render_mode", mode)
        imgs[mode] = rt.render(16).copy()
    scene.set_option("render_mode", 0)
    pix = np.array(pts)
    cpp = cs.render(ocam, 16, pixels=pix, precision="double")
    npy = O.render(gs, ocam, depth=16, pixels=pix, return_layers=True) if 'return_layers' in O.render.__code__.co_varnames else O.render(gs, ocam, depth=16, pixels=pix)
    for k, (i, j) in enumerate(pts):
        print(f"view {view} px ({i},{j}): lists {imgs[0][i,j]} fused {imgs[1][i,j]} cpp64 {cpp['rgb'][k]} numpy64 {npy['rgb'][k]} nlayers cpp {cpp['nlayers'][k]} numpy nhit {npy['nhit'][k]}", flush=True)
