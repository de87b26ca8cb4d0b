"""One-off: EVERY pixel of a full-size bench frame against the float64 C++ oracle."""
import sys, time, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.')
from oracle import ref_cpu, ref_numpy as O
from rtgs.camera import Camera
from rtgs.orbit import focal_from_fov, orbit_pose
from rtgs.ray_tracer import RayTracer
from rtgs.scene import Scene
from rtgs.synthetic import CONFIGS, FOV_DEG, ORBIT_R, make_scene
name = sys.argv[1]; views = [int(v) for v in sys.argv[2:]]
n, seed, deg, (W, H) = CONFIGS[name]
a = make_scene(n, seed, deg)
scene = Scene().from_arrays(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
cs = ref_cpu.CpuScene(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
f = focal_from_fov(H, FOV_DEG)
pix = ref_cpu.all_pixels(W, H, 1)
for view in views:
    pos, rot = orbit_pose(2 * np.pi * view / 64, np.pi / 2, ORBIT_R)
    cam = Camera(pos, rot, (W, H), (f, f))
    rt = RayTracer((W, H), scene, cam, t_cut=0.0)
    img = rt.render(16).copy()
    t0 = time.time()
    ref = cs.render(O.CameraParams(np.asarray(pos), np.asarray(rot), W, H, (f, f)), 16, pixels=pix, precision="double")
    d = np.abs(img[pix[:, 0], pix[:, 1]].astype(np.float64) - ref["rgb"]).max(axis=1)
    print(f"{name} view {view}: {len(pix)} px in {time.time()-t0:.1f}s cpu; max-abs {d.max():.3e}; >1e-4: {(d>1e-4).sum()}; >1e-3: {(d>1e-3).sum()}; "
          f"psnr {O.psnr(img[pix[:,0],pix[:,1]], ref['rgb']):.1f}", flush=True)
    bad = np.argsort(-d)[:3]
    print("   worst px", [(int(pix[b,0]), int(pix[b,1]), float(d[b]), int(ref['nlayers'][b])) for b in bad])
