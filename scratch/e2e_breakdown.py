import sys, time, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.')
import torch
from rtgs.synthetic import CONFIGS, make_scene
from rtgs.scene import Scene
from rtgs.camera import Camera
from rtgs.ray_tracer import RayTracer
from rtgs.orbit import focal_from_fov, orbit_pose
n, seed, deg, (W, H) = CONFIGS["1m_deg3_1080p"]
a = make_scene(n, seed, deg)
scene = Scene(device=0).from_arrays(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
f = focal_from_fov(H, 60.0)
views = [orbit_pose(2*np.pi*k/64, np.pi/2, 2.2) for k in range(64)]
cam = Camera(views[0][0], views[0][1], (W, H), (f, f), device=0)
rt = RayTracer((W, H), scene, cam, t_cut=1e-4)
for s in range(5):
    cam.position, cam.rotation = views[s]; rt.render(16)
N = 64
scene.set_option("kernel_timing", N)
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(N):
    cam.position, cam.rotation = views[s]; rt.render(16)
dt = (time.perf_counter() - t0) / N
kt = scene.read_kernel_times(N).mean(axis=0)
print("e2e ms/frame", 1e3*dt, "kernels ms", kt, "sum", kt.sum(), "host+sync overhead", 1e3*dt - kt.sum())
# device output for comparison
out = torch.empty((W, H, 3), device="cuda")
t0 = time.perf_counter()
for s in range(N):
    cam.position, cam.rotation = views[s]; rt.render_device(16, out=out)
torch.cuda.synchronize()
dt2 = (time.perf_counter() - t0) / N
kt2 = scene.read_kernel_times(N).mean(axis=0)
print("device ms/frame", 1e3*dt2, "kernels ms", kt2, "sum", kt2.sum())
