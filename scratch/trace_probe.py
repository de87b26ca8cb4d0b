"""Scene.hit (k_trace_closest) on the 1080p camera rays of the bench scene: the launch the ncu inventory captures."""
import sys, numpy as np, torch
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.')
import bench
from rtgs import _native
from rtgs.camera import Camera
from rtgs.scene import Scene
arrays, n, seed, deg, (W, H), nv, phi, what = bench.load_config("1m_deg3_1080p")
scene = Scene().from_arrays(arrays["pos"], arrays["rot"], arrays["scale"], arrays["color"], arrays["opacity"], arrays["sh"])
f, views = bench.make_views(W, H, nv, phi)
cam = Camera(views[0][0], views[0][1], (W, H), (f, f))
rays = torch.from_numpy(np.ascontiguousarray(cam.cam_ray_field.to_numpy().reshape(-1, 8))).cuda()
idx = torch.empty(rays.shape[0], dtype=torch.int32, device="cuda")
t12 = torch.empty((rays.shape[0], 2), dtype=torch.float32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(3):
    e0.record()
    _native.check(_native.load().rtgs_trace_closest(scene.handle, rays.shape[0], rays.data_ptr(), idx.data_ptr(), t12.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream))
    e1.record(); torch.cuda.synchronize()
print(f"k_trace_closest: {rays.shape[0]} rays in {e0.elapsed_time(e1):.3f} ms = {rays.shape[0] / e0.elapsed_time(e1) / 1e3:.1f} Mrays/s, "
      f"hit fraction {(idx >= 0).float().mean().item():.3f}")
