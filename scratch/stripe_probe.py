"""Per-kernel times of one rank's share of a tile-sharded frame, on a single GPU (rank 0 of world 1/2/4/8), in the
render modes given on the command line: python scratch/stripe_probe.py CONFIG MODE [MODE...]"""
import sys, numpy as np, torch
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.')
from rtgs.camera import Camera
from rtgs.orbit import focal_from_fov, orbit_pose
from rtgs.ray_tracer import RayTracer
from rtgs.scene import Scene
from rtgs.synthetic import CONFIGS, FOV_DEG, ORBIT_R, make_scene
name = sys.argv[1] if len(sys.argv) > 1 else "1m_deg3_1080p"
modes = [int(m) for m in sys.argv[2:]] or [0, 2]
n, seed, deg, (W, H) = CONFIGS[name]
a = make_scene(n, seed, deg)
scene = Scene().from_arrays(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
f = focal_from_fov(H, FOV_DEG)
views = [orbit_pose(2 * np.pi * v / 64, np.pi / 2, ORBIT_R) for v in range(64)]
cam = Camera(views[0][0], views[0][1], (W, H), (f, f))
rt = RayTracer((W, H), scene, cam)
out = torch.empty((W, H, 3), dtype=torch.float32, device="cuda")
for mode in modes:
    scene.set_option("render_mode", mode)
    for world in (1, 2, 4, 8):
        for rank in ((0,) if world == 1 else (0, world - 1)):
            scene.set_stripe(world, rank)
            for v in range(5):
                cam.position, cam.rotation = views[v]; rt.render_device(16, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for v in range(32):
                cam.position, cam.rotation = views[v]; rt.render_device(16, out=out)
            e1.record(); torch.cuda.synchronize()
            frame = e0.elapsed_time(e1) / 32
            # two frames in flight on alternating streams (each with its own frame scratch in the library)
            st = [torch.cuda.Stream(), torch.cuda.Stream()]
            e0.record()
            for x in st: x.wait_event(e0)
            for v in range(32):
                cam.position, cam.rotation = views[v]
                with torch.cuda.stream(st[v & 1]):
                    rt.render_device(16, out=out)
            for x in st: torch.cuda.current_stream().wait_stream(x)
            e1.record(); torch.cuda.synchronize()
            frame2 = e0.elapsed_time(e1) / 32
            scene.set_option("kernel_timing", 32)
            for v in range(32):
                cam.position, cam.rotation = views[v]; rt.render_device(16, out=out)
            t = scene.read_kernel_times(32).astype(np.float64)
            scene.set_option("kernel_timing", 0)
            print(f"{name} mode {mode} world {world} rank {rank}: frame {frame:.4f} ms, two streams {frame2:.4f} ms | " +
                  " ".join(f"{k or '-'} {t[:, i].mean():.4f}" for i, k in enumerate(scene.kernel_names)), flush=True)
