"""EVERY pixel of full-size frames of the surface-like scene (where 12 % of the tiles take the fused kernel) against
the float64 C++ oracle, in the three render modes."""
import sys, time, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.')
import bench
from oracle import ref_cpu, ref_numpy as O
from rtgs.camera import Camera
from rtgs.ray_tracer import RayTracer
from rtgs.scene import Scene
a, n, seed, deg, (W, H), nv, phi, what = bench.load_config("surface_1m_1080p")
scene = Scene().from_arrays(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
cs = ref_cpu.CpuScene(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
f, views = bench.make_views(W, H, nv, phi)
pix = ref_cpu.all_pixels(W, H, 1)
for view in [int(v) for v in sys.argv[1:]] or [0]:
    pos, rot = views[view]
    cam = Camera(pos, rot, (W, H), (f, f))
    rt = RayTracer((W, H), scene, cam, t_cut=0.0)
    t0 = time.time()
    ref = cs.render(O.CameraParams(np.asarray(pos), np.asarray(rot), W, H, (f, f)), 16, pixels=pix, precision="double")
    cpu_s = time.time() - t0
    for mode in (0, 2):
        scene.set_option("render_mode", mode)
        img = rt.render(16).copy()
        d = np.abs(img[pix[:, 0], pix[:, 1]].astype(np.float64) - ref["rgb"]).max(axis=1)
        print(f"surface view {view} mode {mode}: {len(pix)} px ({cpu_s:.0f}s cpu); max-abs {d.max():.3e}; >1e-4: {(d>1e-4).sum()}; >1e-3: {(d>1e-3).sum()}; "
              f"psnr {O.psnr(img[pix[:,0],pix[:,1]], ref['rgb']):.1f}", flush=True)
    scene.set_option("render_mode", 0)
