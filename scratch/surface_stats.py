"""What the fused kernel does on the heavy tiles of the surface-like scene (RTGS_STATS_FALLBACK_ONLY=1: the
statistics of a render are those of the fallback k_render launch alone)."""
import os, sys, numpy as np, torch
os.environ["RTGS_STATS_FALLBACK_ONLY"] = "1"
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.')
import bench
from rtgs.camera import Camera
from rtgs.ray_tracer import RayTracer
from rtgs.scene import Scene
arrays, n, seed, deg, (W, H), nv, phi, what = bench.load_config("surface_1m_1080p")
scene = Scene().from_arrays(arrays["pos"], arrays["rot"], arrays["scale"], arrays["color"], arrays["opacity"], arrays["sh"])
f, views = bench.make_views(W, H, nv, phi)
cam = Camera(views[0][0], views[0][1], (W, H), (f, f))
for t_cut in (1e-4, 0.0):
    rt = RayTracer((W, H), scene, cam, t_cut=t_cut)
    out = torch.empty((W, H, 3), dtype=torch.float32, device="cuda")
    for v in (0, 5, 11):
        cam.position, cam.rotation = views[v]
        rt.render_device(16, out=out, collect_stats=True)
        st = rt.last_stats
        tl = max(st["tiles"], 1)
        print(f"t_cut {t_cut} view {v}: fallback tiles {st['tiles']}  per tile: boxes {st['nodes_tested']/tl:.0f} steps {st['traversal_steps']/tl:.1f} "
              f"cands {st['candidates']/tl:.0f} rounds {st['insert_rounds']/tl:.1f} f64 {st['f64_refinements']/tl:.1f}; "
              f"per ray: layers {st['layers']/max(st['rays'],1):.2f} hit {st['rays_hit']/max(st['rays'],1):.3f}", flush=True)
