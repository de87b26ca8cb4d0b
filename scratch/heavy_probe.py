"""Surface-like scene: the frame with heavy groups rendered by k_render (heavy_lists 0) vs listed in depth slabs (2):
per-kernel times, statistics of the slab walk, bit-identity."""
import os, sys, numpy as np, torch
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.')
import bench
from rtgs.camera import Camera
from rtgs.ray_tracer import RayTracer
from rtgs.scene import Scene
arrays, n, seed, deg, (W, H), nv, phi, what = bench.load_config("surface_1m_1080p")
scene = Scene().from_arrays(arrays["pos"], arrays["rot"], arrays["scale"], arrays["color"], arrays["opacity"], arrays["sh"])
f, views = bench.make_views(W, H, nv, phi)
cam = Camera(views[0][0], views[0][1], (W, H), (f, f))
rt = RayTracer((W, H), scene, cam, t_cut=1e-4)
out = torch.empty((W, H, 3), dtype=torch.float32, device="cuda")
ref = {}
for mode in (0, 2):
    scene.set_option("heavy_lists", mode)
    for v in (0, 5, 11):
        cam.position, cam.rotation = views[v]
        rt.render_device(16, out=out, collect_stats=True)
        st = rt.last_stats
        img = out.cpu().numpy().copy()
        if mode == 0: ref[v] = img
        same = "" if mode == 0 else f" identical={np.array_equal(img, ref[v])} maxdiff={np.abs(img - ref[v]).max():.2e}"
        print(f"heavy_lists {mode} view {v}: heavy tiles {st['heavy_groups']} failed {st['heavy_failed']} passes/group "
              f"{st['heavy_passes']/max(st['heavy_groups'],1):.1f} sample tests/group {st['heavy_sample_tests']/max(st['heavy_groups'],1):.0f} "
              f"kcyc/tile walk {st['heavy_cycles_walk']/max(st['heavy_groups'],1)/1e3:.1f} test {st['heavy_cycles_test']/max(st['heavy_groups'],1)/1e3:.1f} pub {st['heavy_cycles_publish']/max(st['heavy_groups'],1)/1e3:.1f} max deferred {st['max_deferred']} retries {st['heavy_retries']} fail list/defer/passes {st['heavy_failed_list']}/{st['heavy_failed_deferred']}/{st['heavy_failed_passes']} steps {st['traversal_steps']} nodes {st['nodes_tested']} fallback tiles {st['fallback_tiles']} cands/tile {st['candidates']/max(st['tiles'],1):.0f}{same}", flush=True)
    # timing: 16-view orbit, 3 rounds, per-kernel events
    scene.set_option("kernel_timing", 64)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rnd in range(3):
        torch.cuda.synchronize()
        ev0.record()
        for v in range(nv):
            cam.position, cam.rotation = views[v]
            rt.render_device(16, out=out)
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / nv
    kt = scene.read_kernel_times(3 * nv).mean(0).round(4).tolist()
    print(f"heavy_lists {mode}: {ms:.3f} ms per frame = {W*H/ms/1e3:.0f} Mrays/s; kernels {kt}", flush=True)
    scene.set_option("kernel_timing", 0)
