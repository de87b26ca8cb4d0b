"""One-off: a 1 M-Gaussian scene made of clusters of very different densities (dense cores, sheets, a uniform haze), 1080p,
cameras outside and inside clusters: every 2nd pixel in i and j against the float64 C++ oracle, all render routes."""
import sys, time, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.')
from oracle import ref_cpu, ref_numpy as O
from rtgs.camera import Camera
from rtgs.orbit import focal_from_fov, orbit_pose
from rtgs.ray_tracer import RayTracer
from rtgs.scene import Scene
rng = np.random.default_rng(777)
n = 1_000_000
nc = 60
centres = rng.uniform(-0.8, 0.8, (nc, 3)); radii = 10 ** rng.uniform(-2.3, -0.5, nc)
k = rng.integers(0, nc, n)
pos = centres[k] + rng.normal(0, 1, (n, 3)) * radii[k, None] * rng.uniform(0, 1, (n, 1)) ** 2
flat = k % 5 == 0
pos[flat, 2] = centres[k[flat], 2] + 0.002 * rng.normal(size=flat.sum())
haze = rng.random(n) < 0.2
pos[haze] = rng.uniform(-1, 1, (haze.sum(), 3))
q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
scale = np.exp(rng.normal(np.log(0.0026), 0.7, (n, 3)))
f32 = np.float32
a = dict(pos=pos.astype(f32), rot=q.astype(f32), scale=scale.astype(f32), color=(1 / (1 + np.exp(-rng.normal(0, 1, (n, 3))))).astype(f32),
         opacity=(1 / (1 + np.exp(-rng.normal(0, 1.5, n)))).astype(f32), sh=rng.normal(0, 0.15, (n, 15, 3)).astype(f32))
scene = Scene().from_arrays(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
print("tree depth", scene.get_option("tree_depth"), "morton bits", scene.get_option("morton_bits"), flush=True)
cs = ref_cpu.CpuScene(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
W, H = 1920, 1080
f = focal_from_fov(H, 60.0)
pix = ref_cpu.all_pixels(W, H, 2)
for view, (th, ph, r, tgt) in enumerate(((0.3, 1.4, 2.2, None), (2.0, 1.0, 0.25, 7), (4.0, 1.9, 0.05, 12), (5.0, 1.5, 0.6, 30))):
    pos_c, rot_c = orbit_pose(th, ph, r)
    pos_c = np.asarray(pos_c, np.float64) + (0 if tgt is None else centres[tgt])
    cam = Camera(pos_c, rot_c, (W, H), (f, f))
    rt = RayTracer((W, H), scene, cam, t_cut=0.0)
    t0 = time.time()
    ref = cs.render(O.CameraParams(np.asarray(pos_c), np.asarray(rot_c), W, H, (f, f)), 16, pixels=pix, precision="double")
    tcpu = time.time() - t0
    for mode in (0, 2, 1):
        scene.set_option("render_mode", mode)
        img = rt.render(16)
        rt.render_device(16, collect_stats=True); st = rt.last_stats
        d = np.abs(img[pix[:, 0], pix[:, 1]].astype(np.float64) - ref["rgb"]).max(axis=1)
        print(f"view {view} (r={r}, cluster {tgt}) mode {mode}: {len(pix)} px ({tcpu:.0f}s cpu) max-abs {d.max():.3e} >1e-4: {(d>1e-4).sum()} >1e-3: {(d>1e-3).sum()} "
              f"kbar {st['layers']/max(st['rays'],1):.2f} cands/tile {st['candidates']/max(st['tiles'],1):.0f} fallback {st['fallback_tiles']} max list {st['max_group_list']}", flush=True)
    scene.set_option("render_mode", 0)
    import torch
    out = torch.empty((W, H, 3), dtype=torch.float32, device="cuda")
    for hl in (0, 2, -1):
        scene.set_option("heavy_lists", max(hl, 0))
        if hl < 0: rt = RayTracer((W, H), scene, cam, t_cut=1e-4)     # the default transmittance cut of bench.py / the CLI
        for _ in range(3): rt.render_device(16, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): rt.render_device(16, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"   heavy_lists {max(hl, 0)} t_cut {rt.t_cut}: {ms:.3f} ms per frame = {W*H/ms/1e3:.0f} Mrays/s", flush=True)
    scene.set_option("heavy_lists", 0)
