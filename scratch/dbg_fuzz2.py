import sys, numpy as np
sys.path.insert(0, 'scratch')
import importlib.util
spec = importlib.util.spec_from_file_location("fuzz2", "scratch/fuzz2.py")
m = importlib.util.module_from_spec(spec)
try:
    spec.loader.exec_module(m)
except SystemExit:
    pass
from oracle import ref_numpy as O
from gpu_util import compare, make_scene
from rtgs.ray_tracer import RayTracer
seed = int(sys.argv[1]); what = sys.argv[2]
D = m.draw(seed)
gs, cam, ocam, depth, tc = D["gs"], D["cam"], D["ocam"], D["depth"], D["tc"]
print({k: D[k] for k in "n S ms aniso kind W H r fov depth tc".split()}, "cam", cam.position, flush=True)
scene = make_scene(gs)
W, H = D["W"], D["H"]
if what == "A":
    ref = O.render(gs, ocam, depth=depth)
    rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
    for mode, lim, pool in ((0, -1, -1), (0, 0, -1), (0, 1 << 30, -1), (1, -1, -1), (0, -1, 0)):
        scene.set_option("render_mode", mode); scene.set_option("heavy_limit", lim); scene.set_option("list_pool_chunks", pool)
        img = rt.render(depth).copy()
        rt.render_device(depth, collect_stats=True); st = rt.last_stats
        df = np.abs(img.astype(np.float64) - ref["rgb"]).max(axis=-1)
        bad = np.argwhere(df > 1e-3)
        nh = np.asarray(ref["nhit"]).reshape(W, H)
        print(f"mode {mode} heavy_limit {lim} pool {pool}: max {df.max():.2e} bad {len(bad)} fallback {st['fallback_tiles']} f64 {st['f64_refinements']}",
              [(int(a), int(b), float(df[a, b]), int(nh[a, b])) for a, b in bad[:8]], flush=True)
if what in ("B", "C"):
    d16 = min(depth, 16)
    rt2 = RayTracer(cam.buf_size, scene, cam, t_cut=tc)
    imgs = {}
    for mode in (1, 0, 2):
        scene.set_option("render_mode", mode)
        print("render mode", mode, flush=True)
        imgs[mode] = rt2.render(d16).copy()
        rt2.render_device(d16, collect_stats=True); st = rt2.last_stats
        print("   fallback", st["fallback_tiles"], "layers", st["layers"], flush=True)
    rt0 = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
    scene.set_option("render_mode", 1)
    exact = rt0.render(d16).copy()
    for mode in (0, 1, 2):
        df = np.abs(imgs[mode] - exact).max(axis=-1)
        print(f"mode {mode} vs t_cut=0: max {df.max():.2e} (bound 4 t_cut = {4*tc:.1e})")
    df = np.abs(imgs[0] - imgs[1]).max(axis=-1); bad = np.argwhere(df > 1e-5)
    print("mode 0 vs 1:", df.max(), len(bad), [(int(a), int(b), float(df[a, b])) for a, b in bad[:8]])
