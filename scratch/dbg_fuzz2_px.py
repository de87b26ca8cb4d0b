import sys, numpy as np
import importlib.util
spec = importlib.util.spec_from_file_location("fuzz2", "scratch/fuzz2.py")
m = importlib.util.module_from_spec(spec)
try: spec.loader.exec_module(m)
except SystemExit: pass
from oracle import ref_numpy as O
from gpu_util import make_scene
from rtgs.ray_tracer import RayTracer
seed = int(sys.argv[1]); pi, pj = int(sys.argv[2]), int(sys.argv[3])
D = m.draw(seed); gs, cam, ocam = D["gs"], D["cam"], D["ocam"]
scene = make_scene(gs)
rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
pix = np.array([[pi, pj]])
o, d = O.camera_rays(ocam, pix)
t1, t2 = O.intersect_all(gs, o[None, :] if o.ndim == 1 else o, d)
t = t1[0]; ok = np.isfinite(t) & (t > 0)
idx = np.argsort(np.where(ok, t, np.inf))[:ok.sum()]
print("pixel", pi, pj, "hits", len(idx))
for k in range(min(22, len(idx))):
    i = idx[k]; gap = (t[idx[k + 1]] - t[i]) / t[i] if k + 1 < len(idx) else 0
    print(f"   layer {k} id {i} t1 {t[i]!r} rel gap to next {gap:.2e} scale {gs.scale[i]} opacity {gs.opacity[i]:.3f}")
for dd in (1, 2, 3, 4, 6, 8, 12, 16):
    ref = O.render(gs, ocam, depth=dd, pixels=pix)
    out = []
    for mode in (0, 1):
        scene.set_option("render_mode", mode)
        img = rt.render(dd)
        out.append(float(np.abs(img[pi, pj] - ref["rgb"][0]).max()))
    print("   depth", dd, "err lists", f"{out[0]:.2e}", "fused", f"{out[1]:.2e}")
