import sys, numpy as np
import importlib.util
spec = importlib.util.spec_from_file_location("fuzz2", "scratch/fuzz2.py")
m = importlib.util.module_from_spec(spec)
try: spec.loader.exec_module(m)
except SystemExit: pass
from oracle import ref_numpy as O
from gpu_util import make_scene
from rtgs.ray_tracer import RayTracer
seed = 25
D = m.draw(seed); gs, cam, ocam = D["gs"], D["cam"], D["ocam"]
W, H = D["W"], D["H"]
scene = make_scene(gs)
print("tree depth", scene.get_option("tree_depth"), "morton bits", scene.get_option("morton_bits"))
ref = O.render(gs, ocam, depth=16)["rgb"].reshape(W, H, 3)
rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
for region in ((16, 24, 4, 8), (16, 16, 8, 16), (0, 0, 32, 32), (0, 0, W, H), (16, 24, 4, 27), (0, 24, 32, 8)):
    x0, y0, w, h = region
    t = rt.render_device(16, tile=region, collect_stats=True)
    img = t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)
    st = rt.last_stats
    sub = img[x0:x0 + w, y0:y0 + h] if img.shape[0] == W else img
    df = np.abs(sub - ref[x0:x0 + w, y0:y0 + h]).max(axis=-1)
    print(region, "img", img.shape, "max err", float(df.max()), "bad", int((df > 1e-3).sum()), "cands", st["candidates"], "tiles", st["tiles"], "fallback", st["fallback_tiles"],
          "max_group_list", st["max_group_list"], "max stack", st["max_lists_stack"], "nodes", st["nodes_tested"])
