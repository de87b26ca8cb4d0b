#!/usr/bin/env python
"""Second reference-run golden: a scene that SATURATES the k-buffer and gives the reference's SAH builder a
multi-level tree.  tests/golden/reference_ksat_fp{64,32}.npz are produced by executing the
REFERENCE'S OWN SOURCE (/root/reference/src/rtgs, imported unmodified, through oracle/taichi_shim) on
tests/data/ksat.ply: 220 large translucent Gaussians, `Scene(1024, 4, 16)` exactly as the reference's CLI builds it
(__main__.py:97), 24x24 pixels, depth 16.  Many rays cross far more than 16 ellipsoids, so the restart rule
`ray.start = t1 + 1e-8` (ray_tracer.py:100-102), the strict `start < t1` acceptance (scene.py:433) and the
truncation after 16 layers are all exercised against reference-produced output.

    python tests/golden/make_golden_ksat.py            # writes both precisions (needs /root/reference)

The PLY itself is written by this repo's `rtgs.synthetic.export_ply` in a subprocess (the reference's package
is also called `rtgs`, so the two cannot share a process)."""
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
PLY = ROOT / "tests" / "data" / "ksat.ply"
N, SEED, RES, DEPTH = 220, 4242, 24, 16
CAM = (0.35, 1.25, 2.6)      # orbit theta, phi, r
FOV = 60.0

MAKE_PLY = f"""
import sys, numpy as np
sys.path.insert(0, {str(ROOT / 'rt-gaussian-splat-renderer_b200')!r})
from rtgs.synthetic import export_ply
from rtgs.utils.math import sigmoid
rng = np.random.default_rng({SEED})
f32 = np.float32
n = {N}
pos = rng.uniform(-1, 1, (n, 3)).astype(f32)
q = rng.normal(size=(n, 4)).astype(f32)
rot = (q / np.linalg.norm(q, axis=-1)[:, None]).astype(f32)
scale = np.exp(rng.normal(np.log(0.17), 0.45, (n, 3))).astype(f32)     # big, anisotropic: ~25 crossings per ray
color = sigmoid(rng.normal(0, 1, (n, 3)).astype(f32)).astype(f32)
opacity = rng.uniform(0.04, 0.45, n).astype(f32)                        # translucent: deep layers still count
sh = rng.normal(0, 0.15, (n, 15, 3)).astype(f32)
export_ply({str(PLY)!r}, dict(pos=pos, rot=rot, scale=scale, color=color, opacity=opacity, sh=sh))
"""


def run(fp):
    os.environ["TAICHI_SHIM_FP"] = fp
    sys.path.insert(0, str(ROOT / "oracle" / "taichi_shim"))
    sys.path.insert(0, str(REF / "src"))
    sys.path.insert(0, str(ROOT))
    import taichi as ti  # the shim
    from rtgs.camera import Camera  # REFERENCE modules
    from rtgs.ray_tracer import RayTracer
    from rtgs.scene import Scene
    from rtgs.utils.types import vec2i
    from oracle.ref_numpy import orbit_pose

    np.seterr(all="ignore")
    scene = Scene(1024, 4, 16)                                   # __main__.py:97
    scene.load_file(PLY, 1.0)
    g = scene.gaussian_field
    n = g.shape[0]
    names = ["sh_10", "sh_11", "sh_12", "sh_20", "sh_21", "sh_22", "sh_23", "sh_24",
             "sh_30", "sh_31", "sh_32", "sh_33", "sh_34", "sh_35", "sh_36"]
    out = {}
    out["pos"] = np.array([np.asarray(g[i].position) for i in range(n)], dtype=np.float32)
    out["rot"] = np.array([np.asarray(g[i].rotation) for i in range(n)], dtype=np.float32)
    out["scale"] = np.array([np.asarray(g[i].scale) for i in range(n)], dtype=np.float32)
    out["color"] = np.array([np.asarray(g[i].color) for i in range(n)], dtype=np.float32)
    out["opacity"] = np.array([float(g[i].opacity) for i in range(n)], dtype=np.float32)
    out["sh"] = np.array([[np.asarray(getattr(g[i], nm)) for nm in names] for i in range(n)], dtype=np.float32)
    nodes = scene.bvh_field
    m = nodes.shape[0]
    out["bvh_int"] = np.array([[nodes[i].left, nodes[i].right, nodes[i].prim_left, nodes[i].prim_right,
                                nodes[i].depth] for i in range(m)], dtype=np.int32)
    # __main__.py:91-92; rounded to float32, which is what Taichi's f32 vec2 kernel argument holds (camera.py:17-29)
    focal = float(np.float32((RES / 2) / np.tan(FOV * np.pi / 360)))
    cpos, crot = orbit_pose(*CAM)
    cam = Camera(ti.math.vec3(cpos), ti.math.vec4(crot), vec2i((RES, RES)), ti.math.vec2(focal, focal))
    rt = RayTracer(vec2i((RES, RES)), scene, cam)
    t0 = time.time()
    for _ in range(DEPTH):                                       # __main__.py:253-256: one layer per call
        rt.sample(DEPTH)
    assert rt.num_samples == 1 and rt.num_steps == 0
    rt.generate_disp_buffer(rt.num_samples, rt.num_steps, DEPTH)
    print(f"fp{fp}: reference pipeline {RES}x{RES} depth {DEPTH} on {n} Gaussians: {time.time() - t0:.1f} s, "
          f"BVH depth {out['bvh_int'][:, 4].max()}")
    out.update(cam_pos=cpos, cam_rot=crot, focal=np.float64(focal), res=np.int32(RES), depth=np.int32(DEPTH),
               sample_buf=rt.sample_buf._data.astype(np.float64),
               attenuation=rt.attenuation_buf._data.astype(np.float64),
               disp=rt.disp_buf._data.astype(np.float64))
    dst = Path(__file__).resolve().parent / f"reference_ksat_fp{fp}.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, f"{dst.stat().st_size / 1024:.1f} KiB")


if __name__ == "__main__":
    assert REF.exists(), "/root/reference is required to (re)generate the golden vectors"
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        subprocess.run([sys.executable, "-c", MAKE_PLY], check=True)
        # the shim's precision is fixed at import: one process each (about a quarter of an hour, side by side)
        procs = [subprocess.Popen([sys.executable, __file__, fp]) for fp in ("64", "32")]
        assert all(p.wait() == 0 for p in procs)
