#!/usr/bin/env python
"""Generate tests/golden/reference_golden_fp{64,32}.npz by executing the REFERENCE'S OWN SOURCE
(/root/reference/src/rtgs/*.py, imported unmodified) through oracle/taichi_shim — a pure-Python
stand-in for Taichi, which cannot be installed in the build container.

    TAICHI_SHIM_FP=64 python tests/golden/make_golden.py      # float64 evaluation (parity truth)
    TAICHI_SHIM_FP=32 python tests/golden/make_golden.py      # float32, Taichi's default precision

Runs only where /root/reference exists (the build container); the .npz files are committed and are
what the tests read.  Contents:
  kat_*    known-answer vectors for the per-function maths (quaternion, Gaussian.cov / bounding_box /
           hit / eval / eval_sh, Bound.hit, Camera.generate_ray_field)
  pipe_*   the full reference pipeline on tests/data/test.ply: Scene.load_file (PLY read, activations,
           the reference's binned-SAH BVH build), Camera, RayTracer.sample x depth, generate_disp_buffer
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT / "oracle" / "taichi_shim"))
sys.path.insert(0, str(REF / "src"))

import taichi as ti  # noqa: E402  (the shim)

from rtgs.bounding_box import Bound  # noqa: E402  (REFERENCE modules)
from rtgs.camera import Camera  # noqa: E402
from rtgs.gaussian import Gaussian, new_gaussian  # noqa: E402
from rtgs.ray import Ray, new_ray  # noqa: E402
from rtgs.ray_tracer import RayTracer  # noqa: E402
from rtgs.scene import Scene  # noqa: E402
from rtgs.utils import quaternion as quat  # noqa: E402
from rtgs.utils.types import vec2i  # noqa: E402

FP = os.environ.get("TAICHI_SHIM_FP", "64")
PIPE_RES = int(os.environ.get("GOLDEN_RES", "40"))
PIPE_SCALE = 30.0
DEPTH = 16


def orbit_pose(theta, phi, r):
    """__main__.py:120-142 (numpy-quaternion is not installed; the pose is an INPUT recorded in the file)."""
    sys.path.insert(0, str(ROOT))
    from oracle.ref_numpy import orbit_pose as op
    return op(theta, phi, r)


def kats(out):
    rng = np.random.default_rng(2024)
    n = 24
    q = rng.normal(size=(n, 4)).astype(np.float32)
    q[: n // 2] /= np.linalg.norm(q[: n // 2], axis=1, keepdims=True)   # half unit, half arbitrary
    p = rng.normal(size=(n, 4)).astype(np.float32)
    v = rng.normal(size=(n, 3)).astype(np.float32)
    out["kat_q"], out["kat_p"], out["kat_v"] = q, p, v
    out["kat_mul"] = np.array([quat.mul(ti.math.vec4(p[i]), ti.math.vec4(q[i])) for i in range(n)])
    out["kat_conj"] = np.array([quat.conj(ti.math.vec4(q[i])) for i in range(n)])
    out["kat_rot"] = np.array([quat.rot_vec3(ti.math.vec4(q[i]), ti.math.vec3(v[i])) for i in range(n)])
    out["kat_mat3"] = np.array([quat.as_rotation_mat3(ti.math.vec4(q[i])) for i in range(n)])

    # Gaussians (unit quaternions as the loader stores them) and rays
    pos = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    rot = rng.normal(size=(n, 4)).astype(np.float32)
    rot = (rot / np.linalg.norm(rot, axis=1, keepdims=True)).astype(np.float32)
    sca = np.exp(rng.normal(-1.5, 0.5, (n, 3))).astype(np.float32)
    col = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    opa = rng.uniform(0.1, 1, n).astype(np.float32)
    sh = rng.normal(0, 0.15, (n, 15, 3)).astype(np.float32)
    ro = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    rd = (pos - ro + rng.normal(0, 0.15, (n, 3))).astype(np.float32)     # aimed near the Gaussian
    rd = (rd / np.linalg.norm(rd, axis=1, keepdims=True)).astype(np.float32)
    names = ["sh_10", "sh_11", "sh_12", "sh_20", "sh_21", "sh_22", "sh_23", "sh_24",
             "sh_30", "sh_31", "sh_32", "sh_33", "sh_34", "sh_35", "sh_36"]
    cov, bmin, bmax, hit, ev, esh = [], [], [], [], [], []
    for i in range(n):
        g = new_gaussian(ti.math.vec3(pos[i]), ti.math.vec4(rot[i]), ti.math.vec3(sca[i]), ti.math.vec3(col[i]),
                         float(opa[i]))
        for k, nm in enumerate(names):
            setattr(g, nm, ti.math.vec3(sh[i, k]))
        ray = new_ray(ti.math.vec3(ro[i]), ti.math.vec3(rd[i]))
        cov.append(np.asarray(g.cov()))
        bb = g.bounding_box()
        bmin.append(np.asarray(bb.p_min))
        bmax.append(np.asarray(bb.p_max))
        t = g.hit(ray)
        hit.append(np.asarray(t))
        tm = (t.x + t.y) / 2 if np.isfinite(t.x) and np.isfinite(t.y) else 1.0
        ev.append(np.asarray(g.eval(ray.get(tm), ray.direction)))
        esh.append(np.asarray(g.eval_sh(ti.math.normalize(ray.direction))))
    out.update(kat_g_pos=pos, kat_g_rot=rot, kat_g_scale=sca, kat_g_color=col, kat_g_opacity=opa, kat_g_sh=sh,
               kat_ray_o=ro, kat_ray_d=rd, kat_cov=np.array(cov), kat_bmin=np.array(bmin), kat_bmax=np.array(bmax),
               kat_hit=np.array(hit), kat_eval=np.array(ev), kat_eval_sh=np.array(esh))

    # Bound.hit
    lo = rng.uniform(-1, 0, (n, 3)).astype(np.float32)
    hi = (lo + rng.uniform(0.1, 1.5, (n, 3))).astype(np.float32)
    bh = []
    for i in range(n):
        bh.append(np.asarray(Bound(ti.math.vec3(lo[i]), ti.math.vec3(hi[i])).hit(
            new_ray(ti.math.vec3(ro[i]), ti.math.vec3(rd[i])))))
    out.update(kat_box_lo=lo, kat_box_hi=hi, kat_box_hit=np.array(bh))

    # Camera.generate_ray_field
    cpos, crot = orbit_pose(0.7, 1.1, 2.2)
    cam = Camera(ti.math.vec3(cpos), ti.math.vec4(crot), vec2i((7, 5)), ti.math.vec2(6.0, 6.5))
    cam.generate_ray_field(cam.position, cam.rotation)
    rays = np.zeros((7, 5, 8))
    for i in range(7):
        for j in range(5):
            r = cam.cam_ray_field[i, j]
            rays[i, j, :3], rays[i, j, 3:6], rays[i, j, 6], rays[i, j, 7] = r.origin, r.direction, r.start, r.end
    out.update(kat_cam_pos=cpos, kat_cam_rot=crot, kat_cam_rays=rays)


def pipeline(out):
    res = PIPE_RES
    scene = Scene(128, 1, 8)                                    # the reference's defaults (scene.py:78-82)
    scene.load_file(ROOT / "tests" / "data" / "test.ply", PIPE_SCALE)
    n = scene.gaussian_field.shape[0]
    g = scene.gaussian_field
    names = ["sh_10", "sh_11", "sh_12", "sh_20", "sh_21", "sh_22", "sh_23", "sh_24",
             "sh_30", "sh_31", "sh_32", "sh_33", "sh_34", "sh_35", "sh_36"]
    out["pipe_pos"] = np.array([np.asarray(g[i].position) for i in range(n)], dtype=np.float32)
    out["pipe_rot"] = np.array([np.asarray(g[i].rotation) for i in range(n)], dtype=np.float32)
    out["pipe_scale"] = np.array([np.asarray(g[i].scale) for i in range(n)], dtype=np.float32)
    out["pipe_color"] = np.array([np.asarray(g[i].color) for i in range(n)], dtype=np.float32)
    out["pipe_opacity"] = np.array([float(g[i].opacity) for i in range(n)], dtype=np.float32)
    out["pipe_sh"] = np.array([[np.asarray(getattr(g[i], nm)) for nm in names] for i in range(n)], dtype=np.float32)
    nodes = scene.bvh_field
    m = nodes.shape[0]
    out["pipe_bvh_min"] = np.array([np.asarray(nodes[i].bound.p_min) for i in range(m)], dtype=np.float64)
    out["pipe_bvh_max"] = np.array([np.asarray(nodes[i].bound.p_max) for i in range(m)], dtype=np.float64)
    out["pipe_bvh_int"] = np.array([[nodes[i].left, nodes[i].right, nodes[i].prim_left, nodes[i].prim_right,
                                     nodes[i].depth] for i in range(m)], dtype=np.int32)
    half_angle = (90.0 * np.pi) / 360                            # __main__.py:91-92, fov 90
    focal = (res / 2) / np.tan(half_angle)
    cpos, crot = orbit_pose(0.0, np.pi / 2, 1.0)                 # the CLI's initial pose
    cam = Camera(ti.math.vec3(cpos), ti.math.vec4(crot), vec2i((res, res)), ti.math.vec2(focal, focal))
    rt = RayTracer(vec2i((res, res)), scene, cam)
    t0 = time.time()
    for _ in range(DEPTH):                                       # __main__.py:253-256: one layer per call
        rt.sample(DEPTH)
    assert rt.num_samples == 1 and rt.num_steps == 0
    rt.generate_disp_buffer(rt.num_samples, rt.num_steps, DEPTH)
    print(f"reference pipeline {res}x{res} depth {DEPTH}: {time.time() - t0:.1f} s")
    out.update(pipe_cam_pos=cpos, pipe_cam_rot=crot, pipe_focal=np.float64(focal), pipe_res=np.int32(res),
               pipe_scale_arg=np.float64(PIPE_SCALE), pipe_depth=np.int32(DEPTH),
               pipe_sample_buf=rt.sample_buf._data.astype(np.float64),
               pipe_attenuation=rt.attenuation_buf._data.astype(np.float64),
               pipe_disp=rt.disp_buf._data.astype(np.float64))


def main():
    assert REF.exists(), "/root/reference is required to (re)generate the golden vectors"
    np.seterr(all="ignore")
    out = {}
    kats(out)
    pipeline(out)
    dst = Path(__file__).resolve().parent / f"reference_golden_fp{FP}.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, f"{dst.stat().st_size / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
