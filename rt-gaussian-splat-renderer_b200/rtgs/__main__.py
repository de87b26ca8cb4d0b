"""Headless driver — the non-GUI part of the reference's viewer (``rtgs/__main__.py:34-152``).

    python -m rtgs -o scene.ply [-r W,H] [-f FOV] [-s SAMPLES] [-d DEPTH] [-v BVH] [--scale S]
                   [--theta T --phi P --radius R | --views N] [--out DIR] [--format png|npy|both]

The flags ``-o/-r/-f/-s/-d/-v/--scale`` and their defaults are the reference's (``__main__.py:41-85``); the
orbit pose is ``update_camera_pose`` (``:120-142``) with the identity global rotation and cursor 0, start pose
``theta=0, phi=pi/2, r=1`` (``:107-109``).  Instead of opening a ti.GUI window it renders the requested pose
(or an N-view orbit sweep ``theta_k = 2 pi k / N``) with ``RayTracer.sample`` exactly like the viewer's
render loop (``:236-252``: ``sample(depth)`` until ``num_samples`` samples are done, then
``generate_disp_buffer``) and writes the display buffer.  With several GPUs (``--gpus N`` or torchrun) the
views are dealt round-robin to the ranks (``rtgs.sharding.views_for_rank``); nothing is exchanged.
"""
from __future__ import annotations

import argparse
import logging
import os
import pathlib
import sys
import time

import numpy as np

logger = logging.getLogger("rtgs")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser("rtgs", description="B200-native 3D Gaussian ray tracer (headless).")
    # the reference's flags, __main__.py:41-85
    p.add_argument("-o", "--open", type=pathlib.Path, help="Path to the .ply Gaussian splatting scene file.")
    p.add_argument("-r", "--res", type=lambda s: tuple(map(int, s.split(","))), default=(960, 540),
                   help="Render resolution W,H")
    p.add_argument("-f", "--fov", type=float, default=90, help="Vertical FOV in degree.")
    p.add_argument("-s", "--sample", type=int, default=1, help="Render sample rate.")
    p.add_argument("-d", "--depth", type=int, default=16, help="Render sample depth.")
    p.add_argument("-v", "--bvh", type=int, default=1024, help="BVH size (accepted; the LBVH has 2n-1 nodes)")
    p.add_argument("--scale", type=float, default=1, help="Global Gaussian Scale")
    # headless additions
    p.add_argument("--synthetic", default=None, help="render a synthetic bench scene (rtgs.synthetic.CONFIGS name) "
                                                     "instead of --open")
    p.add_argument("--export-ply", type=pathlib.Path, default=None,
                   help="with --synthetic: also write the scene as a 62-property 3DGS .ply (the reference can load it)")
    p.add_argument("--theta", type=float, default=0.0)
    p.add_argument("--phi", type=float, default=float(np.pi / 2))
    p.add_argument("--radius", type=float, default=1.0)
    p.add_argument("--views", type=int, default=0, help="orbit sweep: N views theta_k = 2 pi k / N (0 = single pose)")
    p.add_argument("--out", type=pathlib.Path, default=pathlib.Path("rtgs_out"))
    p.add_argument("--format", choices=("png", "npy", "both"), default="png")
    p.add_argument("--t-cut", type=float, default=1e-4,
                   help="transmittance early-termination threshold (0 = off, the reference's behaviour: it has no "
                        "early-out; 1e-4 changes a pixel by < 1e-4)")
    p.add_argument("--sh-layout", choices=("channel_major", "interleaved"), default="channel_major",
                   help="how f_rest_* map to the SH triples (Scene.load_file): channel_major = the 3DGS file layout "
                        "(default), interleaved = the flat reinterpretation the reference-run goldens contain")
    p.add_argument("--device", type=int, default=None)
    return p


def to_display(disp: np.ndarray) -> np.ndarray:
    """(W,H,3) field layout (origin bottom-left, [i,j] = column,row) -> (H,W,3) uint8, origin top-left; values
    clipped to [0,1] like ti.GUI.set_image does."""
    img = np.clip(np.asarray(disp, np.float32), 0.0, 1.0).transpose(1, 0, 2)[::-1]
    return (img * 255.0 + 0.5).astype(np.uint8)


def write_image(path: pathlib.Path, disp: np.ndarray, fmt: str) -> list[pathlib.Path]:
    out = []
    if fmt in ("npy", "both"):
        np.save(path.with_suffix(".npy"), np.asarray(disp, np.float32))
        out.append(path.with_suffix(".npy"))
    if fmt in ("png", "both"):
        from PIL import Image
        Image.fromarray(to_display(disp)).save(path.with_suffix(".png"))
        out.append(path.with_suffix(".png"))
    return out


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(name)s: %(message)s")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = args.device if args.device is not None else int(os.environ.get("LOCAL_RANK", "0"))

    from .camera import Camera
    from .orbit import focal_from_fov, orbit_pose
    from .ray_tracer import RayTracer
    from .scene import Scene
    from .sharding import views_for_rank
    from .utils.types import vec2i

    res = tuple(args.res)
    focal = focal_from_fov(res[1], args.fov)                    # __main__.py:88-92
    t0 = time.perf_counter()
    scene = Scene(args.bvh, 4, 16, device=device)               # __main__.py:97
    if args.synthetic:
        from .synthetic import CONFIGS, make_scene
        n, seed, deg, _ = CONFIGS[args.synthetic]
        a = make_scene(n, seed, deg)
        if args.export_ply is not None:
            from .synthetic import export_ply
            export_ply(args.export_ply, a)
            logger.info("wrote %s", args.export_ply)
        scene.from_arrays(a["pos"], a["rot"], a["scale"], a["color"], a["opacity"], a["sh"])
    elif args.open is not None:
        scene.load_file(args.open, args.scale, sh_layout=args.sh_layout)
    else:
        print("rtgs: one of -o/--open or --synthetic is required", file=sys.stderr)
        return 2
    logger.info("scene ready: %d Gaussians, LBVH built in %.1f ms", scene.num_gaussians, 1e3 * (time.perf_counter() - t0))

    pos, rot = orbit_pose(args.theta, args.phi, args.radius)
    camera = Camera(pos, rot, vec2i(res), (focal, focal), device=device)
    tracer = RayTracer(vec2i(res), scene, camera, t_cut=args.t_cut)
    poses = [(args.theta, "frame")] if args.views <= 0 else \
        [(2 * np.pi * k / args.views, f"view_{k:03d}") for k in views_for_rank(args.views, rank, world)]
    args.out.mkdir(parents=True, exist_ok=True)
    t0 = time.perf_counter()
    if args.sample == 1 and len(poses) > 1:
        # one sample per view: the display buffer is the sample itself, so the sweep can run through the two-deep
        # pipeline (RayTracer.sweep: view k+1 renders while view k is copied out and written)
        for k, img in tracer.sweep([orbit_pose(theta, args.phi, args.radius) for theta, _ in poses], args.depth):
            for f in write_image(args.out / poses[k][1], img, args.format):
                logger.info("wrote %s", f)
        poses_done, poses = poses, []
    for theta, name in poses:
        camera.position, camera.rotation = orbit_pose(theta, args.phi, args.radius)
        tracer.clear_sample()                                   # __main__.py:214-216 on a camera move
        tracer.num_steps = tracer.num_samples = 0
        while tracer.num_samples < args.sample:                 # __main__.py:236-247
            tracer.sample(args.depth)
        tracer.generate_disp_buffer(tracer.num_samples, tracer.num_steps, args.depth)
        for f in write_image(args.out / name, tracer.disp_buf.to_numpy(), args.format):
            logger.info("wrote %s", f)
    dt = time.perf_counter() - t0
    if args.sample == 1 and not poses:
        poses = poses_done
    if poses:
        logger.info("rank %d: %d view(s) of %dx%d in %.3f s incl. image output", rank, len(poses), res[0], res[1], dt)
    return 0


if __name__ == "__main__":
    sys.exit(main())
