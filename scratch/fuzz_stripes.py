"""One-off fuzz: stripe-sharded frames (Scene.set_stripe, N 'ranks' on one GPU into one buffer) of clustered scenes, random
sizes and world sizes, all three modes, two streams: bit-identical to the single launch."""
import sys, numpy as np, torch
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import ref_numpy as O
from gpu_util import make_scene
from rtgs.camera import Camera
from rtgs.orbit import focal_from_fov, orbit_pose
from rtgs.ray_tracer import RayTracer
lo, hi = int(sys.argv[1]), int(sys.argv[2]); nfail = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(53000 + seed)
    n = int(10 ** rng.uniform(2, 4.3))
    nc = int(rng.integers(1, 8)); centres = rng.uniform(-0.8, 0.8, (nc, 3)); radii = 10 ** rng.uniform(-2.5, -0.3, nc)
    k = rng.integers(0, nc, n)
    pos = centres[k] + rng.normal(0, 1, (n, 3)) * radii[k, None]
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    gs = O.GaussianSet(pos=pos, rot=q, scale=np.exp(rng.normal(np.log(10 ** rng.uniform(-2.7, -1.0)), 0.5, (n, 3))),
                       color=rng.uniform(0, 1, (n, 3)), opacity=rng.uniform(0.05, 0.95, n), sh=rng.normal(0, 0.15, (n, 15, 3)))
    scene = make_scene(gs)
    W, H = int(rng.integers(33, 400)), int(rng.integers(16, 200))
    r = float(10 ** rng.uniform(-1.3, 0.5)); tgt = centres[int(rng.integers(0, nc))]
    pos_c, rot_c = orbit_pose(float(rng.uniform(0, 6.28)), float(rng.uniform(0.3, 2.8)), r)
    f = focal_from_fov(H, float(rng.uniform(25, 110)))
    cam = Camera(np.asarray(pos_c) + tgt, rot_c, (W, H), (f, f))
    rt = RayTracer(cam.buf_size, scene, cam, t_cut=float(rng.choice([0.0, 1e-4])))
    world = int(rng.choice([2, 3, 5, 8])); depth = int(rng.choice([16, 16, 20]))
    s2 = torch.cuda.Stream()
    res = []
    for mode in (0, 2, 1):
        scene.set_option("render_mode", mode)
        scene.set_stripe()
        ref = rt.render_device(depth).clone()
        buf = torch.full_like(ref, -7.0)
        for rk in range(world):
            scene.set_stripe(world, rk)
            if rk % 2:
                s2.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s2): rt.render_device(depth, out=buf)
                torch.cuda.current_stream().wait_stream(s2)
            else:
                rt.render_device(depth, out=buf)
        scene.set_stripe()
        torch.cuda.synchronize()
        res.append(bool(torch.equal(buf, ref)))
    scene.set_option("render_mode", 0)
    ok = all(res); nfail += not ok
    print(f"seed {seed}: n={n} {W}x{H} world={world} depth={depth} r={r:.3f} identical {res}{'' if ok else '  <<<<<< FAIL'}", flush=True)
print("failures", nfail)
