import sys, numpy as np
sys.path.insert(0, 'rt-gaussian-splat-renderer_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import importlib.util
spec = importlib.util.spec_from_file_location("fuzz2", "scratch/fuzz2.py")
m = importlib.util.module_from_spec(spec)
try: spec.loader.exec_module(m)
except SystemExit: pass
except Exception as e: print("import:", e)
from oracle import ref_numpy as O
seed = 25
D = m.draw(seed); gs, ocam = D["gs"], D["ocam"]
W, H = D["W"], D["H"]
i0, j0 = 16, 24
pix = np.array([(i, j) for i in range(i0, i0 + 4) for j in range(j0, j0 + 8)])
o, d = O.camera_rays(ocam, pix)
t1, t2 = O.intersect_all(gs, o[None, :] if o.ndim == 1 else o, d)
hit = np.isfinite(t1) & (t1 > 0)
hit_any = hit.any(0)
print("Gaussians hit by the tile's rays:", hit_any.sum())
# the filter in float64 maths (exact support function) and with fp16 correlations
cov = O.covariance(gs.rot, gs.scale)
p = gs.pos.astype(np.float64)
h = np.sqrt(3 * np.stack([cov[:, 0, 0], cov[:, 1, 1], cov[:, 2, 2]], 1))
rho2 = 2 * np.stack([cov[:, 0, 1] / np.sqrt(cov[:, 0, 0] * cov[:, 1, 1]), cov[:, 0, 2] / np.sqrt(cov[:, 0, 0] * cov[:, 2, 2]),
                     cov[:, 1, 2] / np.sqrt(cov[:, 1, 1] * cov[:, 2, 2])], 1)
rho2h = rho2.astype(np.float16).astype(np.float64)
q = ocam.rotation.astype(np.float64)
fx, fy = float(ocam.focal[0]), float(ocam.focal[1])
def plane(axis, c):
    n = np.array([1.0, 0.0, c]) if axis == 0 else np.array([0.0, 1.0, c])
    return O.rot_vec3(q[None, :], n[None, :])[0]
def side(n, r2):
    dist = (p - o) @ n
    u = n[None, :] * h
    s2 = 1.003 * (u ** 2).sum(1) + u[:, 0] * (u[:, 1] * r2[:, 0] + u[:, 2] * r2[:, 1]) + u[:, 1] * u[:, 2] * r2[:, 2]
    reach = dist ** 2 <= s2
    return (dist >= 0) | reach, (dist <= 0) | reach, dist, s2
for r2, name in ((rho2, "exact rho"), (rho2h, "fp16 rho")):
    nl = plane(0, (i0 - 0.5 * W) / fx); nh = plane(0, (i0 + 4 - 0.5 * W) / fx)
    ml = plane(1, (j0 - 0.5 * H) / fy); mh = plane(1, (j0 + 8 - 0.5 * H) / fy)
    lo_i, _, d1, s1 = side(nl, r2); _, hi_i, d2, s2_ = side(nh, r2)
    lo_j, _, d3, s3 = side(ml, r2); _, hi_j, d4, s4 = side(mh, r2)
    passed = lo_i & hi_i & lo_j & hi_j
    missed = np.nonzero(hit_any & ~passed)[0]
    print(name, "filter passes", passed.sum(), "missed hits:", missed)
    for g in missed:
        print("   g", g, "pos", p[g], "scale", gs.scale[g], "h", h[g], "rho2", r2[g], "lo_i hi_i lo_j hi_j", lo_i[g], hi_i[g], lo_j[g], hi_j[g],
              "dists", d1[g], d2[g], d3[g], d4[g], "s", np.sqrt(s1[g]), np.sqrt(s2_[g]), np.sqrt(s3[g]), np.sqrt(s4[g]), "rays hit", hit[:, g].sum())
