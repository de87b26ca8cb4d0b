"""Offline feasibility check for depth-capped candidate lists on the surface-like scene: for groups whose frustum holds
more than 896 boxes, how many candidates lie nearer (by depth along the group's centre direction) than the farthest
16th hit of the group's rays?"""
import sys, numpy as np
sys.path.insert(0, "rt-gaussian-splat-renderer_b200"); sys.path.insert(0, ".")
from rtgs.synthetic import make_surface_scene, SURFACE_CONFIGS, ORBIT_R, FOV_DEG
from oracle import ref_numpy as rn

n, seed, (W, H), views, phi = SURFACE_CONFIGS["surface_1m_1080p"]
sc = make_surface_scene(n, seed)
view = int(sys.argv[1]) if len(sys.argv) > 1 else 0
pos, rot = rn.orbit_pose(2 * np.pi * view / views, phi, ORBIT_R)
f = rn.focal_from_fov(H, FOV_DEG)
cam = rn.CameraParams(pos, rot, W, H, (f, f))
lo, hi = rn.bounding_box_tight(sc["pos"], sc["rot"], sc["scale"])
o = cam.position.astype(np.float64)
GI, GJ = 8, 16
rng = np.random.default_rng(1)

def dirs_of(pix):
    return rn.camera_rays(cam, pix)[1]

def edge_dir(i, j):   # direction through pixel-corner coordinates (i, j) (not centres)
    fx = float(cam.focal[0])
    dc = np.array([(i - 0.5 * W) / fx, (j - 0.5 * H) / fx, -1.0])
    return rn.rot_vec3(cam.rotation.astype(np.float64)[None, :], dc[None, :])[0]

def frustum_cands(i0, i1, j0, j1):
    c00, c10, c01, c11 = edge_dir(i0, j0), edge_dir(i1, j0), edge_dir(i0, j1), edge_dir(i1, j1)
    cen = edge_dir(0.5 * (i0 + i1), 0.5 * (j0 + j1))
    planes = [np.cross(c00, c01), np.cross(c11, c10), np.cross(c10, c00), np.cross(c01, c11)]
    m = np.ones(n, bool)
    for pn in planes:
        if pn @ cen < 0: pn = -pn       # inside = positive
        # farthest box corner along pn
        far = np.where(pn > 0, hi, lo)
        m &= ((far - o) @ pn) >= 0
    return np.nonzero(m)[0], cen / np.linalg.norm(cen)

res = []
extra = []
gi_list = rng.integers(0, W // GI, 400); gj_list = rng.integers(0, (H + GJ - 1) // GJ, 400)
heavy = 0
for gi, gj in zip(gi_list, gj_list):
    i0, j0 = gi * GI, gj * GJ
    idx, c = frustum_cands(i0, i0 + GI, j0, min(j0 + GJ, H))
    if len(idx) <= 896: continue
    heavy += 1
    pix = np.array([(i, j) for i in range(i0, i0 + GI) for j in range(j0, min(j0 + GJ, H))])
    d = dirs_of(pix)
    gs = rn.GaussianSet(sc["pos"][idx], sc["rot"][idx], sc["scale"][idx], sc["color"][idx], sc["opacity"][idx])
    t1, t2 = rn.intersect_all(gs, o[None, :], d)
    t1 = np.where(t1 > 0, t1, np.inf)
    ts = np.sort(t1, axis=1)
    t16 = ts[:, 15]
    nhits = np.isfinite(t1).sum(1)
    zn = np.minimum((lo[idx] - o) * c, (hi[idx] - o) * c).sum(1)
    cap = t16.max()
    # per tile caps
    tile_counts = []
    for ti in range(2):
        for tj in range(2):
            sel = (pix[:, 0] - i0) // 4 == ti
            sel &= (pix[:, 1] - j0) // 8 == tj
            if sel.any(): tile_counts.append(t16[sel].max())
    lat = ((pix[:, 0] - i0) % 2 == 0) & ((pix[:, 1] - j0) % 2 == ((pix[:, 0] - i0) // 2) % 2)
    cap_s = t16[lat].max()
    extra.append((cap / cap_s, int((zn < cap_s * 1.015).sum()), int((zn < cap_s * 1.03).sum()), int(lat.sum())))
    hit_any = np.isfinite(t1).any(0)
    res.append((len(idx), int((zn < cap).sum()) if np.isfinite(cap) else -1, int(hit_any.sum()), int(nhits.min()),
                float(np.median(nhits)), float(ts[:, 0].min()), float(cap),
                int((hit_any & (zn < cap)).sum()) if np.isfinite(cap) else -1))
print("sampled 400 groups, heavy", heavy)
r = np.array([x[:4] for x in res])
fin = r[:, 1] >= 0
print("finite cap:", fin.mean(), "slab<=896 among all heavy:", ((r[:, 1] >= 0) & (r[:, 1] <= 896)).mean())
print("median frustum", np.median(r[:, 0]), "median slab", np.median(r[fin, 1]))

e = np.array(extra)
print("lattice rays", e[0, 3], "cap_true/cap_sample: max", e[:, 0].max(), "frac <=1.05:", (e[:, 0] <= 1.05).mean(), "<=1.10:", (e[:, 0] <= 1.10).mean())
print("slab count at 1.015 x sample cap: median", np.median(e[:, 1]), "max", e[:, 1].max(), "<=896:", (e[:, 1] <= 896).mean())
print("slab count at 1.03 x sample cap: median", np.median(e[:, 2]), "max", e[:, 2].max(), "<=896:", (e[:, 2] <= 896).mean())
