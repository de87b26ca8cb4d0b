"""GPU LBVH vs the integer specification (oracle/lbvh_ref.py): Morton codes, sort order and
hierarchy ids BIT-EXACT; boxes contain the exact ellipsoid extents and match within ulps."""
import numpy as np
import pytest

from oracle import lbvh_ref as L
from oracle import ref_numpy as O

from gpu_util import make_scene, random_set

pytestmark = pytest.mark.gpu


def _check(gs, bits=30):
    scene = make_scene(gs, morton_bits=bits)
    lb = scene.read_lbvh()
    ref = L.build(gs.pos, bits=bits)
    n = gs.n
    assert np.array_equal(lb["morton"], ref["codes"]), "Morton codes differ"
    assert np.array_equal(lb["sorted_idx"], ref["sorted_idx"]), "sort order differs"
    assert np.array_equal(lb["child"], ref["child"]), "hierarchy differs"
    assert np.array_equal(lb["parent"], ref["parent"])
    # leaf boxes: tight sqrt(3)-sigma extents (float64) must be inside, and within 1e-5 relative
    sidx = ref["sorted_idx"]
    lo, hi = O.bounding_box_tight(gs.pos[sidx], gs.rot[sidx], gs.scale[sidx])
    bb = lb["aabb"]
    leaf = bb[n - 1:]
    assert (leaf[:, :3] <= lo).all() and (leaf[:, 3:] >= hi).all(), "leaf box does not contain the ellipsoid"
    ext = (hi - lo).max(axis=1, keepdims=True)
    assert (np.abs(leaf[:, :3] - lo) <= 1e-5 * ext + 1e-6 * np.abs(lo)).all()
    assert (np.abs(leaf[:, 3:] - hi) <= 1e-5 * ext + 1e-6 * np.abs(hi)).all()
    # internal boxes = exact float32 unions of the GPU's own leaf boxes
    bmin, bmax = L.refit(ref["child"], leaf[:, :3], leaf[:, 3:])
    assert np.array_equal(bb[:, :3], bmin) and np.array_equal(bb[:, 3:], bmax)
    return scene


@pytest.mark.parametrize("n", [1, 2, 3, 16, 255, 2049, 100_000])
def test_lbvh_bit_exact(n):
    _check(random_set(n, seed=100 + n, sh=False))


def test_lbvh_duplicates_and_flat_axis():
    gs = random_set(5000, seed=9, sh=False)
    gs.pos[100:400] = gs.pos[100]       # identical centres -> identical codes, order by index
    gs.pos[:, 1] = 0.25                 # zero extent on y
    _check(gs)


def test_lbvh_one_million():
    gs = random_set(1_000_000, seed=1002, mean_scale=0.0026, sh=False)
    scene = _check(gs)
    bf = scene.bvh_field
    assert bf.shape == (2 * gs.n - 1,)
    root = bf[0]
    assert root.prim_left == 0 and root.prim_right == gs.n and root.depth == 0


@pytest.mark.parametrize("n", [1, 2, 3, 16, 255, 2049, 100_000])
def test_lbvh63_bit_exact(n):
    """The optional 63-bit codes (Scene(morton_bits=63), RTGS_OPT_MORTON_BITS): codes, (code, index) order and the
    duplicate-aware Karras hierarchy bit-exact against oracle/lbvh_ref.py build(bits=63)."""
    _check(random_set(n, seed=300 + n, sh=False), bits=63)


def test_lbvh63_duplicates_flat_axis_and_outliers():
    gs = random_set(20000, seed=19, sh=False)
    gs.pos[100:400] = gs.pos[100]       # identical centres -> identical codes, told apart by sorted position
    gs.pos[:, 1] = 0.25                 # zero extent on y
    gs.pos[7] = (4000.0, 0.25, -3000.0)  # far outliers: 30-bit codes would put the cloud into a few cells
    gs.pos[8] = (-5000.0, 0.25, 2500.0)
    _check(gs, bits=63)
    assert len(np.unique(L.morton30(gs.pos))) < 50 < 15000 < len(np.unique(L.morton63(gs.pos)))


def test_image_does_not_depend_on_the_code_width():
    """Same frame from the 30-bit and the 63-bit tree, bit for bit; with far outliers the wide tree visits far
    fewer boxes (the reason it exists)."""
    from gpu_util import make_camera
    from rtgs.ray_tracer import RayTracer
    gs = random_set(30000, seed=23, mean_scale=0.02)
    gs.pos[0] = (3000.0, 2000.0, -4000.0)
    gs.pos[1] = (-2500.0, -3500.0, 3000.0)
    cam, _ = make_camera(0.4, 1.2, 2.5, 320, 200)
    imgs, boxes = [], []
    for bits in (30, 63):
        scene = make_scene(gs, morton_bits=bits)
        rt = RayTracer(cam.buf_size, scene, cam, t_cut=0.0)
        imgs.append(rt.render(16).copy())
        rt.render_device(16, collect_stats=True)
        boxes.append(rt.last_stats["nodes_tested"])
    assert np.array_equal(imgs[0], imgs[1])
    assert boxes[1] < boxes[0]


def test_code_width_is_chosen_automatically():
    """Default Scene(): the 30-bit tree of the specification, unless more than an eighth of its codes repeat."""
    gs = random_set(20000, seed=29, sh=False)
    assert make_scene(gs).morton_bits == 30
    gs.pos[3] = (5000.0, -4000.0, 3000.0)
    scene = make_scene(gs)
    assert scene.morton_bits == 63
    ref = L.build(gs.pos, bits=63)
    lb = scene.read_lbvh()
    assert np.array_equal(lb["morton"], ref["codes"]) and np.array_equal(lb["child"], ref["child"])
    assert make_scene(gs, morton_bits=30).morton_bits == 30
