"""GPU LBVH vs the integer specification (oracle/lbvh_ref.py): Morton codes, sort order and
hierarchy ids BIT-EXACT; boxes contain the exact ellipsoid extents and match within ulps."""
import numpy as np
import pytest

from oracle import lbvh_ref as L
from oracle import ref_numpy as O

from gpu_util import make_scene, random_set

pytestmark = pytest.mark.gpu


def _check(gs):
    scene = make_scene(gs)
    lb = scene.read_lbvh()
    ref = L.build(gs.pos)
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
