"""Checks of the integer LBVH specification (oracle/lbvh_ref.py): known-answer Morton codes,
key uniqueness / order, and structural validity of the Karras hierarchy."""
import numpy as np
import pytest

from oracle import lbvh_ref as L


def test_expand_bits_kat():
    assert L.expand_bits(np.array([0b1], np.uint32))[0] == 0b1
    assert L.expand_bits(np.array([0b11], np.uint32))[0] == 0b1001
    assert L.expand_bits(np.array([1023], np.uint32))[0] == 0x09249249


def test_morton_kat():
    pos = np.array([[0, 0, 0], [1, 1, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.5, 0.5, 0.5]], np.float32)
    c = L.morton30(pos)
    assert c[0] == 0 and c[1] == 0x3FFFFFFF
    assert c[2] == 0x09249249 << 2 and c[3] == 0x09249249 << 1 and c[4] == 0x09249249
    assert c[5] == (L.expand_bits(np.array([512], np.uint32))[0] * 7)
    # zero extent on an axis -> that axis contributes 0
    flat = np.array([[0, 5, 0], [1, 5, 1]], np.float32)
    assert (L.morton30(flat) & (0x09249249 << 1)).max() == 0


@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 1000])
def test_hierarchy_is_a_valid_tree(n):
    rng = np.random.default_rng(n)
    pos = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    if n >= 64:
        pos[10:20] = pos[10]          # duplicates: identical codes, ordered by index
    b = L.build(pos)
    keys = b["keys"]
    assert (np.diff(keys.astype(np.uint64)) > 0).all() if n > 1 else True
    assert np.array_equal(np.sort(b["sorted_idx"]), np.arange(n))
    child, parent, rng_ = b["child"], b["parent"], b["rng"]
    if n == 1:
        assert child.shape == (0, 2) and parent.tolist() == [-1]
        return
    assert parent[0] == -1 and (parent[1:] >= 0).all()
    # every node except the root is the child of exactly one internal node
    assert np.array_equal(np.sort(child.reshape(-1)), np.arange(1, 2 * n - 1))
    # ranges: root covers everything, children split the parent's range at gamma
    assert rng_[0].tolist() == [0, n - 1]
    for i in range(n - 1):
        l, r = child[i]
        lf, ll = (l - (n - 1), l - (n - 1)) if l >= n - 1 else rng_[l]
        rf, rl = (r - (n - 1), r - (n - 1)) if r >= n - 1 else rng_[r]
        assert lf == rng_[i][0] and rl == rng_[i][1] and ll + 1 == rf


def test_morton63_kat_and_tree():
    assert L.expand_bits21(np.array([1], np.uint64))[0] == 1
    assert L.expand_bits21(np.array([0b11], np.uint64))[0] == 0b1001
    assert L.expand_bits21(np.array([0x1FFFFF], np.uint64))[0] == 0x1249249249249249
    pos = np.array([[0, 0, 0], [1, 1, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    c = L.morton63(pos)
    assert c[0] == 0 and c[1] == 0x7FFFFFFFFFFFFFFF
    assert c[2] == 0x1249249249249249 << 2 and c[3] == 0x1249249249249249 << 1 and c[4] == 0x1249249249249249
    # the top 30 bits of the wide code are the 30-bit code
    rng = np.random.default_rng(3)
    p = rng.uniform(-1, 1, (4000, 3)).astype(np.float32)
    assert np.array_equal((L.morton63(p) >> np.uint64(33)).astype(np.uint32), L.morton30(p))
    # duplicates: order by index, hierarchy still a valid tree with the positional tie-break
    for n in (2, 3, 7, 64, 1000):
        q = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        if n >= 7:
            q[2:6] = q[2]
        b = L.build(q, bits=63)
        keys, sidx, child, rng_ = b["keys"], b["sorted_idx"], b["child"], b["rng"]
        assert (keys[1:] >= keys[:-1]).all() and np.array_equal(np.sort(sidx), np.arange(n))
        same = keys[1:] == keys[:-1]
        assert (np.diff(sidx.astype(np.int64))[same] > 0).all()
        assert np.array_equal(np.sort(child.reshape(-1)), np.arange(1, 2 * n - 1))
        assert rng_[0].tolist() == [0, n - 1]
        for i in range(n - 1):
            l, r = child[i]
            lf, ll = (l - (n - 1), l - (n - 1)) if l >= n - 1 else rng_[l]
            rf, rl = (r - (n - 1), r - (n - 1)) if r >= n - 1 else rng_[r]
            assert lf == rng_[i][0] and rl == rng_[i][1] and ll + 1 == rf


def test_refit_contains_leaves():
    rng = np.random.default_rng(1)
    n = 500
    pos = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    b = L.build(pos)
    p = pos[b["sorted_idx"]]
    bmin, bmax = L.refit(b["child"], p - 0.01, p + 0.01)
    assert np.allclose(bmin[0], p.min(0) - 0.01) and np.allclose(bmax[0], p.max(0) + 0.01)
    for i in range(n - 1):
        f, l = b["rng"][i]
        assert np.array_equal(bmin[i], (p[f:l + 1] - np.float32(0.01)).min(0))
