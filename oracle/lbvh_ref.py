"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — integer LBVH specification in NumPy.

The reference has no Morton codes, no sort and no LBVH (its builder, scene.py:162-404, is a
host-driven binned-SAH build that the north star replaces), so THIS file is the definition that
"bit-exact Morton codes and sort keys" is measured against (SURVEY.md §7 "Specs to freeze").
The split key is the Gaussian centre, as in the reference (scene.py:263-266).

Spec
----
lo/hi  = per-axis min/max of the float32 centres;  inv = 1/(hi-lo) in float32 (0 if hi == lo)
x      = (p - lo) * inv            float32 subtract, then float32 multiply
u      = uint32(min(max(x * 1024, 0), 1023))         truncation
code   = expand(ux) << 2 | expand(uy) << 1 | expand(uz)        30 bits (Karras bit spread)
key    = uint64(code) << 32 | uint32(original index)           unique
order  = ascending key
tree   = Karras 2012 over the sorted keys with delta(i,j) = clz64(key_i ^ key_j)
ids    = internal [0, n-1), leaves [n-1, 2n-1) in sorted order; root = internal 0

Optional wide variant (morton63 / build(pos, bits=63)): 21 bits per axis, order = ascending (code, index),
delta(i,j) = clz64(code_i ^ code_j), or 64 + clz64(i ^ j) for equal codes.
"""
from __future__ import annotations

import numpy as np


def scene_bounds(pos):
    pos = np.asarray(pos, dtype=np.float32)
    return pos.min(axis=0), pos.max(axis=0)


def expand_bits(v):
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def morton30(pos, lo=None, hi=None):
    """30-bit Morton codes of float32 centres (uint32)."""
    pos = np.asarray(pos, dtype=np.float32)
    if lo is None or hi is None:
        lo, hi = scene_bounds(pos)
    lo = np.asarray(lo, dtype=np.float32)
    hi = np.asarray(hi, dtype=np.float32)
    ext = (hi - lo).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(ext > 0, np.float32(1.0) / ext, np.float32(0.0)).astype(np.float32)
    x = ((pos - lo).astype(np.float32) * inv).astype(np.float32)
    q = np.minimum(np.maximum((x * np.float32(1024.0)).astype(np.float32), np.float32(0.0)),
                   np.float32(1023.0))
    u = q.astype(np.uint32)
    with np.errstate(over="ignore"):
        return (expand_bits(u[:, 0]) << np.uint32(2)) | (expand_bits(u[:, 1]) << np.uint32(1)) | expand_bits(u[:, 2])


def expand_bits21(v):
    """Spread the low 21 bits of v two apart (uint64)."""
    v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    v = (v | (v << np.uint64(32))) & np.uint64(0x001F00000000FFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x001F0000FF0000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


def morton63(pos, lo=None, hi=None):
    """63-bit Morton codes (uint64), the optional wide variant (RTGS_OPT_MORTON_BITS = 63):
    x as for morton30, u = uint32(min(max(x * 2^21, 0), 2^21 - 1)), code = spread(ux) << 2 | spread(uy) << 1 |
    spread(uz).  Codes may repeat: the order is ascending (code, original index) and the hierarchy tells equal
    codes apart by sorted position (karras(..., duplicates=True))."""
    pos = np.asarray(pos, dtype=np.float32)
    if lo is None or hi is None:
        lo, hi = scene_bounds(pos)
    lo = np.asarray(lo, dtype=np.float32)
    hi = np.asarray(hi, dtype=np.float32)
    ext = (hi - lo).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(ext > 0, np.float32(1.0) / ext, np.float32(0.0)).astype(np.float32)
    x = ((pos - lo).astype(np.float32) * inv).astype(np.float32)
    q = np.minimum(np.maximum((x * np.float32(2097152.0)).astype(np.float32), np.float32(0.0)),
                   np.float32(2097151.0))
    u = q.astype(np.uint32)
    return (expand_bits21(u[:, 0]) << np.uint64(2)) | (expand_bits21(u[:, 1]) << np.uint64(1)) | expand_bits21(u[:, 2])


def sort_codes63(codes):
    """(sorted 63-bit codes, sorted original indices uint32): ascending (code, index)."""
    order = np.argsort(codes, kind="stable")
    return codes[order], order.astype(np.uint32)


def sort_keys(codes):
    """(sorted 64-bit keys, sorted original indices uint32)."""
    n = codes.shape[0]
    keys = (codes.astype(np.uint64) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    keys = np.sort(keys)
    return keys, (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def _clz64(x):
    """Count leading zeros of non-zero uint64 values."""
    x = x.astype(np.uint64)
    n = np.zeros(x.shape, dtype=np.int64)
    for shift in (32, 16, 8, 4, 2, 1):
        hi = x >> np.uint64(shift)
        nz = hi != 0
        x = np.where(nz, hi, x)
        n += np.where(nz, 0, shift)
    return n


def karras(keys, duplicates=False):
    """Karras 2012 hierarchy.  Returns child (n-1,2) int32 with node ids in the unified id
    space (leaf k -> n-1+k), parent (2n-1,) int32 (root: -1), rng (n-1,2) int32 = the sorted
    leaf range [first,last] covered by each internal node.  duplicates=True: `keys` are sorted codes that
    may repeat; equal ones are told apart by position, delta = 64 + clz64(i ^ j) (Karras 2012, section 4)."""
    keys = np.asarray(keys, dtype=np.uint64)
    n = keys.shape[0]
    if n == 1:
        return np.zeros((0, 2), np.int32), np.array([-1], np.int32), np.zeros((0, 2), np.int32)
    i = np.arange(n - 1, dtype=np.int64)

    def delta(a, b):
        ok = (b >= 0) & (b < n)
        bb = np.clip(b, 0, n - 1)
        x = keys[a] ^ keys[bb]
        if duplicates:
            same = ok & (x == 0)
            pos = (a.astype(np.uint64) ^ bb.astype(np.uint64)) | (~same).astype(np.uint64)
            return np.where(same, 64 + _clz64(pos), np.where(ok, _clz64(x | (~ok | same).astype(np.uint64)), -1))
        return np.where(ok, _clz64(x | (~ok).astype(np.uint64)), -1)

    d = np.where(delta(i, i + 1) - delta(i, i - 1) >= 0, 1, -1).astype(np.int64)
    # sign(): delta values are never equal for unique keys except both -1 (n == 1, excluded).
    dmin = delta(i, i - d)
    lmax = np.full(n - 1, 2, dtype=np.int64)
    while True:
        grow = delta(i, i + lmax * d) > dmin
        if not grow.any():
            break
        lmax = np.where(grow, lmax * 2, lmax)
    l = np.zeros(n - 1, dtype=np.int64)
    t = lmax // 2
    while (t >= 1).any():
        act = t >= 1
        cond = act & (delta(i, i + (l + t) * d) > dmin)
        l = np.where(cond, l + t, l)
        t = np.where(act, t // 2, 0)
    j = i + l * d
    dnode = delta(i, j)
    s = np.zeros(n - 1, dtype=np.int64)
    t = l.copy()
    act = np.ones(n - 1, dtype=bool)
    while act.any():
        t = np.where(act, (t + 1) >> 1, t)
        cond = act & (delta(i, i + (s + t) * d) > dnode)
        s = np.where(cond, s + t, s)
        act = act & (t > 1)
    gamma = i + s * d + np.minimum(d, 0)
    first = np.minimum(i, j)
    last = np.maximum(i, j)
    left = np.where(first == gamma, (n - 1) + gamma, gamma)
    right = np.where(last == gamma + 1, (n - 1) + gamma + 1, gamma + 1)
    child = np.stack([left, right], axis=-1).astype(np.int32)
    parent = np.full(2 * n - 1, -1, dtype=np.int32)
    parent[child[:, 0]] = i.astype(np.int32)
    parent[child[:, 1]] = i.astype(np.int32)
    return child, parent, np.stack([first, last], axis=-1).astype(np.int32)


def refit(child, leaf_min, leaf_max):
    """Bottom-up AABBs for the unified id space: (2n-1,3) min and max (float32 min/max are
    exact, so the result is independent of evaluation order)."""
    n = leaf_min.shape[0]
    bmin = np.empty((2 * n - 1, 3), dtype=np.float32)
    bmax = np.empty((2 * n - 1, 3), dtype=np.float32)
    bmin[n - 1:] = leaf_min
    bmax[n - 1:] = leaf_max
    if n == 1:
        return bmin, bmax
    done = np.zeros(2 * n - 1, dtype=bool)
    done[n - 1:] = True
    pending = np.arange(n - 1)
    while pending.size:
        ready = done[child[pending, 0]] & done[child[pending, 1]]
        r = pending[ready]
        assert r.size, "cycle in hierarchy"
        bmin[r] = np.minimum(bmin[child[r, 0]], bmin[child[r, 1]])
        bmax[r] = np.maximum(bmax[child[r, 0]], bmax[child[r, 1]])
        done[r] = True
        pending = pending[~ready]
    return bmin, bmax


def build(pos, bits=30):
    """Full integer pipeline: dict(codes, keys, sorted_idx, child, parent, rng).  bits=63: the wide variant
    (keys = the sorted codes)."""
    if bits == 63:
        codes = morton63(pos)
        keys, sidx = sort_codes63(codes)
        child, parent, rng = karras(keys, duplicates=True)
        return dict(codes=codes, keys=keys, sorted_idx=sidx, child=child, parent=parent, rng=rng)
    codes = morton30(pos)
    keys, sidx = sort_keys(codes)
    child, parent, rng = karras(keys)
    return dict(codes=codes, keys=keys, sorted_idx=sidx, child=child, parent=parent, rng=rng)
