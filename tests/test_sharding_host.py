"""Host-side helpers of rtgs.sharding that need no GPU."""
import numpy as np

from rtgs import sharding


def test_numa_binding_never_raises_and_reports():
    r = sharding.bind_to_gpu_numa_node(0)
    assert isinstance(r, dict) and "bound" in r
    assert r["bound"] or "why" in r


def test_views_and_tiles_partition_the_work():
    for world in (1, 2, 3, 8):
        views = sorted(v for r in range(world) for v in sharding.views_for_rank(64, r, world))
        assert views == list(range(64))
        cols = np.concatenate([sharding.stripe_columns(1920, r, world) for r in range(world)])
        assert np.array_equal(np.sort(cols), np.arange(1920))
