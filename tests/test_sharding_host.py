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


def test_peer_frame_arrival_targets():
    """PeerFrame hands frames over with one arrive counter per buffer: frame f uses buffer f % buffers, and rank 0
    waits for world x (number of frames that have used that buffer so far).  Simulate producers running ahead by
    up to `buffers` frames: the target of a frame is reached exactly when all ranks have delivered that frame."""
    for buffers in (1, 2, 3, 4):
        for f in range(1, 40):
            brute = sum(1 for g in range(1, f + 1) if g % buffers == f % buffers)
            assert sharding.PeerFrame.uses_of_buffer(f, buffers) == brute
    world, buffers = 3, 2
    arrive = [0] * buffers
    delivered = [0] * world
    rng = np.random.default_rng(0)
    consumed = 0
    for _ in range(400):
        r = int(rng.integers(world))
        nxt = delivered[r] + 1
        if nxt - buffers <= consumed and nxt <= 30:          # the grant: buffer reuse needs frame nxt - buffers consumed
            delivered[r] = nxt
            arrive[nxt % buffers] += 1
        f = consumed + 1
        target = world * sharding.PeerFrame.uses_of_buffer(f, buffers)
        if arrive[f % buffers] >= target:
            assert min(delivered) >= f                       # rank 0 never sees a frame before every rank delivered it
            consumed = f
    assert consumed == 30
