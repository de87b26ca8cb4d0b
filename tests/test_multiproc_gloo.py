"""Host-side logic of the multi-GPU path with world_size 2 on CPU (gloo): the view / tile partition
functions used by bench.py and the headless driver cover the work exactly once, and a gathered
framebuffer assembled from per-rank tile renders equals the full frame (render stubbed by the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rtgs.sharding import assemble_tiles, tiles_for_rank, views_for_rank


def test_partitions_cover_exactly_once():
    for world in (1, 2, 3, 4, 8):
        views = sorted(v for r in range(world) for v in views_for_rank(64, r, world))
        assert views == list(range(64))
        for (W, H) in ((1920, 1080), (3840, 2160), (100, 37)):
            cover = np.zeros((W, H), np.int32)
            for r in range(world):
                for (x0, y0, w, h) in tiles_for_rank(W, H, r, world):
                    assert w > 0 and h > 0 and x0 % 32 == 0
                    cover[x0:x0 + w, y0:y0 + h] += 1
            assert (cover == 1).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, W, H, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # stand-in renderer: a deterministic function of the pixel, evaluated only on this rank's tiles
        def render_tile(x0, y0, w, h):
            ii, jj = np.meshgrid(np.arange(x0, x0 + w), np.arange(y0, y0 + h), indexing="ij")
            return np.stack([np.sin(ii * 0.1) * jj, ii + 0.5 * jj, ii * jj % 7], axis=-1).astype(np.float32)

        # the gather bench.py --sharding tiles uses: every rank holds a frame with only its stripes filled in
        # (what Scene.set_stripe(world, rank) renders); rank 0 ends up with the whole frame
        from rtgs.sharding import StripeGather
        part = torch.zeros((W, H, 3), dtype=torch.float32)
        for (x0, y0, w, h) in tiles_for_rank(W, H, rank, world):
            part[x0:x0 + w, y0:y0 + h] = torch.from_numpy(render_tile(x0, y0, w, h))
        got = StripeGather(W, H, rank, world, torch.device("cpu"))(part, dist)
        # the host-side assembly of concatenated strip payloads (headless driver)
        payloads = [np.concatenate([render_tile(*t).reshape(-1) for t in tiles_for_rank(W, H, r, world)])
                    for r in range(world)]
        if rank == 0:
            want = render_tile(0, 0, W, H)
            q.put(bool(np.array_equal(got.numpy(), want)) and
                  bool(np.array_equal(assemble_tiles(W, H, world, payloads), want)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_tile_gather_equals_full_frame():
    world, W, H = 2, 144, 48      # 4.5 stripes: ragged last stripe, unequal stripe counts
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
