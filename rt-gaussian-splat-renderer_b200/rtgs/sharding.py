"""Work partitioning for one-process-per-GPU rendering.  The scene is replicated on
every GPU; frames are split by camera view or by image tile; there is no collective on the data path
(only the finished framebuffer pieces are gathered)."""
from __future__ import annotations

import numpy as np

TILE_COLUMNS = 32   # tile = 32 pixel columns x full height: contiguous in the (W,H,3) i-major image


def views_for_rank(n_views: int, rank: int, world: int) -> list[int]:
    """View k -> rank k mod world."""
    return list(range(rank, n_views, world))


def tiles_for_rank(W: int, H: int, rank: int, world: int) -> list[tuple[int, int, int, int]]:
    """32-column strips dealt round-robin to the ranks (interleaving balances dense and empty image
    regions).  Returns (x0, y0, w, h) regions."""
    strips = [(x0, 0, min(TILE_COLUMNS, W - x0), H) for x0 in range(0, W, TILE_COLUMNS)]
    return strips[rank::world]


def assemble_tiles(W: int, H: int, world: int, payloads) -> np.ndarray:
    """Rebuild the (W,H,3) frame from the per-rank concatenated strip buffers (rank order)."""
    frame = np.empty((W, H, 3), dtype=np.float32)
    for rank, buf in enumerate(payloads):
        off = 0
        for (x0, y0, w, h) in tiles_for_rank(W, H, rank, world):
            n = w * h * 3
            frame[x0:x0 + w, y0:y0 + h] = np.asarray(buf[off:off + n]).reshape(w, h, 3)
            off += n
    return frame
