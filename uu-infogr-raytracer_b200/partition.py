"""Row-tile partition of a frame across cooperating GPUs — the host-side mirror of the arithmetic in
csrc/rtb200.cu (make_params / render_loop): tile t of `tile_rows` rows belongs to rank t % world.
Used by bench.py (NCCL-gather comparison path) and by the world_size-2 gloo tests; no CUDA here."""
from __future__ import annotations

import numpy as np


def n_tiles(height: int, tile_rows: int) -> int:
    return (height + tile_rows - 1) // tile_rows


def tiles_of_rank(height: int, tile_rows: int, rank: int, world: int) -> np.ndarray:
    return np.arange(rank, n_tiles(height, tile_rows), world, dtype=np.int64)


def rows_of_rank(height: int, tile_rows: int, rank: int, world: int) -> np.ndarray:
    """Sorted row indices owned by `rank`."""
    t = tiles_of_rank(height, tile_rows, rank, world)
    rows = (t[:, None] * tile_rows + np.arange(tile_rows)[None, :]).reshape(-1)
    return rows[rows < height]


def max_rows_per_rank(height: int, tile_rows: int, world: int) -> int:
    return max(len(rows_of_rank(height, tile_rows, r, world)) for r in range(world))


def pack_rows(frame, rows, pad_to: int):
    """frame: [h, w] tensor/array; returns [pad_to, w] with this rank's rows first (zero padded) — the gather payload."""
    import torch
    out = torch.zeros((pad_to, frame.shape[1]), dtype=frame.dtype, device=frame.device)
    idx = torch.as_tensor(rows, device=frame.device)
    out[: len(rows)] = frame.index_select(0, idx)
    return out


def assemble(gathered, height: int, width: int, tile_rows: int):
    """gathered: list (one per rank) of [pad_to, w] tensors -> the full [h, w] frame on gathered[0]'s device."""
    import torch
    world = len(gathered)
    frame = torch.empty((height, width), dtype=gathered[0].dtype, device=gathered[0].device)
    for r in range(world):
        rows = rows_of_rank(height, tile_rows, r, world)
        frame[torch.as_tensor(rows, device=frame.device)] = gathered[r][: len(rows)]
    return frame
