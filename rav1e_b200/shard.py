"""Tile / frame sharding across ranks and the one real exchange step of the path:
an all-gather of fixed-size per-block winner records.

rav1e treats tiles as the parallel unit (encoder.rs:3249-3257, api/lookahead.rs:275-283):
candidate evaluations inside a tile only READ the shared reference planes, so units shard with no
data-path collective.  The rank that owns the entropy coder needs every tile's winners, hence one
`all_gather_into_tensor` (NCCL over NVLink on GPUs, gloo in the CPU tests) per stage.
"""
import numpy as np


def owned_units(n_units, rank, world):
    """Round-robin ownership: unit t -> rank t % world (8 tiles -> 8 GPUs one each)."""
    return list(range(rank, n_units, world))


def units_per_rank(n_units, world):
    return (n_units + world - 1) // world


def gather_records(local, n_units, rank, world, group=None):
    """local: uint8 tensor [units_per_rank, record_bytes] holding this rank's units in ownership
    order (padded with zeros when n_units % world != 0).  Returns a uint8 tensor
    [n_units, record_bytes] in GLOBAL unit order on every rank."""
    import torch
    import torch.distributed as dist
    upr = units_per_rank(n_units, world)
    assert local.shape[0] == upr and local.dtype == torch.uint8
    if world == 1:
        return local[:n_units]
    out = torch.empty((world * upr,) + tuple(local.shape[1:]), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    # rank-major [r][k] -> unit t = k * world + r
    out = out.view(world, upr, *local.shape[1:]).transpose(0, 1).reshape(world * upr, *local.shape[1:])
    return out[:n_units]


def tile_grid(frame_w, frame_h, tile_cols_log2, tile_rows_log2, sb_size=64):
    """Uniform tile grid as TilingInfo::from_target_tiles lays it out (tiling/tiler.rs:97-132):
    tile width/height in superblocks = ceil(frame_sb / tiles), last tile takes the remainder.
    Returns a list of (x, y, w, h) luma rectangles in raster order."""
    sb_cols = (frame_w + sb_size - 1) // sb_size
    sb_rows = (frame_h + sb_size - 1) // sb_size
    tw = (sb_cols + (1 << tile_cols_log2) - 1) >> tile_cols_log2
    th = (sb_rows + (1 << tile_rows_log2) - 1) >> tile_rows_log2
    cols = (sb_cols + tw - 1) // tw
    rows = (sb_rows + th - 1) // th
    tiles = []
    for r in range(rows):
        for c in range(cols):
            x, y = c * tw * sb_size, r * th * sb_size
            tiles.append((x, y, min(tw * sb_size, frame_w - x), min(th * sb_size, frame_h - y)))
    return tiles


def blocks_in_rect(rect, bw, bh, block_dtype):
    x0, y0, w, h = rect
    xs = np.arange(x0, x0 + w - bw + 1, bw)
    ys = np.arange(y0, y0 + h - bh + 1, bh)
    b = np.zeros(len(xs) * len(ys), block_dtype)
    b["x"] = np.tile(xs, len(ys))
    b["y"] = np.repeat(ys, len(xs))
    return b
