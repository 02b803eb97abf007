"""World-size-2 gloo test of the multi-GPU host logic: tiles are owned round-robin,
each rank produces winner records for its own tiles only, and one all-gather delivers all of
them to every rank in global tile order.  The per-tile "winners" here come from the oracle's
full search on the rank's tiles, so the assembled result must equal a single-process pass."""
import os
import socket

import numpy as np
import pytest

from rav1e_b200 import backend as B
from rav1e_b200 import shard
from tests import oracle_lib as O


def test_tile_grid_matches_reference_layout():
    # config 5: 4K, tile_cols_log2 = 2, tile_rows_log2 = 1 -> 15x17 SB tiles = 960x1088 px
    tiles = shard.tile_grid(3840, 2160, 2, 1)
    assert len(tiles) == 8
    assert tiles[0] == (0, 0, 960, 1088) and tiles[7] == (2880, 1088, 960, 1072)
    assert sum(w * h for _, _, w, h in tiles) == 3840 * 2160
    assert shard.owned_units(8, 3, 8) == [3] and shard.owned_units(8, 1, 2) == [1, 3, 5, 7]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _tile_winners(tile_index, tiles, cur, ref):
    blocks = shard.blocks_in_rect(tiles[tile_index], 16, 16, B.BLOCK_DTYPE)
    return O.full_search_blocks(cur, ref, blocks, 16, 16, 8, 8, 4, 320, threads=1)


def _make_frame():
    W, H, PAD = 256, 128, 64
    rng = np.random.default_rng(0)
    cur, ref = O.Plane(W, H, PAD), O.Plane(W, H, PAD)
    img = rng.integers(0, 256, (H, W), dtype=np.uint8)
    ref.fill_from(img)
    cur.fill_from(np.roll(img, (2, -4), axis=(0, 1)))
    return cur, ref, shard.tile_grid(W, H, 1, 1)           # 2x2 tiles of 128x64


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cur, ref, tiles = _make_frame()
    nb = 8 * 4                                             # 16x16 blocks per 128x64 tile
    upr = shard.units_per_rank(len(tiles), world)
    local = torch.zeros((upr, nb * 16), dtype=torch.uint8)
    for k, t in enumerate(shard.owned_units(len(tiles), rank, world)):
        local[k] = torch.from_numpy(_tile_winners(t, tiles, cur, ref).view(np.uint8).copy())
    allw = shard.gather_records(local, len(tiles), rank, world)
    np.save(out_path % rank, allw.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    out = str(tmp_path / "winners_%d.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    cur, ref, tiles = _make_frame()
    want = np.stack([_tile_winners(t, tiles, cur, ref).view(np.uint8) for t in range(len(tiles))])
    for r in range(2):
        np.testing.assert_array_equal(np.load(out % r), want)
