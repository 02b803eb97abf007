"""BASELINE config 0 as a parity case: the luma of tests/small_input.y4m (5 frames, 64x64, 8 bit;
fixture tests/golden/small_input_luma.npy extracted by tests/golden/make_golden.py) run through
the speed-10 intra shape of the path — 32x32 partitions only (speedsettings.rs:184-188), the 13
RAV1E_INTRA_MODES (predict.rs:42-56) predicted per block, SATD against the source picks the mode
(rdo.rs:1477-1500's pruning metric), the winner's residual goes through the 32x32 DCT_DCT.

Plumbing, not the encoder: neighbours come from the SOURCE frame (edge-replicated) instead of the
reconstruction, and every edge counts as available, so that the inputs are real picture content
without rebuilding rav1e's reconstruction loop.  Both sides (oracle / CUDA) consume the same edges.
"""
import os

import numpy as np

BS, PAD = 32, 64
# (PredictionMode discriminant, variant BOTH, angle): DC, H, V, PAETH, SMOOTH_H, SMOOTH, SMOOTH_V, D45..D67
MODES13 = [(0, 3, 0), (2, 3, 180), (1, 3, 90), (12, 3, 0), (11, 3, 0), (9, 3, 0), (10, 3, 0), (3, 3, 45),
           (4, 3, 135), (5, 3, 113), (6, 3, 157), (7, 3, 203), (8, 3, 67)]


def load_luma():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_input_luma.npy"))


def blocks_of(frame):
    h, w = frame.shape
    return [(x, y) for y in range(0, h, BS) for x in range(0, w, BS)]


def intra_edge(padded, x, y, n=2 * BS):
    """IntraEdge layout (partition.rs:600-637): top-left at [128], left bottom->top ending at
    [127], above from [129]; n pixels each way."""
    e = np.zeros(257, np.uint8)
    px, py = x + PAD, y + PAD
    e[128] = padded[py - 1, px - 1]
    e[128 - n:128] = padded[py:py + n, px - 1][::-1]
    e[129:129 + n] = padded[py - 1, px:px + n]
    return e


def padded(frame):
    return np.pad(frame, PAD, mode="edge")
