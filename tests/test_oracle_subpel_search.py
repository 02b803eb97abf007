"""subpel_diamond_search (me.rs:1311-1383) of the oracle: invariants that hold for any correct
restatement - the result is never worse than the start, it is a strict local minimum of the last
radius' diamond, and on a pure half-pel-interpolated shift it walks to the true vector."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib as O

BLOCK = np.dtype([("x", "<i2"), ("y", "<i2")])
CAND = np.dtype([("block", "<u4"), ("mv_row", "<i2"), ("mv_col", "<i2")])


def search(ocur, oref, blocks, start, n, use_satd, lam, allow_hp, bd=8):
    L = O.lib()
    L.orc_subpel_diamond_search_blocks.restype = None
    L.orc_subpel_diamond_search_blocks.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t] + [C.c_int] * 3 + \
        [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p] + [C.c_int] * 3 + [C.c_void_p, C.c_int]
    res = start.copy()
    L.orc_subpel_diamond_search_blocks(ocur.origin_ptr(), ocur.stride, oref.origin_ptr(), oref.stride, ocur.bpp,
                                       2 * ((ocur.width + 7) >> 3), 2 * ((ocur.height + 7) >> 3), O.ptr(blocks),
                                       len(blocks), n, n, int(use_satd), lam, None, int(allow_hp), 0, bd, O.ptr(res), 0)
    return res


@pytest.mark.parametrize("allow_hp", [False, True])
def test_result_is_a_local_minimum_never_worse_than_the_start(allow_hp):
    rng = np.random.default_rng(4)
    W, H, PAD, n = 128, 96, 64, 16
    base = rng.integers(0, 256, (H + 8, W + 8)).astype(np.float64)
    k = np.ones(5) / 5
    base = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 1, base)
    base = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 0, base)
    ref = np.ascontiguousarray(np.rint(base[4:4 + H, 4:4 + W]).astype(np.uint8))
    cur = np.ascontiguousarray(np.rint((base[4:4 + H, 4:4 + W] + base[4:4 + H, 5:5 + W]) / 2).astype(np.uint8))  # +1/2 px
    ocur, oref = O.Plane(W, H, PAD), O.Plane(W, H, PAD)
    ocur.fill_from(cur)
    oref.fill_from(ref)
    xs, ys = np.arange(16, W - 32, 16), np.arange(16, H - 32, 16)
    blocks = np.zeros(len(xs) * len(ys), BLOCK)
    blocks["x"], blocks["y"] = np.tile(xs, len(ys)), np.repeat(ys, len(xs))
    nb = len(blocks)
    lam = 0
    cands = np.zeros(nb, CAND)
    cands["block"] = np.arange(nb)
    sad, cost = O.fullpel_candidates(ocur, oref, blocks, cands, n, n, False, lam)
    ME = O.ME_RESULT_DTYPE
    start = np.zeros(nb, ME)
    start["cost"], start["sad"] = cost, sad
    res = search(ocur, oref, blocks, start, n, False, lam, allow_hp)
    assert (res["cost"] <= start["cost"]).all()
    # half-pel shift to the right: most blocks end on (row 0, col +4)
    assert (np.abs(res["mv_col"] - 4) <= 1).mean() > 0.8 and (np.abs(res["mv_row"]) <= 1).mean() > 0.8
    r = 1 if allow_hp else 2
    for dr, dc in ((r, 0), (0, r), (-r, 0), (0, -r)):
        nbr = np.zeros(nb, CAND)
        nbr["block"] = np.arange(nb)
        nbr["mv_row"], nbr["mv_col"] = res["mv_row"] + dr, res["mv_col"] + dc
        _, c2 = O.subpel_candidates(ocur, oref, blocks, nbr, n, n, False, lam, None, allow_hp, 0, 8)
        assert (c2 >= res["cost"]).all()
