"""GPU parity for the device-side search stages of full_pixel_me (me.rs:692-856): CUDA (through
the C ABI) == the oracle's restatement, bit exact, block by block.

The reference has no stored vectors for these stages (they need rustc); the oracle restates
get_best_predictor / fullpel_diamond_search / hexagon_search / uneven_multi_hex_search line by
line on top of the pinned SAD, and these tests pin the CUDA path to it.
"""
import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests import oracle_lib as O

pytestmark = pytest.mark.gpu


def predictor_subsets(nblocks, nsubsets, seed, max_px, sizes=(1, 5, 4), drop_median_every=5):
    """Random predictor subsets per block (1/8-pel units, full-pel aligned like real predictors
    plus a few sub-pel ones: get_subset_predictors does not quantise every entry)."""
    rng = np.random.default_rng(seed)
    counts = np.zeros((nblocks, nsubsets), np.int64)
    if nsubsets == 1:
        counts[:, 0] = rng.integers(0, 10, nblocks)          # an empty all_mvs is legal
    else:
        counts[:, 0] = 1
        counts[::drop_median_every, 0] = 0                   # no median (`if let Some`)
        counts[:, 1] = rng.integers(0, sizes[1] + 1, nblocks)
        counts[:, 2] = rng.integers(0, sizes[2] + 1, nblocks)
    offs = np.zeros(nblocks * nsubsets + 1, np.uint32)
    offs[1:] = np.cumsum(counts.reshape(-1))
    n = int(offs[-1])
    p = np.zeros(n, B.CAND_DTYPE)
    mv = rng.integers(-max_px, max_px + 1, (n, 2)) * 8
    sub = rng.random(n) < 0.15
    mv[sub] += rng.integers(-7, 8, (int(sub.sum()), 2))
    p["mv_row"], p["mv_col"] = mv[:, 0], mv[:, 1]
    return p, offs


SEARCH_CASES = [
    # (w, h, dtype, bit_depth, nsubsets, umh_range, lambda, smooth)
    (16, 16, np.uint8, 8, 1, 0, 600, True),
    (16, 16, np.uint8, 8, 3, 24, 600, True),
    (16, 16, np.uint8, 8, 3, 24, 0, False),      # noise planes: long walks, many ties on rate 0
    (8, 8, np.uint8, 8, 3, 24, 300, True),
    (32, 32, np.uint8, 8, 3, 24, 1500, True),
    (64, 64, np.uint8, 8, 1, 0, 1500, True),
    (16, 8, np.uint8, 8, 3, 16, 300, True),
    (4, 4, np.uint8, 8, 3, 24, 100, True),
    (16, 16, np.uint16, 10, 3, 24, 2400, True),
    (32, 16, np.uint16, 12, 1, 0, 9000, True),
    # blocks beyond the lanes' org-word cache with fewer than 8 words per row (the uncached path
    # covers two rows per step) and its 16-bit counterpart
    (16, 64, np.uint8, 8, 3, 24, 900, True),
    (16, 64, np.uint8, 8, 1, 0, 0, False),
    (8, 32, np.uint16, 10, 3, 16, 900, True),
    (64, 16, np.uint16, 10, 1, 0, 2400, True),
]


@pytest.mark.parametrize("w,h,dtype,bd,nsubsets,umh_range,lam,smooth", SEARCH_CASES)
def test_search_matches_oracle(w, h, dtype, bd, nsubsets, umh_range, lam, smooth):
    W, H, PAD = 352, 288, 160
    cur, ref = G.make_planes(W, H, PAD, dtype, seed=w * 7 + h + nsubsets, bit_depth=bd, smooth=smooth,
                             shift=(5, -3))
    ocur, dcur = G.both_planes(cur, PAD)
    oref, dref = G.both_planes(ref, PAD)
    blocks = G.grid_blocks(W, H, w, h)
    n = len(blocks)
    preds, offs = predictor_subsets(n, nsubsets, seed=3 + w, max_px=24)
    rng = np.random.default_rng(5)
    pmv = (rng.integers(-40, 41, (n, 4)) * 4).astype(np.int16)
    # thresholds around typical SADs so that every exit of the ladder is taken by some blocks
    thresh = rng.integers(0, (w * h) << (bd - 8 + 3), n).astype(np.uint32) if nsubsets == 3 else None
    want = O.full_pixel_me_blocks(ocur, oref, blocks, preds, offs, nsubsets, w, h, lam, pmv, thresh, umh_range)

    c = G.ctx()
    p = B.me_params(w, h, W, H, lam, bit_depth=bd)
    d_best = G.dev_empty(16 * n)
    c.me_search_dev(dcur, dref, G.to_dev(blocks), n, G.to_dev(preds) if len(preds) else None, G.to_dev(offs),
                    nsubsets, p, d_best, G.to_dev(pmv), G.to_dev(thresh) if thresh is not None else None,
                    umh_range)
    c.synchronize()
    got = G.from_dev(d_best, B.ME_RESULT_DTYPE)[:n]
    for f in ("cost", "sad", "mv_row", "mv_col"):
        np.testing.assert_array_equal(got[f], want[f], err_msg=f)
    # the searches really moved: some winners are not among the predictors' own positions
    assert (got["cost"] != np.uint64(2**64 - 1)).mean() > 0.9
    for pl in (dcur, dref):
        c.plane_free(pl)


def test_search_frame_edges_and_out_of_range_predictors():
    """Blocks on the frame border with predictors beyond get_mv_range (me.rs:339-362): the range
    check of get_fullpel_mv_rd must reject them in every stage."""
    W, H, PAD, w, h = 128, 96, 192, 16, 16
    cur, ref = G.make_planes(W, H, PAD, np.uint8, seed=9)
    ocur, dcur = G.both_planes(cur, PAD)
    oref, dref = G.both_planes(ref, PAD)
    blocks = G.grid_blocks(W, H, w, h)
    n = len(blocks)
    preds, offs = predictor_subsets(n, 3, seed=77, max_px=60)      # range is only +-(16 + w) px past the edge
    thresh = np.zeros(n, np.uint32)                                # never exit early
    want = O.full_pixel_me_blocks(ocur, oref, blocks, preds, offs, 3, w, h, 800, None, thresh, 24)
    c = G.ctx()
    p = B.me_params(w, h, W, H, 800)
    d_best = G.dev_empty(16 * n)
    c.me_search_dev(dcur, dref, G.to_dev(blocks), n, G.to_dev(preds), G.to_dev(offs), 3, p, d_best, None,
                    G.to_dev(thresh), 24)
    c.synchronize()
    got = G.from_dev(d_best, B.ME_RESULT_DTYPE)[:n]
    for f in ("cost", "sad", "mv_row", "mv_col"):
        np.testing.assert_array_equal(got[f], want[f], err_msg=f)
    for pl in (dcur, dref):
        c.plane_free(pl)


def test_search_multi_pair_equals_per_pair():
    W, H, PAD, w, h = 192, 128, 128, 16, 16
    c = G.ctx()
    imgs = [G.make_planes(W, H, PAD, np.uint8, seed=40 + k, shift=(k, -k)) for k in range(3)]
    dpl = [(c.plane_from_host(a, PAD), c.plane_from_host(b, PAD)) for a, b in imgs]
    NP = 5
    curs = [dpl[k % 3][0] for k in range(NP)]
    refs = [dpl[(k + 1) % 3][1] for k in range(NP)]
    grid = G.grid_blocks(W, H, w, h)
    nb = len(grid)
    blocks = np.tile(grid, NP)
    preds, offs = predictor_subsets(nb * NP, 3, seed=11, max_px=20)
    thresh = np.random.default_rng(2).integers(0, 4000, nb * NP).astype(np.uint32)
    p = B.me_params(w, h, W, H, 500)
    d_blocks, d_preds, d_offs, d_thresh = map(G.to_dev, (blocks, preds, offs, thresh))
    d_best = G.dev_empty(16 * nb * NP)
    ends = [(k + 1) * nb for k in range(NP)]
    c.me_search_multi_dev(B.PlanePairs(curs, refs, ends, ends), d_blocks, nb * NP, d_preds, d_offs, 3, p, d_best,
                          None, d_thresh, 24)
    c.synchronize()
    got = G.from_dev(d_best, B.ME_RESULT_DTYPE)[:nb * NP].copy()
    d_one = G.dev_empty(16 * nb)
    for k in range(NP):
        o = offs[3 * k * nb:3 * (k + 1) * nb + 1]
        c.me_search_dev(curs[k], refs[k], G.to_dev(grid), nb, G.to_dev(preds[o[0]:o[-1]]),
                        G.to_dev((o - o[0]).astype(np.uint32)), 3, p, d_one, None,
                        G.to_dev(thresh[k * nb:(k + 1) * nb]), 24)
        c.synchronize()
        np.testing.assert_array_equal(got[k * nb:(k + 1) * nb], G.from_dev(d_one, B.ME_RESULT_DTYPE)[:nb])
    for a, b in dpl:
        c.plane_free(a)
        c.plane_free(b)
