"""b200_subpel_rdo_dev (sub-pel refinement fused with the winner's residual + forward transform) ==
oracle: get_subpel_mv_rd per candidate (8-tap MC + SAD / SATD + mv cost), first minimum per block, then
put_8tap of the winner -> diff -> forward_transform."""
import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests import oracle_lib as O

pytestmark = pytest.mark.gpu

CASES = [
    # (n, dtype, bd, use_satd, filter_mode, tx_type)
    (16, np.uint8, 8, True, 0, 0),
    (16, np.uint16, 10, True, 0, 0),       # BASELINE configs[3]: 10 bit, REGULAR, SATD, DCT_DCT 16x16
    (16, np.uint16, 12, False, 2, 1),
    (8, np.uint8, 8, True, 1, 3),
    (8, np.uint16, 10, False, 0, 9),
    (32, np.uint8, 8, True, 0, 0),
    (32, np.uint16, 10, True, 3, 9),
]


@pytest.mark.parametrize("n,dtype,bd,use_satd,mode,tx_type", CASES)
def test_fused_subpel_rdo_matches_oracle(n, dtype, bd, use_satd, mode, tx_type):
    import torch
    c = G.ctx()
    W, H, PAD = 256, 160, 96
    cur, ref = G.make_planes(W, H, PAD, dtype, seed=n + bd, bit_depth=bd, shift=(2, -1))
    ocur, dcur = G.both_planes(cur, PAD)
    oref, dref = G.both_planes(ref, PAD)
    blocks = G.grid_blocks(W, H, n, n)
    nb = len(blocks)
    cands, offs = G.random_cands(nb, 9, 6, seed=5, fullpel=False)     # sub-pel vectors, 4..13 per block
    # a few vectors far outside the frame: dropped by the mv-range check (some blocks lose all of them)
    cands["mv_col"][::17] = 16000
    rng = np.random.default_rng(3)
    pmv = (rng.integers(-16, 17, (nb, 4)) * 2).astype(np.int16)
    lam = 700
    ts = {8: 1, 16: 2, 32: 3}[n]
    want_sad, want_cost = O.subpel_candidates(ocur, oref, blocks, cands, n, n, use_satd, lam, pmv, filter_mode=mode,
                                              bit_depth=bd)
    p = B.me_params(n, n, W, H, lam, use_satd=use_satd, bit_depth=bd)
    d_sad, d_cost, d_best = G.dev_empty(4 * len(cands)), G.dev_empty(8 * len(cands)), G.dev_empty(16 * nb)
    ct = torch.int16 if bd == 8 else torch.int32
    d_coef = torch.full((nb, n * n), 77, dtype=ct, device="cuda")
    c.subpel_rdo_dev(dcur, dref, G.to_dev(blocks), nb, G.to_dev(cands), len(cands), G.to_dev(offs), p, mode, ts,
                     tx_type, G.to_dev(pmv), d_sad, d_cost, d_best, d_coef)
    c.synchronize()
    np.testing.assert_array_equal(G.from_dev(d_sad, np.uint32)[:len(cands)], want_sad)
    np.testing.assert_array_equal(G.from_dev(d_cost, np.uint64)[:len(cands)], want_cost)
    best = G.from_dev(d_best, B.ME_RESULT_DTYPE)[:nb]
    coef = d_coef.cpu().numpy()
    EMPTY = np.uint64(2**64 - 1)
    resid = np.zeros((nb, n, n), np.int16)
    have = np.zeros(nb, bool)
    for b in range(nb):
        lo, hi = int(offs[b]), int(offs[b + 1])
        if lo == hi or want_cost[lo:hi].min() == EMPTY:
            assert best[b]["cost"] == EMPTY
            continue
        k = lo + int(np.argmin(want_cost[lo:hi]))
        assert best[b]["cost"] == want_cost[k] and best[b]["sad"] == want_sad[k]
        assert (best[b]["mv_row"], best[b]["mv_col"]) == (cands[k]["mv_row"], cands[k]["mv_col"])
        mr, mc = int(cands[k]["mv_row"]), int(cands[k]["mv_col"])
        x, y = int(blocks[b]["x"]), int(blocks[b]["y"])
        pred = O.put_8tap(oref, x + (mc >> 3), y + (mr >> 3), n, n, (mc << 1) & 15, (mr << 1) & 15, mode, mode, bd)
        resid[b] = cur[y:y + n, x:x + n].astype(np.int32) - pred.astype(np.int32)
        have[b] = True
    want_coef = O.forward_transform_batch(resid, ts, tx_type, bd, coeff_i32=(bd > 8)).reshape(nb, n * n)
    np.testing.assert_array_equal(coef[have], want_coef[have])
    assert (coef[~have] == 0).all()
    for pl in (dcur, dref):
        c.plane_free(pl)


@pytest.mark.parametrize("n,dtype,bd,use_satd,allow_hp", [(16, np.uint8, 8, True, False), (16, np.uint16, 10, True, True),
                                                           (8, np.uint8, 8, False, True), (32, np.uint16, 10, True, False)])
def test_subpel_diamond_search_matches_oracle(n, dtype, bd, use_satd, allow_hp):
    """b200_subpel_search_dev == the oracle's subpel_diamond_search (me.rs:1311-1383) from the same
    full-pel starting results, incl. the transform of the final vector's residual."""
    import ctypes as C
    import torch
    c = G.ctx()
    W, H, PAD = 256, 160, 96
    cur, ref = G.make_planes(W, H, PAD, dtype, seed=3 * n + bd, bit_depth=bd, shift=(2, -1))
    ocur, dcur = G.both_planes(cur, PAD)
    oref, dref = G.both_planes(ref, PAD)
    blocks = G.grid_blocks(W, H, n, n)
    nb = len(blocks)
    rng = np.random.default_rng(8)
    lam = 500
    # starting points: full-pel vectors near the true shift with their SAD-based cost (the full-pel stage's winner)
    start = np.zeros(nb, B.ME_RESULT_DTYPE)
    cands = np.zeros(nb, B.CAND_DTYPE)
    cands["block"] = np.arange(nb)
    cands["mv_row"] = (rng.integers(-3, 2, nb)) * 8
    cands["mv_col"] = (rng.integers(0, 5, nb)) * 8
    pmv = (rng.integers(-8, 9, (nb, 4)) * 2).astype(np.int16)
    sad, cost = O.fullpel_candidates(ocur, oref, blocks, cands, n, n, False, lam, pmv, allow_hp=allow_hp)
    start["cost"], start["sad"], start["mv_row"], start["mv_col"] = cost, sad, cands["mv_row"], cands["mv_col"]
    want = start.copy()
    L = O.lib()
    L.orc_subpel_diamond_search_blocks.restype = None
    L.orc_subpel_diamond_search_blocks.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t] + [C.c_int] * 3 + \
        [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p] + [C.c_int] * 3 + [C.c_void_p, C.c_int]
    L.orc_subpel_diamond_search_blocks(ocur.origin_ptr(), ocur.stride, oref.origin_ptr(), oref.stride, ocur.bpp,
                                       2 * ((W + 7) >> 3), 2 * ((H + 7) >> 3), O.ptr(blocks), nb, n, n, int(use_satd), lam,
                                       O.ptr(pmv), int(allow_hp), 0, bd, O.ptr(want), 0)
    ts = {8: 1, 16: 2, 32: 3}[n]
    p = B.me_params(n, n, W, H, lam, allow_hp=allow_hp, use_satd=use_satd, bit_depth=bd)
    d_best = G.dev_empty(16 * nb)
    d_coef = torch.zeros((nb, n * n), dtype=torch.int16 if bd == 8 else torch.int32, device="cuda")
    c.subpel_search_dev(dcur, dref, G.to_dev(blocks), nb, G.to_dev(start), p, d_best, 0, ts, 0, G.to_dev(pmv), d_coef)
    c.synchronize()
    got = G.from_dev(d_best, B.ME_RESULT_DTYPE)[:nb]
    for f in ("cost", "sad", "mv_row", "mv_col"):
        np.testing.assert_array_equal(got[f], want[f], err_msg=f)
    assert (got["mv_row"] != start["mv_row"]).any() or (got["mv_col"] != start["mv_col"]).any()   # the search moved
    resid = np.zeros((nb, n, n), np.int16)
    for b in range(nb):
        mr, mc = int(want[b]["mv_row"]), int(want[b]["mv_col"])
        x, y = int(blocks[b]["x"]), int(blocks[b]["y"])
        pred = O.put_8tap(oref, x + (mc >> 3), y + (mr >> 3), n, n, (mc << 1) & 15, (mr << 1) & 15, 0, 0, bd)
        resid[b] = cur[y:y + n, x:x + n].astype(np.int32) - pred.astype(np.int32)
    want_coef = O.forward_transform_batch(resid, ts, 0, bd, coeff_i32=(bd > 8)).reshape(nb, n * n)
    np.testing.assert_array_equal(d_coef.cpu().numpy(), want_coef)
    for pl in (dcur, dref):
        c.plane_free(pl)
