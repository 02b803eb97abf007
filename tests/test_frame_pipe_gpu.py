"""b200_frame_pipe (host buffers, one call per frame): each pushed frame is searched against the
previously pushed one; winners, coefficients and the quantize chain must equal the oracle run on the
same frames and the same expanded candidate lists."""
import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests import oracle_lib as O
from tests.test_oracle_quantize import chain as oracle_chain

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("quant", [False, True])
@pytest.mark.parametrize("use_centers", [False, True])
def test_pushed_frames_match_oracle(quant, use_centers):
    W, H, PAD = 320, 176, 96
    NS, NT, LAM = 24, 6, 900
    c = B.Context(0)
    pipe = B.FramePipe(c, W, H, PAD, (16, 16), LAM, NS, NT, 40, tx_size=2, tx_type=0,
                       dc_quant=88 if quant else 0, ac_quant=72 if quant else 0)
    nb = pipe.nblocks
    blocks = G.grid_blocks(W, H, 16, 16)
    assert nb == len(blocks)
    rng = np.random.default_rng(7)
    frames = []
    for f in range(3):
        cur, ref = G.make_planes(W, H, PAD, np.uint8, seed=20 + f, shift=(3 + f, -2))
        frames.append(cur)
    pipe.push(frames[0])                                   # nothing to search against yet
    for f in (1, 2):
        so = rng.integers(-40, 41, (nb, NS, 2)).astype(np.int8)
        to = rng.integers(-40, 41, (nb, NT, 2)).astype(np.int8)
        cen = (rng.integers(-6, 7, (nb, 2)) * 8).astype(np.int16) if use_centers else None
        best, best2 = np.zeros(nb, B.ME_RESULT_DTYPE), np.zeros(nb, B.ME_RESULT_DTYPE)
        coef = np.zeros((nb, 256), np.int16)
        eob, dist = np.zeros(nb, np.uint16), np.zeros(nb, np.uint64)
        pipe.push(frames[f], so, to, cen, best, best2, coef, eob if quant else None, dist if quant else None)
        c.synchronize()
        ocur, oref = O.Plane(W, H, PAD), O.Plane(W, H, PAD)
        ocur.fill_from(frames[f])
        oref.fill_from(frames[f - 1])
        for offs, per, satd, got in ((so, NS, False, best), (to, NT, True, best2)):
            cands = np.zeros(nb * per, B.CAND_DTYPE)
            cands["block"] = np.repeat(np.arange(nb, dtype=np.uint32), per)
            base = np.zeros((nb, 2), np.int16) if cen is None else cen
            mv = (base[:, None, :].astype(np.int32) + 8 * offs.astype(np.int32)).astype(np.int16).reshape(-1, 2)
            cands["mv_row"], cands["mv_col"] = mv[:, 0], mv[:, 1]
            sad, cost = O.fullpel_candidates(ocur, oref, blocks, cands, 16, 16, satd, LAM)
            for b in range(nb):
                k = b * per + int(np.argmin(cost[b * per:(b + 1) * per]))
                assert got[b]["cost"] == cost[k] and got[b]["sad"] == sad[k]
                assert (got[b]["mv_row"], got[b]["mv_col"]) == (cands[k]["mv_row"], cands[k]["mv_col"])
        # transform leg: residual against the reference displaced by the SAD winner
        resid = np.zeros((nb, 16, 16), np.int16)
        for i, b in enumerate(blocks):
            x, y = int(b["x"]) + PAD, int(b["y"]) + PAD
            dx, dy = int(best[i]["mv_col"]) // 8, int(best[i]["mv_row"]) // 8   # full-pel vectors: exact
            resid[i] = ocur.data[y:y + 16, x:x + 16].astype(np.int32) - oref.data[y + dy:y + dy + 16, x + dx:x + dx + 16]
        want = O.forward_transform_batch(resid, 2, 0, 8).reshape(nb, 256)
        if quant:
            q, _, we, wd = oracle_chain(want, 2, 0, 88, 72, False)
            np.testing.assert_array_equal(coef, q)
            np.testing.assert_array_equal(eob, we)
            np.testing.assert_array_equal(dist, wd)
        else:
            np.testing.assert_array_equal(coef, want)
    pipe.close()
    c.close()


@pytest.mark.parametrize("bd", [8, 10])
@pytest.mark.parametrize("use_centers", [False, True])
def test_resident_lists_and_packed_output(bd, use_centers):
    """b200_frame_pipe_set_lists + b200_frame_pipe_push_packed: the lists go up once, a push sends the frame only,
    and the quantized coefficients come back as coeffs[scan[0..eob)] per block.  Everything must equal the dense
    push of the same frames with the same lists (which the test above pins to the oracle)."""
    from tests.test_oracle_quantize import L as OL
    W, H, PAD = 320, 176, 96
    NS, NT, LAM = 24, 6, 900
    dtype = np.uint8 if bd == 8 else np.uint16
    ct = np.int16 if bd == 8 else np.int32
    dcq, acq = 88 << (bd - 8), 72 << (bd - 8)
    c = B.Context(0)
    mk = lambda: B.FramePipe(c, W, H, PAD, (16, 16), LAM, NS, NT, 40, tx_size=2, tx_type=0, dc_quant=dcq, ac_quant=acq,
                             bit_depth=bd)
    dense, packed_pipe, small_pipe = mk(), mk(), mk()
    nb = dense.nblocks
    rng = np.random.default_rng(11)
    frames = [G.make_planes(W, H, PAD, dtype, seed=40 + f, shift=(2 + f, -1), bit_depth=bd)[0] for f in range(4)]
    so = rng.integers(-40, 41, (nb, NS, 2)).astype(np.int8)
    to = rng.integers(-40, 41, (nb, NT, 2)).astype(np.int8)
    cen = (rng.integers(-6, 7, (nb, 2)) * 8).astype(np.int16) if use_centers else None
    packed_pipe.set_lists(so, to, cen)
    small_pipe.set_lists(so, to, cen)
    scan = np.zeros(256, np.uint16)
    OL().orc_scan_order(2, 0, scan.ctypes.data, None)
    dense.push(frames[0])
    assert packed_pipe.push_packed(frames[0]) == 0             # the first push only uploads
    assert small_pipe.push_packed(frames[0]) == 0
    for f in (1, 2, 3):
        b1, b2 = np.zeros(nb, B.ME_RESULT_DTYPE), np.zeros(nb, B.ME_RESULT_DTYPE)
        q = np.zeros((nb, 256), ct)
        e1, d1 = np.zeros(nb, np.uint16), np.zeros(nb, np.uint64)
        dense.push(frames[f], so, to, cen, b1, b2, q, e1, d1)
        g1, g2 = np.zeros(nb, B.ME_RESULT_DTYPE), np.zeros(nb, B.ME_RESULT_DTYPE)
        e2, d2 = np.zeros(nb, np.uint16), np.zeros(nb, np.uint64)
        buf = np.full(nb * 256, -1, ct)
        total = packed_pipe.push_packed(frames[f], g1, g2, e2, d2, buf)
        c.synchronize()
        np.testing.assert_array_equal(g1, b1)
        np.testing.assert_array_equal(g2, b2)
        np.testing.assert_array_equal(e2, e1)
        np.testing.assert_array_equal(d2, d1)
        assert total == int(e1.astype(np.int64).sum()) and 0 < total <= nb * 256
        want = np.concatenate([q[b][scan[:e1[b]]] for b in range(nb)])
        np.testing.assert_array_equal(buf[:total], want)
        assert (buf[total:] == -1).all()                       # nothing beyond the total is written
        # a buffer that is too small: the total is still reported, the prefix that fits is copied
        small = np.full(100, -1, ct)
        assert small_pipe.push_packed(frames[f], packed=small) == total
        c.synchronize()
        k = min(100, total)
        np.testing.assert_array_equal(small[:k], want[:k])
        assert (small[k:] == -1).all()
    # per-push lists still work on a pipe that has resident ones, and replace them
    b1 = np.zeros(nb, B.ME_RESULT_DTYPE)
    packed_pipe.push(frames[1], so, to, cen, b1, None, None, None, None)
    c.synchronize()
    with pytest.raises(B.B200Error):
        packed_pipe.push_packed(frames[2])                      # resident lists are gone
    for pp in (dense, packed_pipe, small_pipe):
        pp.close()
    c.close()
