"""GPU parity: CUDA put_8tap / prep_8tap / mc_avg == oracle, bit exact — the CUDA counterpart of
the reference's asm==rust tests (asm/x86/mc.rs:624-833: 10 filter pairs x 4 MVs on 8x8),
extended to every block size the encoder uses, 8/10/12 bit and chroma decimation."""
import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests import oracle_lib as O

pytestmark = pytest.mark.gpu

SIZES = [(2, 2), (4, 4), (4, 8), (8, 4), (8, 8), (16, 16), (16, 8), (32, 32), (64, 64), (128, 128),
         (64, 16), (8, 32)]


def planes(dtype, bd, seed):
    rng = np.random.default_rng(seed)
    W, H, PAD = 256, 160, 160
    img = rng.integers(0, 1 << bd, (H, W)).astype(dtype)
    return (W, H, PAD), *G.both_planes(img, PAD)


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10), (np.uint16, 12)])
def test_percall_all_filter_pairs(dtype, bd):
    """b200_put_8tap / b200_prep_8tap / b200_mc_avg (host pointers, byte strides)."""
    L = B.lib()
    (W, H, PAD), op, dp = planes(dtype, bd, 1)
    isz = op.data.itemsize
    for mode_x in range(4):
        for mode_y in range(4):
            for cf, rf in ((0, 0), (0, 9), (6, 0), (11, 3)):
                w, h = 8, 8
                want = O.put_8tap(op, 20, 24, w, h, cf, rf, mode_x, mode_y, bd)
                got = np.zeros((h, w), dtype)
                L.b200_put_8tap(got.ctypes.data, w * isz, op.at(20, 24), op.stride * isz, w, h, cf, rf,
                                mode_x, mode_y, bd)
                np.testing.assert_array_equal(got, want)
                wantp = O.prep_8tap(op, 20, 24, w, h, cf, rf, mode_x, mode_y, bd)
                gotp = np.zeros((h, w), np.int16)
                L.b200_prep_8tap(gotp.ctypes.data, op.at(20, 24), op.stride * isz, w, h, cf, rf,
                                 mode_x, mode_y, bd)
                np.testing.assert_array_equal(gotp, wantp)
    t1 = O.prep_8tap(op, 20, 24, 16, 8, 3, 5, 0, 0, bd)
    t2 = O.prep_8tap(op, 33, 41, 16, 8, 12, 0, 1, 2, bd)
    got = np.zeros((8, 16), dtype)
    L.b200_mc_avg(got.ctypes.data, 16 * isz, t1.ctypes.data, t2.ctypes.data, 16, 8, bd)
    np.testing.assert_array_equal(got, O.mc_avg(t1, t2, bd))
    G.ctx().plane_free(dp)


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10), (np.uint16, 12)])
@pytest.mark.parametrize("w,h", SIZES)
def test_batched_blocks_match_oracle(dtype, bd, w, h):
    import torch
    c = G.ctx()
    (W, H, PAD), op, dp = planes(dtype, bd, w * 7 + h)
    blocks = G.grid_blocks(W, H, max(w, 8), max(h, 8))[:96]
    n = len(blocks)
    rng = np.random.default_rng(5)
    mvs = rng.integers(-60 * 8, 60 * 8 + 1, (n, 2)).astype(np.int16)
    special = [(0, 0), (8, -16), (0, 3), (-5, 0)]      # copy, full-pel, H only, V only
    if n < 8:                  # 128x128 yields only two blocks on this plane: repeat them
        blocks = np.ascontiguousarray(np.tile(blocks, 4))
        n = len(blocks)
        mvs = rng.integers(-60 * 8, 60 * 8 + 1, (n, 2)).astype(np.int16)
    mvs[:4] = special
    tdt = torch.uint8 if dtype == np.uint8 else torch.int16
    for mode_x, mode_y, xdec, ydec in ((0, 0, 0, 0), (1, 2, 0, 0), (3, 3, 0, 0), (0, 0, 1, 1)):
        for kind in (0, 1):
            want = O.mc_blocks(op, blocks, mvs, w, h, mode_x, mode_y, bd, xdec, ydec, kind)
            d_out = torch.empty((n, h, w), dtype=tdt if kind == 0 else torch.int16, device="cuda")
            c.mc_blocks_dev(dp, G.to_dev(blocks), G.to_dev(mvs), n, w, h, mode_x, mode_y, bd, xdec,
                            ydec, kind, d_out)
            c.synchronize()
            got = d_out.cpu().numpy().view(want.dtype)
            np.testing.assert_array_equal(got, want, err_msg=f"{mode_x}{mode_y} dec{xdec}{ydec} kind{kind}")
    # compound: avg of two preps
    mv2 = rng.integers(-40 * 8, 40 * 8 + 1, (n, 2)).astype(np.int16)
    t1 = O.mc_blocks(op, blocks, mvs, w, h, 0, 0, bd, kind=1)
    t2 = O.mc_blocks(op, blocks, mv2, w, h, 0, 0, bd, kind=1)
    want = np.stack([O.mc_avg(a, b, bd) for a, b in zip(t1, t2)])
    d_dst = torch.empty((n, h, w), dtype=tdt, device="cuda")
    c.mc_avg_dev(torch.from_numpy(t1).cuda(), torch.from_numpy(t2).cuda(), d_dst, n, w, h, bd)
    c.synchronize()
    np.testing.assert_array_equal(d_dst.cpu().numpy().view(want.dtype), want)
    c.plane_free(dp)


def test_preconditions_rejected():
    """mc.rs:256-257: odd heights and non-power-of-two widths are assertion failures."""
    import torch
    c = G.ctx()
    (W, H, PAD), op, dp = planes(np.uint8, 8, 0)
    blocks = G.grid_blocks(W, H, 16, 16)[:2]
    d_out = torch.empty(4096, dtype=torch.uint8, device="cuda")
    for w, h in ((12, 8), (8, 7), (256, 8)):
        with pytest.raises(B.B200Error) as e:
            c.mc_blocks_dev(dp, G.to_dev(blocks), None, 2, w, h, 0, 0, 8, 0, 0, 0, d_out)
        assert e.value.status == B.ERR_ARG
    c.plane_free(dp)
