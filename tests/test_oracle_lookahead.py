"""Lookahead kernels of the oracle (Plane::downsampled + pad, estimate_intra_costs,
estimate_importance_block_difference, estimate_inter_costs' cost part; api/lookahead.rs) against
independent numpy models.  No stored vectors exist upstream for these (parity unpinned)."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib as O

H8 = np.array([[1]], dtype=np.int64)
for _ in range(3):
    H8 = np.block([[H8, H8], [H8, -H8]])


def L():
    l = O.lib()
    vp, pd, i32 = C.c_void_p, C.c_ssize_t, C.c_int
    l.orc_plane_downsample.restype = None
    l.orc_plane_downsample.argtypes = [vp, pd, i32, i32, vp, pd, i32, i32, i32, i32]
    l.orc_estimate_intra_costs.restype = None
    l.orc_estimate_intra_costs.argtypes = [vp, pd, i32, i32, i32, i32, vp]
    l.orc_importance_block_difference.restype = C.c_double
    l.orc_importance_block_difference.argtypes = [vp, pd, vp, pd, i32, i32, i32]
    l.orc_estimate_inter_costs.restype = C.c_double
    l.orc_estimate_inter_costs.argtypes = [vp, pd, vp, pd, i32, i32, i32, vp, vp]
    return l


def satd8(a, b):
    d = a.astype(np.int64) - b.astype(np.int64)
    return (int(np.abs(H8 @ d @ H8.T).sum()) + 4) >> 3


def padded(img, pad):
    return np.pad(img, pad, mode="edge")


def downsample(img, dst_pad, pad_w=None, pad_h=None):
    """oracle call on an edge-replicated copy of img; returns the whole padded destination"""
    h, w = img.shape
    src = padded(img, 2)
    w2, h2 = (w + 1) // 2, (h + 1) // 2
    dst = np.zeros((h2 + 2 * dst_pad, w2 + 2 * dst_pad), img.dtype)
    isz = img.itemsize
    L().orc_plane_downsample(src.ctypes.data + (2 * src.shape[1] + 2) * isz, src.shape[1], w, h,
                             dst.ctypes.data + (dst_pad * dst.shape[1] + dst_pad) * isz, dst.shape[1], dst_pad, isz,
                             pad_w or w2, pad_h or h2)
    return dst


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 12)])
@pytest.mark.parametrize("w,h", [(64, 48), (63, 47), (30, 17)])
def test_downsample_is_the_rounded_box_filter_with_replicated_borders(dtype, bd, w, h):
    rng = np.random.default_rng(w * h + bd)
    img = rng.integers(0, 1 << bd, (h, w)).astype(dtype)
    got = downsample(img, 5)
    e = padded(img, 1).astype(np.int64)[1:, 1:]            # one replicated row / column for odd sizes
    w2, h2 = (w + 1) // 2, (h + 1) // 2
    box = (e[0:2 * h2:2, 0:2 * w2:2] + e[0:2 * h2:2, 1:2 * w2:2] + e[1:2 * h2:2, 0:2 * w2:2] + e[1:2 * h2:2, 1:2 * w2:2] + 2) >> 2
    np.testing.assert_array_equal(got, padded(box.astype(dtype), 5))


def test_downsample_pads_from_the_frame_size_not_the_plane_size():
    """quarter resolution of a 1917-wide frame: plane width ((1917+1)/2+1)/2 = 480 but Plane::pad
    replicates from (1917 + 2) >> 2 = 479 columns (v_frame pad(): (w + xdec) >> xdec)"""
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (10, 959)).astype(np.uint8)          # the half-resolution plane
    got = downsample(img, 3, pad_w=479, pad_h=5)
    assert got.shape == (5 + 6, 480 + 6)
    vis = got[3:8, 3:483]
    assert (vis[:, 479] == vis[:, 478]).all()                       # column 479 is a replica
    assert (got[:, 3 + 479:] == got[:, 3 + 478:3 + 479]).all()


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10)])
def test_intra_costs_match_the_model(dtype, bd):
    rng = np.random.default_rng(bd)
    W, H = 88, 52                                                   # 11 x 6 importance blocks (4 rows ignored)
    img = (rng.integers(0, 1 << bd, (H, W)) // 3 + np.linspace(0, (1 << bd) // 2, W)[None, :]).astype(dtype)
    costs = np.zeros((H // 8, W // 8), np.uint32)
    L().orc_estimate_intra_costs(img.ctypes.data, W, W, H, img.itemsize, bd, costs.ctypes.data)
    for by in range(H // 8):
        for bx in range(W // 8):
            x, y = 8 * bx, 8 * by
            blk = img[y:y + 8, x:x + 8]
            left = img[y:y + 8, x - 1].astype(np.int64) if x else None
            above = img[y - 1, x:x + 8].astype(np.int64) if y else None
            if x and y:
                dc = (left.sum() + above.sum() + 8) // 16
            elif x:
                dc = (left.sum() + 4) // 8
            elif y:
                dc = (above.sum() + 4) // 8
            else:
                dc = 128 << (bd - 8)
            assert costs[by, bx] == satd8(blk, np.full((8, 8), dc)), (bx, by)


def test_importance_block_difference_and_inter_costs():
    rng = np.random.default_rng(5)
    W, H, PAD = 96, 64, 24
    a = rng.integers(0, 256, (H, W)).astype(np.uint8)
    b = np.clip(a.astype(np.int64) + rng.integers(-20, 21, (H, W)), 0, 255).astype(np.uint8)
    l = L()
    got = l.orc_importance_block_difference(a.ctypes.data, W, b.ctypes.data, W, W, H, 1)
    tot = 0
    for by in range(H // 8):
        for bx in range(W // 8):
            sa, sb = int(a[8 * by:8 * by + 8, 8 * bx:8 * bx + 8].sum()), int(b[8 * by:8 * by + 8, 8 * bx:8 * bx + 8].sum())
            tot += abs((sa + 32) // 64 - (sb + 32) // 64)
    assert got == tot / ((W // 8) * (H // 8))
    # inter costs: vectors of both signs incl. sub-pel parts; (x * 64 + mv) / 8 truncates toward zero
    bp = padded(b, PAD)
    mvs = (rng.integers(-PAD * 8 + 8, PAD * 8 - 8, (H // 8, W // 8, 2))).astype(np.int16)
    costs = np.zeros((H // 8, W // 8), np.uint32)
    got = l.orc_estimate_inter_costs(a.ctypes.data, W, bp.ctypes.data + PAD * bp.shape[1] + PAD, bp.shape[1], W, H, 1,
                                     mvs.ctypes.data, costs.ctypes.data)
    tot = 0
    for by in range(H // 8):
        for bx in range(W // 8):
            rx = int(np.trunc((bx * 64 + int(mvs[by, bx, 1])) / 8))
            ry = int(np.trunc((by * 64 + int(mvs[by, bx, 0])) / 8))
            want = satd8(a[8 * by:8 * by + 8, 8 * bx:8 * bx + 8], bp[PAD + ry:PAD + ry + 8, PAD + rx:PAD + rx + 8])
            assert costs[by, bx] == want
            tot += want
    assert got == tot / ((W // 8) * (H // 8))
