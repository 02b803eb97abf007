"""GPU parity for the lookahead kernels (rav1e_b200/csrc/lookahead.cu) == oracle/lookahead.c:
Plane::downsampled (+ pad), estimate_intra_costs, estimate_inter_costs (cost part, f64 mean bit
exact), estimate_importance_block_difference; 1080p size-independent properties."""
import ctypes as C

import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests import oracle_lib as O
from tests.test_oracle_lookahead import L as OL, downsample as oracle_downsample, padded

pytestmark = pytest.mark.gpu


def download_padded(c, p, dtype):
    """the whole padded area of a device plane"""
    full = np.zeros((p.height + 2 * p.pad, p.stride), dtype)
    base = p.data - (p.pad * p.stride) * full.itemsize - (p.data - p.alloc - p.pad * p.stride * full.itemsize) % (p.stride * full.itemsize)
    c.check(c.L.b200_memcpy_d2h(c.h, full.ctypes.data, base, full.nbytes))
    lead = (p.data - base) // full.itemsize - p.pad * p.stride
    return full[:, lead - p.pad:lead + p.width + p.pad]


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10)])
@pytest.mark.parametrize("w,h", [(128, 72), (127, 71), (90, 33)])
def test_downsample_matches_oracle(dtype, bd, w, h):
    c = G.ctx()
    rng = np.random.default_rng(w + h + bd)
    img = rng.integers(0, 1 << bd, (h, w)).astype(dtype)
    src = c.plane_from_host(img, 8)
    w2, h2 = (w + 1) // 2, (h + 1) // 2
    for pad_w, pad_h in ((w2, h2), (w2 - 1, h2 - 1)):
        dst = c.plane_alloc(w2, h2, 12, img.itemsize)
        c.plane_downsample_dev(src, dst, pad_w, pad_h)
        c.synchronize()
        got = download_padded(c, dst, dtype)
        np.testing.assert_array_equal(got, oracle_downsample(img, 12, pad_w, pad_h))
        c.plane_free(dst)
    c.plane_free(src)


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10), (np.uint16, 12)])
def test_intra_and_inter_costs_match_oracle(dtype, bd):
    c = G.ctx()
    W, H, PAD = 352, 200, 96
    cur, ref = G.make_planes(W, H, PAD, dtype, seed=bd, bit_depth=bd, shift=(5, -3))
    ocur, dcur = G.both_planes(cur, PAD)
    oref, dref = G.both_planes(ref, PAD)
    l = OL()
    wb, hb = W // 8, H // 8
    want = np.zeros((hb, wb), np.uint32)
    isz = cur.itemsize
    l.orc_estimate_intra_costs(ocur.at(0, 0), ocur.stride, W, H, isz, bd, want.ctypes.data)
    d_costs = G.dev_empty(4 * wb * hb)
    c.estimate_intra_costs_dev(dcur, bd, d_costs)
    c.synchronize()
    np.testing.assert_array_equal(G.from_dev(d_costs, np.uint32)[:wb * hb].reshape(hb, wb), want)
    # inter: vectors reaching into the padding on every side, sub-pel parts of both signs
    rng = np.random.default_rng(2)
    mvs = rng.integers(-(PAD - 8) * 8, (PAD - 8) * 8, (hb, wb, 2)).astype(np.int16)
    want_c = np.zeros((hb, wb), np.uint32)
    want_m = l.orc_estimate_inter_costs(ocur.at(0, 0), ocur.stride, oref.at(0, 0), oref.stride, W, H, isz,
                                        mvs.ctypes.data, want_c.ctypes.data)
    d_c, d_s, d_m = G.dev_empty(4 * wb * hb), G.dev_empty(8), G.dev_empty(8)
    c.estimate_inter_costs_dev(dcur, dref, G.to_dev(mvs), d_c, d_s, d_m)
    c.synchronize()
    np.testing.assert_array_equal(G.from_dev(d_c, np.uint32)[:wb * hb].reshape(hb, wb), want_c)
    assert G.from_dev(d_m, np.float64)[0] == want_m                       # bit exact (one rounded division)
    want_d = l.orc_importance_block_difference(ocur.at(0, 0), ocur.stride, oref.at(0, 0), oref.stride, W, H, isz)
    c.importance_block_difference_dev(dcur, dref, d_s, d_m)
    c.synchronize()
    assert G.from_dev(d_m, np.float64)[0] == want_d
    for p in (dcur, dref):
        c.plane_free(p)


def test_1080p_pyramid_and_costs_properties():
    """size-independent properties at the BASELINE frame size: a constant plane stays constant through
    both pyramid levels and has zero intra cost away from the frame's first block; a plane against
    itself has zero inter cost; searching on the pyramid finds the scaled shift."""
    c = G.ctx()
    W, H, PAD = 1920, 1080, 96
    flat = np.full((H, W), 77, np.uint8)
    p0 = c.plane_from_host(flat, PAD)
    p1 = c.plane_alloc(960, 540, PAD // 2, 1)
    p2 = c.plane_alloc(480, 270, PAD // 4, 1)
    c.plane_downsample_dev(p0, p1, 960, 540)
    c.plane_downsample_dev(p1, p2, 480, 270)
    c.synchronize()
    assert (download_padded(c, p2, np.uint8) == 77).all()
    d_costs = G.dev_empty(4 * 240 * 135)
    c.estimate_intra_costs_dev(p0, 8, d_costs)
    c.synchronize()
    costs = G.from_dev(d_costs, np.uint32)[:240 * 135].reshape(135, 240)
    assert costs[0, 0] == ((128 - 77) * 64 + 4) >> 3 and (costs.reshape(-1)[1:] == 0).all()
    d_s, d_m = G.dev_empty(8), G.dev_empty(8)
    c.estimate_inter_costs_dev(p0, p0, G.to_dev(np.zeros((135, 240, 2), np.int16)), None, d_s, d_m)
    c.synchronize()
    assert G.from_dev(d_m, np.float64)[0] == 0.0
    for p in (p0, p1, p2):
        c.plane_free(p)
