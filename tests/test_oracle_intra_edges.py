"""get_intra_edges (partition.rs:639-898) and its availability rules (recon_intra.rs): the oracle's
regenerated has_tr_* / has_bl_* bitmaps against the digests of all 44 reference tables, and the edge
gather against an independent numpy model + invariants."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from tests import oracle_lib as O

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
# BlockSize in enum order (partition.rs:130-153)
BSIZES = [(4, 4), (4, 8), (8, 4), (8, 8), (8, 16), (16, 8), (16, 16), (16, 32), (32, 16), (32, 32), (32, 64),
          (64, 32), (64, 64), (64, 128), (128, 64), (128, 128), (4, 16), (16, 4), (8, 32), (32, 8), (16, 64), (64, 16)]
MODES = ["DC_PRED", "V_PRED", "H_PRED", "D45_PRED", "D135_PRED", "D113_PRED", "D157_PRED", "D203_PRED", "D67_PRED",
         "SMOOTH_PRED", "SMOOTH_V_PRED", "SMOOTH_H_PRED", "PAETH_PRED", "UV_CFL_PRED"]
ANGLE = {"V_PRED": 90, "H_PRED": 180, "D45_PRED": 45, "D135_PRED": 135, "D113_PRED": 113, "D157_PRED": 157,
         "D203_PRED": 203, "D67_PRED": 67}


def L():
    l = O.lib()
    l.orc_intra_avail_table.argtypes = [C.c_int, C.c_int, C.c_void_p]
    l.orc_get_intra_edges.restype = None
    l.orc_get_intra_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t] + [C.c_int] * 22 + [C.POINTER(C.c_int)] * 2
    return l


def get_edges(img, po, tx, mode, bsize=None, part_bo=None, bxy=(0, 0), rect=None, dec=(0, 0), bd=8, ief=False,
              delta=0):
    """img: the region's pixels (rect origin at [0,0]); returns (edge[257], init_left, init_above)."""
    l = L()
    img = np.ascontiguousarray(img)
    h, w = img.shape
    rect = rect or (0, 0, w, h)
    bsize = bsize or tx
    part_bo = part_bo or (po[0] >> 2, po[1] >> 2)
    edge = np.full(257, 0xAA if img.dtype == np.uint8 else 0xAAAA, img.dtype)
    il, ia = C.c_int(), C.c_int()
    l.orc_get_intra_edges(edge.ctypes.data, img.ctypes.data, img.strides[0] // img.itemsize, img.itemsize,
                          rect[0] + w, rect[1] + h, rect[0], rect[1], rect[2], rect[3], dec[0], dec[1], part_bo[0],
                          part_bo[1], bxy[0], bxy[1], BSIZES.index(bsize), po[0], po[1], tx[0], tx[1], bd,
                          -1 if mode is None else MODES.index(mode), int(ief), delta, C.byref(il), C.byref(ia))
    return edge, il.value, ia.value


def test_every_availability_table_matches_the_reference_digest():
    l = L()
    buf = (C.c_uint8 * 128)()
    for kind, tag in ((0, "tr"), (1, "bl")):
        for i, (w, h) in enumerate(BSIZES):
            info = KATS["intra_avail_tables"][f"has_{tag}_{w}x{h}"]
            n = l.orc_intra_avail_table(kind, i, buf)
            got = bytes(buf[:n])
            assert n == info["n"] and list(got[:4]) == info["first4"], (tag, w, h)
            assert hashlib.sha256(got).hexdigest() == info["sha256"], (tag, w, h)


def model_dc_edges(img, x, y, n, bd):
    """independent model of the lookahead's call (lookahead.rs:59-74): DC_PRED, TX n x n"""
    base = 128 << (bd - 8)
    left = above = None
    if x != 0:                                       # needs_left = x != 0 for DC (p_angle = 0)
        left = img[y:y + n, x - 1][::-1].copy()      # bottom -> top
    # needs_top is always true for DC: `p_angle != 90 && p_angle < 180` holds for p_angle = 0
    # (partition.rs:695) - an asymmetry of the reference that is kept
    if y != 0:
        above = img[y - 1, x:x + n].copy()
    else:
        above = np.full(n, img[0, x - 1] if x != 0 else base - 1, img.dtype)
    return left, above, base


@pytest.mark.parametrize("bd,dtype", [(8, np.uint8), (10, np.uint16)])
def test_lookahead_dc_edges_match_the_model(bd, dtype):
    rng = np.random.default_rng(bd)
    img = rng.integers(0, 1 << bd, (48, 64)).astype(dtype)
    for by in range(6):
        for bx in range(8):
            x, y = 8 * bx, 8 * by
            # lookahead.rs:59-74 passes the importance-block index as the partition offset
            edge, il, ia = get_edges(img, (x, y), (8, 8), "DC_PRED", part_bo=(bx, by), bd=bd)
            left, above, base = model_dc_edges(img, x, y, 8, bd)
            assert il == (8 if x != 0 else 0) and ia == 8
            if left is not None:
                np.testing.assert_array_equal(edge[128 - 8:128], left)
            np.testing.assert_array_equal(edge[129:129 + 8], above)
            assert edge[128] == base                                    # DC needs no top-left


def test_edges_at_the_frame_border_use_the_documented_constants():
    img = np.arange(32 * 32, dtype=np.uint8).reshape(32, 32)
    e, il, ia = get_edges(img, (0, 0), (8, 8), "V_PRED")
    assert (il, ia) == (8, 8) and (e[120:128] == 129).all() and (e[129:137] == 127).all()
    e, il, ia = get_edges(img, (8, 0), (8, 8), "SMOOTH_PRED")          # y == 0, x != 0: above = dst[0][x-1]
    assert (e[129:137] == img[0, 7]).all() and (e[120:128] == img[0:8, 7][::-1]).all()
    e, il, ia = get_edges(img, (0, 8), (8, 8), "SMOOTH_PRED")          # x == 0, y != 0: left = dst[y-1][0]
    assert (e[120:128] == img[7, 0]).all() and (e[129:137] == img[7, 0:8]).all()
    e, _, _ = get_edges(img, (8, 8), (8, 8), "PAETH_PRED")
    assert e[128] == img[7, 7]


def test_rows_and_columns_past_the_visible_area_are_replicated():
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (20, 28)).astype(np.uint8)               # 28 x 20 visible, 16x16 blocks overhang
    e, il, ia = get_edges(img, (16, 16), (16, 16), "H_PRED")
    assert (e[128 - 4:128] == img[16:20, 15][::-1]).all() and (e[128 - 16:128 - 4] == img[19, 15]).all()
    e, il, ia = get_edges(img, (16, 16), (16, 16), "V_PRED")
    assert (e[129:129 + 12] == img[15, 16:28]).all() and (e[129 + 12:129 + 16] == img[15, 27]).all()


def test_directional_modes_extend_both_edges():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (128, 128)).astype(np.uint8)
    # 16x16 partition at (16,16) px: the top-right 16x16 (TR quadrant of the 32x32) is coded before it
    e, il, ia = get_edges(img, (16, 16), (16, 16), "D45_PRED")
    assert (il, ia) == (16, 32)
    # ... wait for the position: (16,16) is the BR quadrant of the 32x32 at the origin: its top-right lies in
    # the next 32x32 (later) -> replicated from the last above pixel
    assert (e[129 + 16:129 + 32] == img[15, 31]).all()
    e, il, ia = get_edges(img, (0, 16), (16, 16), "D45_PRED")          # BL quadrant: top-right is the TR quadrant
    assert (e[129 + 16:129 + 32] == img[15, 16:32]).all()
    e, il, ia = get_edges(img, (16, 0 + 32), (16, 16), "D203_PRED")    # (16,32): TR of the 32x32 at (0,32)
    assert (il, ia) == (32, 16)
    assert (e[128 - 32:128 - 16] == img[48:64, 15][::-1]).all() or (e[128 - 32:128 - 16] == img[47, 15]).all()
    # top-left smoothing for the 90 < angle < 180 zone on large blocks (partition.rs:886-892)
    e0, _, _ = get_edges(img, (16, 16), (16, 16), "D135_PRED", ief=False)
    e1, _, _ = get_edges(img, (16, 16), (16, 16), "D135_PRED", ief=True)
    l, a, tl = int(e0[127]), int(e0[129]), int(e0[128])
    assert tl == img[15, 15] and int(e1[128]) == (l * 5 + tl * 6 + a * 5 + 8) >> 4
