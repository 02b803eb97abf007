"""GPU parity for b200_get_intra_edges_dev == oracle get_intra_edges (partition.rs:639-898): every
prediction mode (and None), angle deltas, edge filter flag, transform sizes inside partitions, luma and
decimated chroma planes, a tile region inside the plane, blocks overhanging the visible area; then the
edges feed b200_predict_intra_dev exactly like the reference's buffers would."""
import ctypes as C

import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests import oracle_lib as O
from tests.test_oracle_intra_edges import BSIZES, L as OL

pytestmark = pytest.mark.gpu

TX = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (4, 8), (8, 4), (8, 16), (16, 8), (16, 32), (32, 16), (32, 64),
      (64, 32), (4, 16), (16, 4), (8, 32), (32, 8), (16, 64), (64, 16)]


def oracle_edges(img_full, rect, dec, bd, it):
    l = OL()
    rx, ry, rw, rh = rect
    reg = img_full[ry:, rx:]
    edge = np.zeros(257, img_full.dtype)
    il, ia = C.c_int(), C.c_int()
    l.orc_get_intra_edges(edge.ctypes.data, reg.ctypes.data, img_full.strides[0] // img_full.itemsize,
                          img_full.itemsize, img_full.shape[1], img_full.shape[0], rx, ry, rw, rh, dec[0], dec[1],
                          int(it["part_x"]), int(it["part_y"]), int(it["bx"]), int(it["by"]), int(it["bsize"]),
                          int(it["po_x"]), int(it["po_y"]), TX[it["tx_size"]][0], TX[it["tx_size"]][1], bd,
                          -1 if it["mode"] == 255 else int(it["mode"]), int(it["enable_ief"]), int(it["angle_delta"]),
                          C.byref(il), C.byref(ia))
    return edge, il.value, ia.value


def make_items(rng, rect_w, rect_h, dec, n):
    """random partitions with their transform blocks; offsets are relative to the region"""
    items = np.zeros(n, B.EDGE_ITEM_DTYPE)
    k = 0
    while k < n:
        bs = int(rng.integers(0, 22))
        bw, bh = BSIZES[bs]
        if bw > 64 or bh > 64:
            continue
        # partition position in 4x4 luma units, aligned to its own size
        pw, ph = max(bw >> dec[0], 4), max(bh >> dec[1], 4)           # plane-domain size of the partition
        px = int(rng.integers(0, max(1, (rect_w + pw - 1) // pw))) * pw
        py = int(rng.integers(0, max(1, (rect_h + ph - 1) // ph))) * ph
        if px >= rect_w or py >= rect_h:
            continue
        # a transform size that tiles the plane-domain partition
        cands = [i for i, (tw, th) in enumerate(TX) if tw <= pw and th <= ph and pw % tw == 0 and ph % th == 0]
        ts = int(rng.choice(cands))
        tw, th = TX[ts]
        bx, by = int(rng.integers(0, pw // tw)), int(rng.integers(0, ph // th))
        x, y = px + bx * tw, py + by * th
        if x >= rect_w or y >= rect_h:
            continue
        mode = int(rng.choice(list(range(14)) + [255]))
        items[k] = (x, y, (px << dec[0]) >> 2, (py << dec[1]) >> 2, bx, by, bs, ts, mode,
                    int(rng.integers(-3, 4)) if 1 <= mode <= 8 else 0, int(rng.integers(0, 2)), 0)
        k += 1
    return items


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10)])
@pytest.mark.parametrize("dec", [(0, 0), (1, 1), (1, 0)])
def test_edges_match_oracle(dtype, bd, dec):
    c = G.ctx()
    rng = np.random.default_rng(bd * 7 + dec[0] * 2 + dec[1])
    PW, PH = 208 >> dec[0], 152 >> dec[1]                 # plane (not a multiple of 64: overhanging blocks)
    img = rng.integers(0, 1 << bd, (PH, PW)).astype(dtype)
    pl = c.plane_from_host(img, 8)
    for rect in ((0, 0, PW, PH), (64 >> dec[0], 64 >> dec[1], 128 >> dec[0], 64 >> dec[1])):   # frame, a tile
        items = make_items(rng, min(rect[2], PW - rect[0]), min(rect[3], PH - rect[1]), dec, 700)
        n = len(items)
        d_edges = G.dev_empty(n * 257 * img.itemsize)
        d_lens = G.dev_empty(2 * n)
        c.get_intra_edges_dev(pl, rect, dec[0], dec[1], bd, G.to_dev(items), n, d_edges, d_lens)
        c.synchronize()
        got = G.from_dev(d_edges, dtype)[:n * 257].reshape(n, 257)
        lens = G.from_dev(d_lens, np.uint8)[:2 * n].reshape(n, 2)
        for k in range(n):
            want, il, ia = oracle_edges(img, rect, dec, bd, items[k])
            assert (int(lens[k, 0]), int(lens[k, 1])) == (il, ia), items[k]
            np.testing.assert_array_equal(got[k, 128 - il:129 + ia], want[128 - il:129 + ia], err_msg=str(items[k]))
            assert (got[k, :128 - il] == 0).all() and (got[k, 129 + ia:] == 0).all()
    c.plane_free(pl)


def test_gathered_edges_drive_the_prediction_kernel():
    """edges built on the device -> b200_predict_intra_dev == oracle edges -> oracle prediction"""
    import torch
    c = G.ctx()
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (128, 128)).astype(np.uint8)
    pl = c.plane_from_host(img, 8)
    rows = []
    for y in range(0, 128, 16):
        for x in range(0, 128, 16):
            for mode in (0, 1, 2, 3, 4, 7, 9, 12):
                rows.append((x, y, x >> 2, y >> 2, 0, 0, BSIZES.index((16, 16)), 2, mode, 0, 1, 0))
    items = np.array(rows, B.EDGE_ITEM_DTYPE)
    n = len(items)
    d_edges, d_lens = G.dev_empty(n * 257), G.dev_empty(2 * n)
    c.get_intra_edges_dev(pl, (0, 0, 128, 128), 0, 0, 8, G.to_dev(items), n, d_edges, d_lens)
    c.synchronize()
    lens = G.from_dev(d_lens, np.uint8)[:2 * n].reshape(n, 2)
    angle = {1: 90, 2: 180, 3: 45, 4: 135, 7: 203}
    pit = np.zeros(n, B.INTRA_ITEM_DTYPE)
    for k, it in enumerate(items):
        x, y, mode = int(it["po_x"]), int(it["po_y"]), int(it["mode"])
        variant = (1 if x else 0) + (2 if y else 0)        # PredictionVariant::new, predict.rs:126-135
        variant = {0: 0, 1: 1, 2: 2, 3: 3}[variant]
        m = mode
        if mode == 12:                                     # predict_intra's PAETH remap, predict.rs:228-234
            m = {0: 0, 2: 1, 1: 2, 3: 12}[variant]
        pit[k] = (k, 0, x, y, angle.get(m, 0), m, variant, 0 if 1 <= m <= 8 else -1, lens[k, 0], lens[k, 1], 0)
    d_out = torch.empty((n, 16, 16), dtype=torch.uint8, device="cuda")
    c.predict_intra_dev(d_edges, G.to_dev(pit), n, None, 16, 16, 8, 128, 128, d_out)
    c.synchronize()
    got = d_out.cpu().numpy()
    for k, it in enumerate(items):
        want_e, il, ia = oracle_edges(img, (0, 0, 128, 128), (0, 0), 8, it)
        p = pit[k]
        want = O.predict_intra(int(p["mode"]), int(p["variant"]), want_e, 16, 16, 8, angle=int(p["angle"]),
                               ief=int(p["ief"]), left_len=il, above_len=ia, plane_w=128, plane_h=128,
                               dst_x=int(p["x"]), dst_y=int(p["y"]))
        np.testing.assert_array_equal(got[k], want, err_msg=str(it))
    c.plane_free(pl)
