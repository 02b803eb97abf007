"""GPU parity: CUDA intra prediction == oracle, bit exact — counterpart of the reference's
asm==rust test (asm/shared/predict.rs:31-169: 20 mode/variant pairs x angles x 3 ief settings x
bd 8/10/12 on 4x4), extended to all 19 TxSize shapes.  The oracle itself is pinned by the
reference KATs (tests/test_oracle_predict.py)."""
import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests import oracle_lib as O
from tests.test_oracle_predict import ANGLES, EXPECTED, kat_edge

pytestmark = pytest.mark.gpu
M = {n: i for i, n in enumerate(O.MODES)}


def run_items(edges, items, w, h, bd, ac=None, plane_w=4096, plane_h=4096):
    import torch
    c = G.ctx()
    n = len(items)
    d_out = torch.empty((n, h, w), dtype=torch.uint8 if bd == 8 else torch.int16, device="cuda")
    c.predict_intra_dev(G.to_dev(edges), G.to_dev(items), n, None if ac is None else G.to_dev(ac), w, h,
                        bd, plane_w, plane_h, d_out)
    c.synchronize()
    return d_out.cpu().numpy().view(edges.dtype)


def test_reference_kats_through_cuda():
    """predict.rs:1514-1619 evaluated by the CUDA kernel (batched) and by the per-call form."""
    e = kat_edge().reshape(1, -1)
    items = np.zeros(len(ANGLES), B.INTRA_ITEM_DTYPE)
    items["mode"], items["variant"], items["ief"] = M["D45_PRED"], 3, -1
    items["angle"] = ANGLES
    items["left_len"] = items["above_len"] = 128
    got = run_items(e, items, 4, 4, 8)
    for k, want in enumerate(EXPECTED):
        assert got[k].reshape(-1).tolist() == want, ANGLES[k]
    L = B.lib()
    dst = np.zeros((4, 4), np.uint8)
    for mode, variant, angle, want in (("DC_PRED", 3, 0, [32] * 16), ("DC_PRED", 2, 0, [35] * 16),
                                       ("DC_PRED", 1, 0, [30] * 16), ("DC_PRED", 0, 0, [128] * 16),
                                       ("PAETH_PRED", 3, 0, [32, 34, 35, 36, 30, 32, 32, 36, 29, 32, 32, 32, 28, 28, 32, 32]),
                                       ("SMOOTH_PRED", 3, 0, [32, 34, 35, 35, 30, 32, 33, 34, 29, 31, 32, 32, 29, 30, 32, 32])):
        L.b200_predict_intra(M[mode], variant, dst.ctypes.data, 4, 4, 4, 8, None, angle, -1,
                             e.ctypes.data, 128, 128, 4096, 4096, 64, 64)
        assert dst.reshape(-1).tolist() == want, mode


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10), (np.uint16, 12)])
def test_all_modes_sizes_match_oracle(dtype, bd):
    rng = np.random.default_rng(bd)
    nedge = 6
    edges = rng.integers(0, 1 << bd, (nedge, O.EDGE_LEN)).astype(dtype)
    edges[1] = np.sort(edges[1])                       # a smooth ramp
    edges[2] = (1 << bd) - 1                           # saturation
    for w, h in O.TX_SIZES:
        full = min(128, w + h)
        rows = []
        for ei in range(nedge):
            for mode in ("DC_PRED",):
                for variant in range(4):
                    rows.append((ei, M[mode], variant, 0, -1, h, w, 64, 64))
            for mode in ("SMOOTH_PRED", "SMOOTH_V_PRED", "SMOOTH_H_PRED", "PAETH_PRED"):
                rows.append((ei, M[mode], 3, 0, -1, h, w, 64, 64))
            for mode, base in O.MODE_ANGLE.items():
                for delta in (-3, -1, 0, 2, 3):
                    for ief in (-1, 0, 1):
                        # dst position: interior and near the plane's right/bottom edge (num_px clipping)
                        for x, y in ((64, 64), (4096 - w // 2, 4096 - h // 2)):
                            rows.append((ei, M[mode], 3, base + 3 * delta, ief, full, full, x, y))
        items = np.zeros(len(rows), B.INTRA_ITEM_DTYPE)
        for k, (ei, mode, variant, angle, ief, ll, al, x, y) in enumerate(rows):
            items[k] = (ei, 0, x, y, angle, mode, variant, ief, ll, al, 0)
        got = run_items(edges, items, w, h, bd)
        for k, (ei, mode, variant, angle, ief, ll, al, x, y) in enumerate(rows):
            want = O.predict_intra(mode, variant, edges[ei], w, h, bd, angle=angle, ief=ief, left_len=ll,
                                   above_len=al, plane_w=4096, plane_h=4096, dst_x=x, dst_y=y)
            np.testing.assert_array_equal(got[k], want, err_msg=f"{w}x{h} {O.MODES[mode]} var{variant} "
                                          f"angle{angle} ief{ief} at ({x},{y})")


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10)])
@pytest.mark.parametrize("xdec,ydec", [(1, 1), (1, 0), (0, 0)])
def test_cfl_ac_and_pred_match_oracle(dtype, bd, xdec, ydec):
    import torch
    c = G.ctx()
    rng = np.random.default_rng(7)
    W, H = 128, 96
    luma = rng.integers(0, 1 << bd, (H, W)).astype(dtype)
    pl = c.plane_from_host(luma, 0)
    for bw, bh, w_pad, h_pad in ((8, 8, 0, 0), (16, 16, 1, 0), (4, 4, 0, 0), (32, 32, 2, 3), (16, 8, 0, 1)):
        lw, lh = bw << xdec, bh << ydec
        blocks = G.grid_blocks(W, H, lw, lh)[:12]
        n = len(blocks)
        d_ac = torch.empty((n, bw * bh), dtype=torch.int16, device="cuda")
        c.pred_cfl_ac_dev(pl, G.to_dev(blocks), n, bw, bh, w_pad, h_pad, xdec, ydec, d_ac)
        c.synchronize()
        got = d_ac.cpu().numpy()
        for i, b in enumerate(blocks):
            sub = np.ascontiguousarray(luma[b["y"]:b["y"] + lh, b["x"]:b["x"] + lw])
            want = O.pred_cfl_ac(sub, bw, bh, w_pad, h_pad, xdec, ydec)
            np.testing.assert_array_equal(got[i], want)
        # CfL prediction: all four DC variants, alphas of both signs
        edges = rng.integers(0, 1 << bd, (2, O.EDGE_LEN)).astype(dtype)
        rows = [(i % 2, i, variant, alpha) for i in range(n) for variant in range(4) for alpha in (-13, 0, 7)]
        items = np.zeros(len(rows), B.INTRA_ITEM_DTYPE)
        for k, (ei, ai, variant, alpha) in enumerate(rows):
            items[k] = (ei, ai, 64, 64, alpha, M["UV_CFL_PRED"], variant, -1, bh, bw, 0)
        out = run_items(edges, items, bw, bh, bd, ac=got)
        for k, (ei, ai, variant, alpha) in enumerate(rows):
            want = O.predict_intra(M["UV_CFL_PRED"], variant, edges[ei], bw, bh, bd, angle=alpha, ac=got[ai],
                                   left_len=bh, above_len=bw)
            np.testing.assert_array_equal(out[k], want)
    c.plane_free(pl)
