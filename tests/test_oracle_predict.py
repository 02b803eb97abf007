"""Pins the oracle's intra predictors against the reference's own known-answer tests
(/root/reference/src/predict.rs:1514-1619 `pred_matches_u8`, :1622-1693 `pred_max`)."""
import numpy as np
import pytest

from tests import oracle_lib as O

M = {n: i for i, n in enumerate(O.MODES)}
NONE, LEFT, TOP, BOTH = range(4)


def kat_edge(dtype=np.uint8):
    """predict.rs:1516-1517: edge_buf[i] = (i + 32).saturating_sub(MAX_TX_SIZE * 2)."""
    return np.maximum(np.arange(O.EDGE_LEN) + 32 - 128, 0).astype(dtype)


def test_pred_matches_u8_kat():
    e = kat_edge()
    p = lambda mode, variant, **kw: O.predict_intra(M[mode], variant, e, 4, 4, 8, **kw).reshape(-1).tolist()
    assert p("DC_PRED", BOTH) == [32] * 16
    assert p("DC_PRED", TOP) == [35] * 16
    # the KAT calls pred_dc_left with the 4-pixel left slice
    assert p("DC_PRED", LEFT) == [30] * 16
    assert p("DC_PRED", NONE) == [128] * 16
    assert p("V_PRED", BOTH, angle=90) == [33, 34, 35, 36] * 4
    assert p("H_PRED", BOTH, angle=180) == [31] * 4 + [30] * 4 + [29] * 4 + [28] * 4
    assert p("PAETH_PRED", BOTH) == [32, 34, 35, 36, 30, 32, 32, 36, 29, 32, 32, 32, 28, 28, 32, 32]
    assert p("SMOOTH_PRED", BOTH) == [32, 34, 35, 35, 30, 32, 33, 34, 29, 31, 32, 32, 29, 30, 32, 32]
    assert p("SMOOTH_H_PRED", BOTH) == [31, 33, 34, 35, 30, 33, 34, 35, 29, 32, 34, 34, 28, 31, 33, 34]
    assert p("SMOOTH_V_PRED", BOTH) == [33, 34, 35, 36, 31, 31, 32, 33, 30, 30, 30, 31, 29, 30, 30, 30]


ANGLES = [3, 6, 9, 14, 17, 20, 23, 26, 29, 32, 36, 39, 42, 45, 48, 51, 54, 58, 61, 64, 67, 70, 73,
          76, 81, 84, 87]
EXPECTED = [
    [40] * 16,
    [40] * 16,
    [39] + [40] * 15,
    [37, 38, 39] + [40] * 13,
    [36, 37, 38, 39] + [40] * 12,
    [36, 37, 38, 39, 39] + [40] * 11,
    [35, 36, 37, 38, 38, 39] + [40] * 10,
    [35, 36, 37, 38, 37, 38, 39, 40, 39] + [40] * 7,
    [35, 36, 37, 38, 37, 38, 39, 40, 38, 39] + [40] * 6,
    [35, 36, 37, 38, 36, 37, 38, 39, 38, 39, 40, 40, 39, 40, 40, 40],
    [34, 35, 36, 37, 36, 37, 38, 39, 37, 38, 39, 40, 39, 40, 40, 40],
    [34, 35, 36, 37, 36, 37, 38, 39, 37, 38, 39, 40, 38, 39, 40, 40],
    [34, 35, 36, 37, 35, 36, 37, 38, 36, 37, 38, 39, 37, 38, 39, 40],
    [34, 35, 36, 37, 35, 36, 37, 38, 36, 37, 38, 39, 37, 38, 39, 40],
    [34, 35, 36, 37, 35, 36, 37, 38, 36, 37, 38, 39, 37, 38, 39, 40],
    [34, 35, 36, 37, 35, 36, 37, 38, 35, 36, 37, 38, 36, 37, 38, 39],
    [34, 35, 36, 37, 34, 35, 36, 37, 35, 36, 37, 38, 36, 37, 38, 39],
    [34, 35, 36, 37, 34, 35, 36, 37, 35, 36, 37, 38, 36, 37, 38, 39],
    [34, 35, 36, 37, 34, 35, 36, 37, 35, 36, 37, 38, 35, 36, 37, 38],
    [33, 34, 35, 36, 34, 35, 36, 37, 34, 35, 36, 37, 35, 36, 37, 38],
    [33, 34, 35, 36, 34, 35, 36, 37, 34, 35, 36, 37, 35, 36, 37, 38],
    [33, 34, 35, 36, 34, 35, 36, 37, 34, 35, 36, 37, 34, 35, 36, 37],
    [33, 34, 35, 36, 34, 35, 36, 37, 34, 35, 36, 37, 34, 35, 36, 37],
    [33, 34, 35, 36, 33, 34, 35, 36, 34, 35, 36, 37, 34, 35, 36, 37],
    [33, 34, 35, 36, 33, 34, 35, 36, 34, 35, 36, 37, 34, 35, 36, 37],
    [33, 34, 35, 36, 33, 34, 35, 36, 33, 34, 35, 36, 33, 34, 35, 36],
    [33, 34, 35, 36, 33, 34, 35, 36, 33, 34, 35, 36, 33, 34, 35, 36],
]


def test_directional_kat_27_angles():
    """predict.rs:1568-1617: pred_directional on 4x4 with the 8-pixel left slice, no edge filter.
    (The mode only selects the directional path; any directional mode with this angle works.)"""
    e = kat_edge()
    for angle, want in zip(ANGLES, EXPECTED):
        got = O.predict_intra(M["D45_PRED"], BOTH, e, 4, 4, 8, angle=angle, ief=-1)
        assert got.reshape(-1).tolist() == want, angle


def test_pred_max_12bit_kat():
    """predict.rs:1622-1693: a full-scale 12-bit edge stays full scale through every predictor."""
    e = np.full(O.EDGE_LEN, 4095, np.uint16)
    for mode, angle in (("DC_PRED", 0), ("H_PRED", 180), ("V_PRED", 90), ("PAETH_PRED", 0),
                        ("SMOOTH_PRED", 0), ("SMOOTH_H_PRED", 0), ("SMOOTH_V_PRED", 0)):
        got = O.predict_intra(M[mode], BOTH, e, 4, 4, 12, angle=angle, left_len=4, above_len=4)
        assert (got == 4095).all(), mode


@pytest.mark.parametrize("w,h", [(4, 4), (8, 8), (16, 8), (8, 32), (64, 64)])
def test_directional_zones_are_consistent_with_v_and_h(w, h):
    """A constant edge predicts a constant block in every zone, with and without the edge
    filter / upsampling (they are convex combinations of edge pixels)."""
    e = np.full(O.EDGE_LEN, 77, np.uint8)
    for mode, base in O.MODE_ANGLE.items():
        for delta in (-9, -3, 0, 3, 9):
            for ief in (-1, 0, 1):
                got = O.predict_intra(M[mode], BOTH, e, w, h, 8, angle=base + delta, ief=ief,
                                      left_len=w + h, above_len=w + h)
                assert (got == 77).all(), (mode, delta, ief)


def test_cfl_ac_and_pred():
    rng = np.random.default_rng(0)
    luma = rng.integers(0, 256, (16, 16)).astype(np.uint8)
    ac = O.pred_cfl_ac(luma, 8, 8, 0, 0, 1, 1)
    # 4:2:0: sum of 2x2 luma << 1, minus the rounded average (predict.rs:1040-1062)
    s = (luma[0::2, 0::2].astype(np.int32) + luma[0::2, 1::2] + luma[1::2, 0::2] + luma[1::2, 1::2]) << 1
    avg = (int(s.sum()) + 32) >> 6
    np.testing.assert_array_equal(ac.reshape(8, 8), s - avg)
    e = np.maximum(np.arange(O.EDGE_LEN) + 32 - 128, 0).astype(np.uint8)
    dc = O.predict_intra(M["DC_PRED"], BOTH, e, 8, 8, 8)
    got = O.predict_intra(M["UV_CFL_PRED"], BOTH, e, 8, 8, 8, angle=5, ac=ac)
    q6 = 5 * ac.astype(np.int32)
    want = np.clip(int(dc[0, 0]) + np.sign(q6) * ((np.abs(q6) + 32) >> 6), 0, 255).reshape(8, 8)
    np.testing.assert_array_equal(got, want)
    # alpha == 0 leaves the DC prediction untouched (predict.rs:1069-1071)
    np.testing.assert_array_equal(O.predict_intra(M["UV_CFL_PRED"], BOTH, e, 8, 8, 8, angle=0, ac=ac), dc)
