"""The committed golden fixtures (tests/golden/reference_kats.json, extracted from the reference's
own unit tests by tests/golden/make_golden.py) drive the oracle directly, and must agree with the
tables the other oracle tests carry inline."""
import ctypes as C
import json
import os

import numpy as np

from tests import oracle_lib as O
from tests.test_oracle_dist import SAD_KAT, SATD_KAT, setup_planes
from tests.test_oracle_predict import ANGLES, EXPECTED, kat_edge

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))
M = {n: i for i, n in enumerate(O.MODES)}


def test_inline_tables_equal_the_extracted_fixtures():
    assert [tuple(t) for t in G["sad"]] == SAD_KAT
    assert [tuple(t) for t in G["satd"]] == SATD_KAT
    assert G["directional_angles"] == ANGLES
    assert G["directional_4x4"] == EXPECTED


def test_oracle_against_golden_dist():
    L = O.lib()
    for dtype, sfx in ((np.uint8, "u8"), (np.uint16, "u16")):
        (inp, xo, yo, st), (rec, xr, yr, sr) = setup_planes(dtype)
        bx, by = G["dist_pattern"]["block_at"]
        for kind in ("sad", "satd"):
            f = getattr(L, f"orc_get_{kind}_{sfx}")
            for w, h, want in G[kind]:
                assert f(O.ptr(inp, (yo + by) * st + xo + bx), st, O.ptr(rec, (yr + by) * sr + xr + bx), sr, w, h) == want


def test_oracle_against_golden_intra():
    e = kat_edge()
    variants = {"DC_PRED": 3, "DC_TOP": 2, "DC_LEFT": 1, "DC_128": 0}
    for name, v in G["intra_4x4_const"].items():
        assert (O.predict_intra(M["DC_PRED"], variants[name], e, 4, 4, 8) == v).all(), name
    angle = {"V_PRED": 90, "H_PRED": 180}
    for name, want in G["intra_4x4"].items():
        got = O.predict_intra(M[name], 3, e, 4, 4, 8, angle=angle.get(name, 0))
        assert got.reshape(-1).tolist() == want, name
    for a, want in zip(G["directional_angles"], G["directional_4x4"]):
        assert O.predict_intra(M["D45_PRED"], 3, e, 4, 4, 8, angle=a).reshape(-1).tolist() == want, a


def test_oracle_against_golden_first_max_element():
    L = O.lib()
    for case in G["first_max_element"]:
        a = np.array(case["input"], np.int32)
        m = C.c_int32()
        i = L.orc_first_max_element(O.ptr(a), len(a), C.byref(m))
        assert [i, m.value] == case["expect"]
