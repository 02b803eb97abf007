"""CPU tests of the oracle's inverse transform (oracle/inv_txfm.c + generated networks) and, through
it, of the FORWARD transform: the reference's own round-trip test (transform/mod.rs:479-617,
`roundtrips`: forward_transform -> inverse_transform_add reproduces the source within a tolerance
of 0..2 per (TxSize, TxType)) is restated with the reference's 44 combinations and tolerances
(fixtures extracted by tests/golden/make_golden.py), many random trials instead of one.

That test pins the forward and inverse restatements against each other under the reference's own
acceptance bounds: a wrong constant, shift, flip or output permutation on either side breaks the
tolerance-0 combinations immediately."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests import oracle_lib as O

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
TX_NAMES = ["TX_4X4", "TX_8X8", "TX_16X16", "TX_32X32", "TX_64X64", "TX_4X8", "TX_8X4", "TX_8X16", "TX_16X8",
            "TX_16X32", "TX_32X16", "TX_32X64", "TX_64X32", "TX_4X16", "TX_16X4", "TX_8X32", "TX_32X8",
            "TX_16X64", "TX_64X16"]


def L():
    l = O.lib()
    l.orc_inverse_transform_add.restype = None
    l.orc_inverse_transform_add.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int,
                                            C.c_int, C.c_int]
    l.orc_inv_txfm_1d.restype = C.c_int
    l.orc_inv_txfm_1d.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    return l


def inverse_add(coef, dst, ts, tt, bd):
    l = L()
    dst = np.ascontiguousarray(dst)
    coef = np.ascontiguousarray(coef)
    l.orc_inverse_transform_add(coef.ctypes.data, int(coef.dtype == np.int32), dst.ctypes.data, dst.shape[1],
                                dst.itemsize, ts, tt, bd)
    return dst


def test_rect_ratio_kat():
    """transform/mod.rs:519-552 log_tx_ratios: the rectangular scaling decision of both directions."""
    for name, want in KATS["log_tx_ratios"]:
        w, h = O.TX_SIZES[TX_NAMES.index(name)]
        assert int(np.log2(w)) - int(np.log2(h)) == want


@pytest.mark.parametrize("pixel", [np.uint8, np.uint16])
def test_reference_roundtrips(pixel):
    """transform/mod.rs:479-617 with the reference's combinations and tolerances (roundtrips_u8 /
    roundtrips_u16: 8-bit values in either pixel type, bd = 8)."""
    rng = np.random.default_rng(0)
    for name, tname, tol in KATS["roundtrips"]:
        ts, tt = TX_NAMES.index(name), O.TX_TYPE_NAMES.index(tname)
        w, h = O.TX_SIZES[ts]
        worst = 0
        for _ in range(60):
            src = rng.integers(0, 256, (h, w)).astype(pixel)
            dst = rng.integers(0, 256, (h, w)).astype(pixel)
            res = (src.astype(np.int16) - dst.astype(np.int16)).reshape(1, h, w)
            freq = O.forward_transform_batch(res, ts, tt, 8, coeff_i32=(pixel == np.uint16))[0]
            rec = inverse_add(freq, dst, ts, tt, 8)
            worst = max(worst, int(np.abs(rec.astype(np.int32) - src.astype(np.int32)).max()))
        assert worst <= tol, (name, tname, worst, tol)


def test_roundtrip_of_every_other_valid_pair():
    """The reference lists 44 pairs; the remaining valid (size, type) pairs below 64 points must
    round-trip as well (tolerance 2, the loosest the reference uses; 64-point transforms drop the
    upper frequencies and are excluded upstream too)."""
    rng = np.random.default_rng(1)
    listed = {(a, b) for a, b, _ in KATS["roundtrips"]}
    for ts, tt in O.valid_txfm_combos():
        w, h = O.TX_SIZES[ts]
        if max(w, h) == 64 or (TX_NAMES[ts], O.TX_TYPE_NAMES[tt]) in listed:
            continue
        for _ in range(6):
            src = rng.integers(0, 256, (h, w)).astype(np.uint8)
            dst = rng.integers(0, 256, (h, w)).astype(np.uint8)
            res = (src.astype(np.int16) - dst.astype(np.int16)).reshape(1, h, w)
            freq = O.forward_transform_batch(res, ts, tt, 8)[0]
            rec = inverse_add(freq, dst, ts, tt, 8)
            assert int(np.abs(rec.astype(np.int32) - src.astype(np.int32)).max()) <= 2, (TX_NAMES[ts], tt)


def test_1d_networks_against_float_transforms():
    """Generated butterflies vs the textbook transforms they implement: the N-point AV1 inverse DCT
    is sqrt(N/2) times the orthonormal inverse DCT-II (one half_btf rounding per stage)."""
    l = L()
    rng = np.random.default_rng(2)
    for n in (4, 8, 16, 32, 64):
        k = np.arange(n)
        # orthonormal DCT-II matrix
        M = np.sqrt(2.0 / n) * np.cos(np.pi * (2 * k[None, :] + 1) * k[:, None] / (2 * n))
        M[0] /= np.sqrt(2.0)
        for _ in range(20):
            x = rng.integers(-2000, 2001, n).astype(np.int32)
            out = np.zeros(n, np.int32)
            assert l.orc_inv_txfm_1d(0, n, x.ctypes.data, out.ctypes.data, 20)
            want = (M.T @ x) * np.sqrt(n / 2.0)
            # 12-bit cosines: ~2^-12 relative error per stage on top of one rounding per stage
            assert np.abs(out - want).max() <= 2.0 + 1e-3 * np.abs(want).max(), (n, np.abs(out - want).max())
    # identity scalings: sqrt2, 2, 2 sqrt2, 4
    for n, g in ((4, np.sqrt(2.0)), (8, 2.0), (16, 2 * np.sqrt(2.0)), (32, 4.0)):
        x = rng.integers(-3000, 3001, n).astype(np.int32)
        out = np.zeros(n, np.int32)
        assert l.orc_inv_txfm_1d(3, n, x.ctypes.data, out.ctypes.data, 20)
        assert np.abs(out - x * g).max() <= 0.51 + 2e-4 * np.abs(x * g).max()   # 5793 / 4096 vs sqrt(2)
    # combinations the reference leaves unimplemented
    out = np.zeros(64, np.int32)
    for kind, n in ((1, 32), (1, 64), (2, 32), (3, 64), (4, 8)):
        assert l.orc_inv_txfm_1d(kind, n, out.ctypes.data, out.ctypes.data, 20) == 0


def test_flipadst_is_reversed_adst_and_wht_is_lossless():
    l = L()
    rng = np.random.default_rng(3)
    for n in (4, 8, 16):
        x = rng.integers(-1500, 1501, n).astype(np.int32)
        a, f = np.zeros(n, np.int32), np.zeros(n, np.int32)
        l.orc_inv_txfm_1d(1, n, x.ctypes.data, a.ctypes.data, 20)
        l.orc_inv_txfm_1d(2, n, x.ctypes.data, f.ctypes.data, 20)
        assert (f == a[::-1]).all()
    # WHT_WHT 4x4 is exactly invertible (tolerance 0 upstream) even on extreme residuals
    for _ in range(50):
        src = rng.integers(0, 256, (4, 4)).astype(np.uint8)
        dst = rng.choice([0, 255], (4, 4)).astype(np.uint8)
        res = (src.astype(np.int16) - dst.astype(np.int16)).reshape(1, 4, 4)
        freq = O.forward_transform_batch(res, 0, 16, 8)[0]
        assert (inverse_add(freq, dst, 0, 16, 8) == src).all()
