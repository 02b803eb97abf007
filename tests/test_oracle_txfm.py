"""Cross-checks that pin the oracle's forward transform as far as the reference allows.

rav1e stores no golden coefficients ("parity unpinned" for this row).  What can
be checked independently of the restatement:
  * every 1-D network equals the double-precision orthonormal DCT-II / DST-IV / DST-VII it
    is documented to approximate (forward_shared.rs comments), to a few LSB;
  * the 2-D driver (forward.rs:71-161): flips, the per-size shift triple, the transposed
    32x32-chunked output order and the i16 / i32 coefficient types, against a float model
      out = 2^(s0+s1+s2) * C_h X C_w^T   (SURVEY §8c scaling law);
  * the valid (size, type) set has the 160 members the reference sweeps
    (transform/mod.rs:420-467).
"""
import numpy as np
import pytest

from tests import oracle_lib as O

SHIFT_SUM = {  # sum of FWD_TXFM_SHIFT_LS rows (forward_shared.rs:22-40): same for 8/10/12 bit
    (4, 4): 3, (8, 8): 3, (16, 16): 3, (32, 32): 2, (64, 64): 1, (4, 8): 3, (8, 4): 3,
    (8, 16): 3, (16, 8): 3, (16, 32): 2, (32, 16): 2, (32, 64): 1, (64, 32): 1, (4, 16): 3,
    (16, 4): 3, (8, 32): 3, (32, 8): 3, (16, 64): 2, (64, 16): 2}


def mat_1d(kind, n):
    i = np.arange(n)[None, :].astype(np.float64)
    k = np.arange(n)[:, None].astype(np.float64)
    if kind == "DCT":
        m = np.cos(np.pi * (2 * i + 1) * k / (2 * n)) * np.sqrt(2.0 / n)
        m[0, :] = np.sqrt(1.0 / n)
        return m
    if kind in ("ADST", "FLIPADST"):
        if n == 4:      # DST-VII
            return np.sin(np.pi * (i + 1) * (2 * k + 1) / (2 * n + 1)) * (2.0 / np.sqrt(2 * n + 1))
        return np.sin(np.pi * (2 * i + 1) * (2 * k + 1) / (4 * n)) * np.sqrt(2.0 / n)   # DST-IV
    if kind == "IDTX":
        return np.eye(n)
    raise KeyError(kind)


def types_1d(tx_type):
    name = O.TX_TYPE_NAMES[tx_type]
    if name == "IDTX":
        return "IDTX", "IDTX"
    if name.startswith("V_"):
        return name[2:], "IDTX"
    if name.startswith("H_"):
        return "IDTX", name[2:]
    v, h = name.split("_")
    return v, h


def float_model(x, tx_size, tx_type):
    """x: (n, h, w) -> (n, h, w) float coefficients, row = vertical frequency."""
    w, h = O.TX_SIZES[tx_size]
    vt, ht = types_1d(tx_type)
    x = x.astype(np.float64)
    if vt == "FLIPADST":
        x = x[:, ::-1, :]
    if ht == "FLIPADST":
        x = x[:, :, ::-1]
    y = np.einsum("kr,nrc->nkc", mat_1d(vt, h), x)
    y = np.einsum("lc,nkc->nkl", mat_1d(ht, w), y)
    return y * (1 << SHIFT_SUM[(w, h)])


def unpack(out, tx_size):
    """Undo forward.rs:135-159: out[n, idx] -> coeff[n, r, c]."""
    w, h = O.TX_SIZES[tx_size]
    r = np.arange(h)[:, None]
    c = np.arange(w)[None, :]
    hs, wc = min(h, 32), min(w, 32)
    idx = (r >= 32) * hs * wc + h * 32 * (c >= 32) + (c % 32) * hs + (r % 32)
    return out[:, idx]


def test_valid_combo_count():
    combos = O.valid_txfm_combos()
    assert len(combos) == 160
    assert (0, 16) in combos and (1, 16) not in combos          # WHT only for 4x4
    assert (3, 9) in combos and (3, 1) not in combos            # 32x32: DCT_DCT + IDTX
    assert [t for s, t in combos if s == 4] == [0]              # 64x64: DCT_DCT only


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_against_float_orthonormal_model(bd):
    rng = np.random.default_rng(bd)
    lim = (1 << bd) - 1
    worst = 0.0
    for tx_size, tx_type in O.valid_txfm_combos():
        if tx_type == 16:
            continue
        w, h = O.TX_SIZES[tx_size]
        x = rng.integers(-lim, lim + 1, (6, h, w)).astype(np.int16)
        got = unpack(O.forward_transform_batch(x, tx_size, tx_type, bd, coeff_i32=True), tx_size)
        want = float_model(x, tx_size, tx_type)
        err = np.abs(got - want).max()
        worst = max(worst, err)
        # each lifting network is within ~4.5 LSB of orthonormal at its own scale; after the
        # second pass and the 2^k output scale the bound below has >2x margin
        assert err <= 12 * (1 << SHIFT_SUM[(w, h)]) / 2 + 8, (tx_size, tx_type, err)
    assert worst > 0          # it is an integer approximation, not the float model itself


def test_wht_constant_block():
    """fwht4 (forward_shared.rs:1778-1796) with shift [0,0,2] (:42): a constant block of 5 gives
    columns [10,0,0,0] (s0=10, s1=0, s2=5, q1=0, q0=10), rows [20,0,0,0]; shift[2] = +2 means
    av1_round_shift_array(.., bit = -2), i.e. << 2 (mod.rs:331-334): 80."""
    x = np.full((1, 4, 4), 5, np.int16)
    out = unpack(O.forward_transform_batch(x, 0, 16, 8, coeff_i32=True), 0)
    assert out[0, 0, 0] == 80
    assert np.count_nonzero(out) == 1


def test_i16_output_truncates_like_as_cast():
    """8-bit pixels use i16 coefficients (`T::cast_from`, forward.rs:157): same low 16 bits."""
    rng = np.random.default_rng(0)
    x = rng.integers(-255, 256, (4, 16, 16)).astype(np.int16)
    a = O.forward_transform_batch(x, 2, 0, 8, coeff_i32=False)
    b = O.forward_transform_batch(x, 2, 0, 8, coeff_i32=True)
    np.testing.assert_array_equal(a, b.astype(np.int16))


def test_output_order_64():
    """64-point sizes store four 32x32 chunks (forward.rs:135-159): a horizontal-frequency-0
    DC-only input lands at index 0, and the chunk bases follow the reference's order."""
    x = np.full((1, 64, 64), 3, np.int16)
    out = O.forward_transform_batch(x, 4, 0, 8, coeff_i32=True)[0]
    assert out[0] != 0 and np.count_nonzero(np.abs(out) > 2) == 1
