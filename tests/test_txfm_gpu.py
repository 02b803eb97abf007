"""GPU parity: CUDA forward transform == oracle, bit exact, for every valid (TxSize, TxType)
pair the reference sweeps (transform/mod.rs:420-467; 160 pairs) at 8/10/12 bit — the CUDA
counterpart of the reference's asm==rust test (asm/shared/transform/forward.rs:54-110, input
range -255..255 at bd 8), extended to HBD ranges and both coefficient types."""
import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_all_valid_pairs_match_oracle(bd):
    c = G.ctx()
    rng = np.random.default_rng(100 + bd)
    lim = (1 << bd) - 1
    for tx_size, tx_type in O.valid_txfm_combos():
        w, h = O.TX_SIZES[tx_size]
        n = 37 if w * h <= 1024 else 5          # ragged vs the CTA's blocks-in-flight
        x = rng.integers(-lim, lim + 1, (n, h, w)).astype(np.int16)
        for i32 in ([False, True] if bd == 8 else [True]):
            want = O.forward_transform_batch(x, tx_size, tx_type, bd, coeff_i32=i32)
            got = c.fwd_txfm_batch(x, tx_size, tx_type, bd, coeff_i32=i32)
            assert got.dtype == want.dtype
            np.testing.assert_array_equal(got, want, err_msg=f"size {tx_size} type {tx_type} i32 {i32}")


def test_strided_input_and_device_resident():
    """Input rows wider than the block (a residual tile inside a 64x64 scratch, as
    encode_tx_block passes it) and device-resident buffers."""
    import torch
    c = G.ctx()
    rng = np.random.default_rng(3)
    n, w, h, stride = 50, 16, 16, 64
    scratch = rng.integers(-255, 256, (n, h, stride)).astype(np.int16)
    x = np.ascontiguousarray(scratch[:, :, :w])
    want = O.forward_transform_batch(x, 2, 1, 8, coeff_i32=False)
    d_in = torch.from_numpy(scratch).cuda()
    d_out = torch.empty((n, w * h), dtype=torch.int16, device="cuda")
    c.fwd_txfm_dev(d_in, h * stride, stride, d_out, n, 2, 1, 8, False)
    c.synchronize()
    np.testing.assert_array_equal(d_out.cpu().numpy(), want)


def test_invalid_pairs_are_rejected():
    """forward.rs:75 asserts valid_av1_transform: 64x64 ADST, 32x32 ADST, 8x8 WHT are errors."""
    c = G.ctx()
    x = np.zeros((1, 64, 64), np.int16)
    for tx_size, tx_type in ((4, 1), (3, 3), (1, 16), (9, 2)):
        w, h = O.TX_SIZES[tx_size]
        with pytest.raises(B.B200Error) as e:
            c.fwd_txfm_batch(np.zeros((1, h, w), np.int16), tx_size, tx_type)
        assert e.value.status == B.ERR_ARG
        assert not O.valid_txfm(tx_size, tx_type)


def test_percall_reference_signature():
    """b200_forward_transform(input, output, stride, tx_size, tx_type, bd, coeff_is_i32)."""
    L = B.lib()
    rng = np.random.default_rng(9)
    x = rng.integers(-255, 256, (1, 8, 8)).astype(np.int16)
    out = np.zeros(64, np.int16)
    L.b200_forward_transform(x.ctypes.data, out.ctypes.data, 8, 1, 3, 8, 0)
    np.testing.assert_array_equal(out, O.forward_transform_batch(x, 1, 3, 8, coeff_i32=False)[0])


def test_1080p_frame_sweep_checksum():
    """BASELINE config 3 shape: whole 1080p residual frame tiled by 16x16 DCT_DCT; compare a
    checksum of all coefficients plus a 1/97 sample against the oracle."""
    c = G.ctx()
    rng = np.random.default_rng(0)
    n = (1920 // 16) * (1080 // 16)
    x = rng.integers(-255, 256, (n, 16, 16)).astype(np.int16)
    got = c.fwd_txfm_batch(x, 2, 0, 8)
    want = O.forward_transform_batch(x, 2, 0, 8)
    assert int(got.astype(np.int64).sum()) == int(want.astype(np.int64).sum())
    np.testing.assert_array_equal(got[::97], want[::97])
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10)])
def test_fused_residual_transform(dtype, bd):
    """b200_fwd_txfm_residual_dev == oracle(diff -> forward_transform), and the standalone
    residual kernel agrees with numpy (encoder.rs:1533-1544)."""
    import torch
    c = G.ctx()
    W, H, PAD = 256, 128, 96
    cur, ref = G.make_planes(W, H, PAD, dtype, seed=5, bit_depth=bd)
    ocur, dcur = G.both_planes(cur, PAD)
    oref, dref = G.both_planes(ref, PAD)
    blocks = G.grid_blocks(W, H, 16, 16)
    n = len(blocks)
    rng = np.random.default_rng(1)
    best = np.zeros(n, B.ME_RESULT_DTYPE)
    best["mv_col"] = rng.integers(-20, 21, n) * 8 + rng.integers(-7, 8, n)
    best["mv_row"] = rng.integers(-20, 21, n) * 8 + rng.integers(-7, 8, n)
    best["cost"][::7] = np.uint64(2**64 - 1)          # empty results mean zero motion
    resid = np.zeros((n, 16, 16), np.int16)
    for i, b in enumerate(blocks):
        dx = int(np.trunc(best["mv_col"][i] / 8)) if best["cost"][i] != np.uint64(2**64 - 1) else 0
        dy = int(np.trunc(best["mv_row"][i] / 8)) if best["cost"][i] != np.uint64(2**64 - 1) else 0
        x, y = int(b["x"]), int(b["y"])
        p = PAD
        cb = ocur.data[p + y:p + y + 16, p + x:p + x + 16].astype(np.int32)
        rb = oref.data[p + y + dy:p + y + dy + 16, p + x + dx:p + x + dx + 16].astype(np.int32)
        resid[i] = cb - rb
    d_blocks, d_best = G.to_dev(blocks), G.to_dev(best)
    d_res = torch.empty((n, 16, 16), dtype=torch.int16, device="cuda")
    c.block_residual_dev(dcur, dref, d_blocks, n, d_best, 16, 16, d_res)
    i32 = bd > 8
    d_out = torch.empty((n, 256), dtype=torch.int32 if i32 else torch.int16, device="cuda")
    c.fwd_txfm_residual_dev(dcur, dref, d_blocks, n, d_best, d_out, 2, 1, bd)
    c.synchronize()
    np.testing.assert_array_equal(d_res.cpu().numpy(), resid)
    want = O.forward_transform_batch(resid, 2, 1, bd, coeff_i32=i32)
    np.testing.assert_array_equal(d_out.cpu().numpy(), want)


@pytest.mark.parametrize("dtype,bd,npairs", [(np.uint8, 8, 3), (np.uint16, 10, 2), (np.uint8, 8, 35)])
def test_fused_residual_transform_multi_pair(dtype, bd, npairs):
    """b200_fwd_txfm_residual_multi_dev over several plane pairs == one
    b200_fwd_txfm_residual_dev call per pair (itself pinned to the oracle above)."""
    import torch
    c = G.ctx()
    W, H, PAD = 128, 64, 64
    imgs = [G.make_planes(W, H, PAD, dtype, seed=70 + k, bit_depth=bd) for k in range(3)]
    dpl = [(c.plane_from_host(a, PAD), c.plane_from_host(b, PAD)) for a, b in imgs]
    curs = [dpl[k % 3][0] for k in range(npairs)]
    refs = [dpl[(k + 2) % 3][1] for k in range(npairs)]
    grid = G.grid_blocks(W, H, 16, 16)
    rng = np.random.default_rng(3)
    subs = [np.ascontiguousarray(grid[: int(rng.integers(0, len(grid) + 1))]) for _ in range(npairs)]
    blocks = np.concatenate(subs)
    n = len(blocks)
    best = np.zeros(n, B.ME_RESULT_DTYPE)
    best["mv_col"] = rng.integers(-20, 21, n) * 8
    best["mv_row"] = rng.integers(-20, 21, n) * 8
    ends = np.cumsum([len(x) for x in subs]).astype(np.uint32)
    i32 = bd > 8
    dt = torch.int32 if i32 else torch.int16
    d_blocks, d_best = G.to_dev(blocks), G.to_dev(best)
    d_out = torch.zeros((n, 256), dtype=dt, device="cuda")
    c.fwd_txfm_residual_multi_dev(B.PlanePairs(curs, refs, ends, ends), d_blocks, n, d_best, d_out, 2, 0, bd)
    c.synchronize()
    got = d_out.cpu().numpy()
    lo = 0
    for k, sub in enumerate(subs):
        if len(sub) == 0:
            continue
        d_one = torch.zeros((len(sub), 256), dtype=dt, device="cuda")
        c.fwd_txfm_residual_dev(curs[k], refs[k], G.to_dev(sub), len(sub), G.to_dev(best[lo:lo + len(sub)]),
                                d_one, 2, 0, bd)
        c.synchronize()
        np.testing.assert_array_equal(got[lo:lo + len(sub)], d_one.cpu().numpy())
        lo += len(sub)
    for a, b in dpl:
        c.plane_free(a)
        c.plane_free(b)


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10)])
def test_subpel_predict_diff_transform_chain(dtype, bd):
    """BASELINE config 4 chain on the device: put_8tap at sub-pel vectors -> residual against the
    packed prediction -> 16x16 DCT_DCT, equal to the oracle's mc -> diff -> forward_transform."""
    import torch
    c = G.ctx()
    W, H, PAD = 256, 128, 96
    cur, ref = G.make_planes(W, H, PAD, dtype, seed=11, bit_depth=bd)
    ocur, dcur = G.both_planes(cur, PAD)
    oref, dref = G.both_planes(ref, PAD)
    blocks = G.grid_blocks(W, H, 16, 16)
    n = len(blocks)
    mvs = np.random.default_rng(2).integers(-20 * 8, 20 * 8 + 1, (n, 2)).astype(np.int16)
    pred = O.mc_blocks(oref, blocks, mvs, 16, 16, 0, 0, bd)
    resid = np.zeros((n, 16, 16), np.int16)
    for i, b in enumerate(blocks):
        x, y = int(b["x"]) + PAD, int(b["y"]) + PAD
        resid[i] = ocur.data[y:y + 16, x:x + 16].astype(np.int32) - pred[i].astype(np.int32)
    want = O.forward_transform_batch(resid, 2, 0, bd, coeff_i32=bd > 8)
    d_blocks = G.to_dev(blocks)
    d_pred = torch.empty((n, 16, 16), dtype=torch.uint8 if bd == 8 else torch.int16, device="cuda")
    c.mc_blocks_dev(dref, d_blocks, G.to_dev(mvs), n, 16, 16, 0, 0, bd, 0, 0, 0, d_pred)
    d_out = torch.empty((n, 256), dtype=torch.int32 if bd > 8 else torch.int16, device="cuda")
    c.fwd_txfm_pred_dev(dcur, d_pred, d_blocks, n, d_out, 2, 0, bd)
    c.synchronize()
    np.testing.assert_array_equal(d_out.cpu().numpy(), want)
