"""GPU parity for the quantize -> dequantize -> tx-domain distortion chain: CUDA (through the C
ABI) == oracle/quantize.c, bit exact, for every transform size, the three scan classes, both
coefficient types and a spread of step sizes."""
import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests import oracle_lib as O
from tests.test_oracle_quantize import L as OL
from tests.test_oracle_quantize import chain as oracle_chain

pytestmark = pytest.mark.gpu


def synth_coeffs(ts, tt, n, dtype, seed, scale=2500.0):
    ol = OL()
    w, h = O.TX_SIZES[ts]
    area, coded = w * h, ol.orc_coded_tx_area(ts)
    scan = np.zeros(coded, np.uint16)
    ol.orc_scan_order(ts, tt, scan.ctypes.data, None)
    env = np.zeros(area)
    env[:coded][scan] = scale / (1.0 + np.arange(coded)) ** 0.9
    env[coded:] = 3.0
    rng = np.random.default_rng(seed)
    c = np.rint(rng.normal(0, 1, (n, area)) * env).astype(dtype)
    c[0] = 0
    c[1, 1:] = 0
    c[2] = np.rint(rng.normal(0, 40, area)).astype(dtype)      # flat spectrum: long tails of 0/1 levels
    return c


def run_gpu(c, ts, tt, dcq, acq, intra):
    import torch
    ctx = G.ctx()
    n = len(c)
    coded = B.lib().b200_coded_tx_area(ts)
    i32 = c.dtype == np.int32
    dt = torch.int32 if i32 else torch.int16
    d_c = torch.from_numpy(c).cuda()
    d_q = torch.full((n, coded), 55, dtype=dt, device="cuda")       # garbage: the kernel zero-fills
    d_r = torch.full((n, coded), 66, dtype=dt, device="cuda")
    d_e = torch.zeros(n, dtype=torch.int16, device="cuda")
    d_d = torch.zeros(n, dtype=torch.int64, device="cuda")
    ctx.quantize_dev(d_c, n, ts, tt, dcq, acq, intra, i32, d_q, d_r, d_e, d_d)
    ctx.synchronize()
    return (d_q.cpu().numpy(), d_r.cpu().numpy(), d_e.cpu().numpy().view(np.uint16),
            d_d.cpu().numpy().view(np.uint64))


@pytest.mark.parametrize("ts", range(19))
def test_every_size_default_scan(ts):
    for dtype, dcq, acq, intra in ((np.int16, 120, 96, False), (np.int32, 1336, 1828, True)):
        c = synth_coeffs(ts, 0, 40, dtype, seed=ts)
        want = oracle_chain(c, ts, 0, dcq, acq, intra)
        got = run_gpu(c, ts, 0, dcq, acq, intra)
        for g, w_, name in zip(got, want, ("qcoeffs", "rcoeffs", "eob", "tx_dist")):
            np.testing.assert_array_equal(g, w_, err_msg=f"{name} ts={ts} {dtype}")


@pytest.mark.parametrize("ts,tt", [(0, 10), (0, 11), (1, 12), (2, 13), (2, 9), (3, 9), (8, 14), (13, 15), (2, 4)])
@pytest.mark.parametrize("dcq,acq", [(4, 4), (33, 28), (300, 260), (5347, 7312)])
def test_scan_classes_and_step_sizes(ts, tt, dcq, acq):
    for dtype in (np.int16, np.int32):
        c = synth_coeffs(ts, tt, 33, dtype, seed=ts * 17 + tt, scale=6000.0)
        for intra in (True, False):
            want = oracle_chain(c, ts, tt, dcq, acq, intra)
            got = run_gpu(c, ts, tt, dcq, acq, intra)
            for g, w_, name in zip(got, want, ("qcoeffs", "rcoeffs", "eob", "tx_dist")):
                np.testing.assert_array_equal(g, w_, err_msg=f"{name} ts={ts} tt={tt} {dtype} intra={intra}")


def test_chained_after_the_fused_residual_transform():
    """encode_tx_block's order on the device: diff + forward_transform (b200_fwd_txfm_residual_dev)
    then the quantize chain, against the oracle's forward_transform + chain."""
    import torch
    c = G.ctx()
    W, H, PAD = 256, 128, 96
    cur, ref = G.make_planes(W, H, PAD, np.uint8, seed=8)
    ocur, dcur = G.both_planes(cur, PAD)
    oref, dref = G.both_planes(ref, PAD)
    blocks = G.grid_blocks(W, H, 16, 16)
    n = len(blocks)
    resid = np.zeros((n, 16, 16), np.int16)
    for i, b in enumerate(blocks):
        x, y = int(b["x"]) + PAD, int(b["y"]) + PAD
        resid[i] = ocur.data[y:y + 16, x:x + 16].astype(np.int32) - oref.data[y:y + 16, x:x + 16].astype(np.int32)
    coef = O.forward_transform_batch(resid, 2, 0, 8)
    want = oracle_chain(coef.reshape(n, 256), 2, 0, 60, 52, False)
    d_coef = torch.empty((n, 256), dtype=torch.int16, device="cuda")
    c.fwd_txfm_residual_dev(dcur, dref, G.to_dev(blocks), n, None, d_coef, 2, 0, 8)
    d_q = torch.empty((n, 256), dtype=torch.int16, device="cuda")
    d_r = torch.empty((n, 256), dtype=torch.int16, device="cuda")
    d_e = torch.zeros(n, dtype=torch.int16, device="cuda")
    d_d = torch.zeros(n, dtype=torch.int64, device="cuda")
    c.quantize_dev(d_coef, n, 2, 0, 60, 52, False, False, d_q, d_r, d_e, d_d)
    c.synchronize()
    np.testing.assert_array_equal(d_q.cpu().numpy(), want[0])
    np.testing.assert_array_equal(d_r.cpu().numpy(), want[1])
    np.testing.assert_array_equal(d_e.cpu().numpy().view(np.uint16), want[2])
    np.testing.assert_array_equal(d_d.cpu().numpy().view(np.uint64), want[3])
    assert (want[2] > 1).mean() > 0.5            # the blocks really carry coefficients
    for pl in (dcur, dref):
        c.plane_free(pl)
