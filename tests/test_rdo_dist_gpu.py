"""GPU parity for the RDO distortion kernels: get_weighted_sse / cdef_dist_kernel through the C
ABI == oracle/rdo_dist.c, bit exact (mirrors the asm==rust random tests of
asm/x86/dist/sse.rs:150-290 and asm/x86/dist/cdef_dist.rs:120-200)."""
import ctypes as C

import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests.test_oracle_rdo_dist import L as OL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10), (np.uint16, 12)])
@pytest.mark.parametrize("w,h", [(4, 4), (8, 8), (16, 16), (32, 16), (64, 64), (128, 128), (4, 16), (16, 4)])
def test_weighted_sse_blocks(dtype, bd, w, h):
    W, H, PAD = 256, 256, 0
    rng = np.random.default_rng(w * 1000 + h + bd)
    maxv = (1 << bd) - 1
    a = rng.integers(0, maxv + 1, (H, W)).astype(dtype)
    b = np.clip(a.astype(np.int64) + rng.integers(-40, 41, a.shape), 0, maxv).astype(dtype)
    scale = rng.integers(1, 1 << 20, (H // 4, W // 4)).astype(np.uint32)
    c = G.ctx()
    da, db = c.plane_from_host(a, PAD), c.plane_from_host(b, PAD)
    blocks = G.grid_blocks(W, H, w, h)
    blocks["x"] = (blocks["x"] + 4 * (np.arange(len(blocks)) % 3)) % (W - w + 1) // 4 * 4   # any 4-px position
    n = len(blocks)
    d_out = G.dev_empty(8 * n)
    c.weighted_sse_dev(da, db, G.to_dev(blocks), n, w, h, G.to_dev(scale), W // 4, d_out)
    c.synchronize()
    got = G.from_dev(d_out, np.uint64)[:n]
    ol = OL()
    fn = ol.orc_weighted_sse_u8 if dtype == np.uint8 else ol.orc_weighted_sse_u16
    isz = a.itemsize
    for i, blk in enumerate(blocks):
        x, y = int(blk["x"]), int(blk["y"])
        want = fn(a.ctypes.data + (y * W + x) * isz, W, b.ctypes.data + (y * W + x) * isz, W,
                  scale.ctypes.data + ((y // 4) * (W // 4) + x // 4) * 4, W // 4, w, h)
        assert int(got[i]) == want, (i, x, y)
    for pl in (da, db):
        c.plane_free(pl)


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10), (np.uint16, 12)])
@pytest.mark.parametrize("w,h", [(8, 8), (4, 4), (8, 4), (4, 8), (6, 8), (8, 3)])
def test_cdef_dist_blocks(dtype, bd, w, h):
    W, H = 192, 128
    rng = np.random.default_rng(w * 100 + h + bd)
    maxv = (1 << bd) - 1
    s = rng.integers(0, maxv + 1, (H, W)).astype(dtype)
    s[:64] = (s[:64].astype(np.int64) // 16 + maxv // 2).astype(dtype)            # a low-variance half
    d = np.clip(s.astype(np.int64) + rng.integers(-12, 13, s.shape), 0, maxv).astype(dtype)
    c = G.ctx()
    ds_, dd = c.plane_from_host(s, 0), c.plane_from_host(d, 0)
    blocks = G.grid_blocks(W, H, 8, 8)
    n = len(blocks)
    d_out, d_raw = G.dev_empty(4 * n), G.dev_empty(12 * n)
    c.cdef_dist_dev(ds_, dd, G.to_dev(blocks), n, w, h, bd, d_out, d_raw)
    c.synchronize()
    got, raw = G.from_dev(d_out, np.uint32)[:n], G.from_dev(d_raw, np.uint32)[:3 * n].reshape(n, 3)
    ol = OL()
    fn = ol.orc_cdef_dist_kernel_u8 if dtype == np.uint8 else ol.orc_cdef_dist_kernel_u16
    isz = s.itemsize
    r3 = (C.c_uint32 * 3)()
    for i, blk in enumerate(blocks):
        x, y = int(blk["x"]), int(blk["y"])
        want = fn(s.ctypes.data + (y * W + x) * isz, W, d.ctypes.data + (y * W + x) * isz, W, w, h, bd, r3)
        assert int(got[i]) == want, (i, x, y)
        assert tuple(int(v) for v in raw[i]) == (r3[0], r3[1], r3[2])
    for pl in (ds_, dd):
        c.plane_free(pl)


def test_percall_forms_match_oracle():
    """The symbols a Rust table entry would bind: host pointers, byte strides."""
    lib = B.lib()
    ol = OL()
    rng = np.random.default_rng(3)
    for dtype, bd in ((np.uint8, 8), (np.uint16, 10)):
        maxv = (1 << bd) - 1
        a = rng.integers(0, maxv + 1, (64, 80)).astype(dtype)
        b = rng.integers(0, maxv + 1, (64, 96)).astype(dtype)
        scale = rng.integers(1, 1 << 18, (16, 24)).astype(np.uint32)
        isz = a.itemsize
        for w, h in ((16, 16), (8, 32), (64, 64)):
            want = (ol.orc_weighted_sse_u8 if isz == 1 else ol.orc_weighted_sse_u16)(
                a.ctypes.data, 80, b.ctypes.data, 96, scale.ctypes.data, 24, w, h)
            got = lib.b200_weighted_sse(a.ctypes.data, 80 * isz, b.ctypes.data, 96 * isz, scale.ctypes.data,
                                        24 * 4, w, h, isz)
            assert got == want
        r3, g3 = (C.c_uint32 * 3)(), (C.c_uint32 * 3)()
        for w, h in ((8, 8), (4, 8), (8, 4), (4, 4)):
            want = (ol.orc_cdef_dist_kernel_u8 if isz == 1 else ol.orc_cdef_dist_kernel_u16)(
                a.ctypes.data, 80, b.ctypes.data, 96, w, h, bd, r3)
            got = lib.b200_cdef_dist_kernel(a.ctypes.data, 80 * isz, b.ctypes.data, 96 * isz, w, h, bd, g3)
            assert got == want and tuple(g3) == tuple(r3)
