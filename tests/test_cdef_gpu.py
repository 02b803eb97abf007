"""GPU parity: CUDA CDEF == oracle, bit exact — counterpart of the reference's asm==rust tests
(asm/x86/cdef.rs:300-500): per-call cdef_dir / cdef_filter with the asm signatures, and the
frame-level batch (cdef_filter_tile semantics) for luma + 4:2:0 / 4:2:2 / 4:4:4 chroma, 8/10/12 bit."""
import ctypes as C

import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10), (np.uint16, 12)])
def test_percall_dir_and_filter(dtype, bd):
    L = B.lib()
    OL = O.lib()
    rng = np.random.default_rng(bd)
    isz = np.dtype(dtype).itemsize
    for t in range(20):
        img = (rng.integers(0, 256, (8, 24)) << (bd - 8)).astype(dtype)     # asm/x86/cdef.rs:376
        if t % 3 == 0:
            img[:] = np.sort(img, axis=t % 2)
        v = C.c_uint32()
        d = L.b200_cdef_dir(img.ctypes.data + 8 * isz, 24 * isz, C.byref(v), bd)
        wd, wv = O.cdef_find_dir(img[:, 8:16], bd)
        assert (d, v.value) == (wd, wv)
    # filter: padded u16 tile with sentinels, all strengths/dirs, three block shapes
    for xdec, ydec in ((0, 0), (1, 1), (1, 0)):
        xs, ys = 8 >> xdec, 8 >> ydec
        for t in range(40):
            tmp = rng.integers(0, 1 << bd, (ys + 4, xs + 4)).astype(np.uint16)
            edges = int(rng.integers(0, 16))
            if not edges & 1:
                tmp[:, :2] = 0x8000
            if not edges & 2:
                tmp[:, -2:] = 0x8000
            if not edges & 4:
                tmp[:2, :] = 0x8000
            if not edges & 8:
                tmp[-2:, :] = 0x8000
            pri = int(rng.integers(0, 16)) << (bd - 8)
            sec = int(rng.choice([0, 1, 2, 4])) << (bd - 8)
            d = int(rng.integers(0, 8))
            damping = int(rng.integers(3, 7)) + bd - 8
            want = np.zeros((ys, xs), dtype)
            OL.orc_cdef_filter_block(O.ptr(want), xs, isz, O.ptr(tmp, 2 * (xs + 4) + 2), xs + 4, pri, sec,
                                     d, damping, bd, xdec, ydec, 15)
            got = np.zeros((ys, xs), dtype)
            L.b200_cdef_filter_block(got.ctypes.data, xs * isz, O.ptr(tmp, 2 * (xs + 4) + 2),
                                     (xs + 4) * 2, pri, sec, d, damping, bd, xdec, ydec)
            np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10), (np.uint16, 12)])
@pytest.mark.parametrize("xdec,ydec", [(1, 1), (1, 0), (0, 0)])
def test_frame_batch_matches_oracle(dtype, bd, xdec, ydec):
    import torch
    c = G.ctx()
    rng = np.random.default_rng(bd * 10 + xdec * 2 + ydec)
    W, H = 200, 136                      # 3.1 x 2.1 superblocks: partial SBs at right/bottom
    base = rng.integers(0, 256, (H, W))
    k = np.ones(3) / 3
    sm = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, base.astype(np.float64))
    luma = (np.clip(sm + rng.normal(0, 6, sm.shape), 0, 255).astype(np.int64) << (bd - 8)).astype(dtype)
    chroma = (rng.integers(0, 256, (H >> ydec, W >> xdec)) << (bd - 8)).astype(dtype)
    skip8 = (rng.random((H // 8, W // 8)) < 0.2).astype(np.uint8)
    sbw, sbh = (W + 63) // 64, (H + 63) // 64
    y_str = rng.integers(0, 64, (sbh, sbw)).astype(np.uint8)
    uv_str = rng.integers(0, 64, (sbh, sbw)).astype(np.uint8)
    damping = 3 + int(rng.integers(0, 4))
    dirs, var = O.cdef_analyze_frame(luma, bd, skip8)
    want_y = O.cdef_filter_plane(luma, 0, 0, 0, W, H, bd, damping, skip8, dirs, var, y_str)
    want_c = O.cdef_filter_plane(chroma, 1, xdec, ydec, W, H, bd, damping, skip8, dirs, var, uv_str)

    d_skip = G.to_dev(skip8)
    d_dir = torch.empty(dirs.size, dtype=torch.uint8, device="cuda")
    d_var = torch.empty(var.size, dtype=torch.int32, device="cuda")
    pl_y = c.plane_from_host(luma, 0)
    c.cdef_find_dir_dev(pl_y, bd, d_skip, d_dir, d_var)
    c.synchronize()
    np.testing.assert_array_equal(d_dir.cpu().numpy().reshape(dirs.shape), dirs)
    np.testing.assert_array_equal(d_var.cpu().numpy().reshape(var.shape), var)
    for img, want, plane, xd, yd, strn in ((luma, want_y, 0, 0, 0, y_str), (chroma, want_c, 1, xdec, ydec, uv_str)):
        pin = c.plane_from_host(img, 0)
        pout = c.plane_from_host(np.zeros_like(img), 0)
        c.cdef_filter_plane_dev(pin, pout, plane, xd, yd, W, H, bd, damping, d_skip, d_dir, d_var,
                                G.to_dev(strn))
        got = np.zeros_like(img)
        c.check(c.L.b200_plane_download(c.h, C.byref(pout), got.ctypes.data, got.strides[0]))
        np.testing.assert_array_equal(got, want)
        c.plane_free(pin)
        c.plane_free(pout)
    c.plane_free(pl_y)


def test_4k_frame_properties():
    """BASELINE config 5 size (3840x2160 luma): flat frame is a fixed point; zero strength is
    identity; a 1/50 sample of 8x8 blocks equals the oracle run on just those blocks."""
    import torch
    c = G.ctx()
    W, H = 3840, 2160
    rng = np.random.default_rng(0)
    luma = rng.integers(0, 256, (H, W)).astype(np.uint8)
    pin = c.plane_from_host(luma, 0)
    pout = c.plane_from_host(np.zeros_like(luma), 0)
    n8 = (W // 8) * (H // 8)
    d_dir = torch.empty(n8, dtype=torch.uint8, device="cuda")
    d_var = torch.empty(n8, dtype=torch.int32, device="cuda")
    c.cdef_find_dir_dev(pin, 8, None, d_dir, d_var)
    sb = np.zeros(((H + 63) // 64, (W + 63) // 64), np.uint8)
    c.cdef_filter_plane_dev(pin, pout, 0, 0, 0, W, H, 8, 5, None, d_dir, d_var, G.to_dev(sb))
    got = np.zeros_like(luma)
    c.check(c.L.b200_plane_download(c.h, C.byref(pout), got.ctypes.data, got.strides[0]))
    np.testing.assert_array_equal(got, luma)
    sb[:] = 4 * 9 + 2
    c.cdef_filter_plane_dev(pin, pout, 0, 0, 0, W, H, 8, 5, None, d_dir, d_var, G.to_dev(sb))
    c.check(c.L.b200_plane_download(c.h, C.byref(pout), got.ctypes.data, got.strides[0]))
    dirs = d_dir.cpu().numpy().reshape(H // 8, W // 8)
    var = d_var.cpu().numpy().reshape(H // 8, W // 8)
    # oracle on a horizontal band of the frame (interior rows keep top/bottom availability)
    y0, y1 = 512, 640
    band = np.ascontiguousarray(luma[y0 - 8:y1 + 8])
    bd_dirs, bd_var = O.cdef_analyze_frame(band, 8)
    np.testing.assert_array_equal(bd_dirs[1:-1], dirs[y0 // 8:y1 // 8])
    np.testing.assert_array_equal(bd_var[1:-1], var[y0 // 8:y1 // 8])
    want = O.cdef_filter_plane(band, 0, 0, 0, W, band.shape[0], 8, 5, None, bd_dirs, bd_var,
                               np.full(((band.shape[0] + 63) // 64, (W + 63) // 64), 4 * 9 + 2, np.uint8))
    np.testing.assert_array_equal(got[y0:y1], want[8:-8])
    c.plane_free(pin)
    c.plane_free(pout)


def test_tile_rects_equal_the_whole_frame():
    """b200_cdef_find_dir_rect_dev / b200_cdef_filter_rect_dev over the tiles of a frame (each tile its
    own launch, like one rank per tile) == the whole-frame calls: taps cross tile borders, the
    sentinel appears only outside the frame."""
    import torch
    from rav1e_b200 import shard
    c = G.ctx()
    rng = np.random.default_rng(77)
    W, H, bd = 256, 192, 8
    luma = rng.integers(0, 256, (H, W)).astype(np.uint8)
    skip8 = (rng.random((H // 8, W // 8)) < 0.15).astype(np.uint8)
    sbw, sbh = (W + 63) // 64, (H + 63) // 64
    strength = rng.integers(0, 64, (sbh, sbw)).astype(np.uint8)
    src = c.plane_from_host(luma, 8)
    whole, tiled = c.plane_from_host(np.zeros_like(luma), 0), c.plane_from_host(np.zeros_like(luma), 0)
    n8 = (H // 8) * (W // 8)
    d_skip, d_str = G.to_dev(skip8), G.to_dev(strength)
    d_dir, d_var = torch.zeros(n8, dtype=torch.uint8, device="cuda"), torch.zeros(n8, dtype=torch.int32, device="cuda")
    d_dir2, d_var2 = torch.zeros_like(d_dir), torch.zeros_like(d_var)
    c.cdef_find_dir_dev(src, bd, d_skip, d_dir, d_var)
    c.cdef_filter_plane_dev(src, whole, 0, 0, 0, W, H, bd, 5, d_skip, d_dir, d_var, d_str)
    for (x, y, w, h) in shard.tile_grid(W, H, 1, 1):                     # 2 x 2 tiles of 128 x 128 / 128 x 64
        r8 = (x // 8, y // 8, w // 8, h // 8)
        c.cdef_find_dir_rect_dev(src, bd, d_skip, d_dir2, d_var2, r8)
        c.cdef_filter_rect_dev(src, tiled, 0, 0, 0, W, H, bd, 5, d_skip, d_dir2, d_var2, d_str, r8)
    c.synchronize()
    assert torch.equal(d_dir, d_dir2) and torch.equal(d_var, d_var2)
    # the same tiles (twice over: 8 items) through the many-items entry point, one call
    tiled2 = c.plane_from_host(np.zeros_like(luma), 0)
    d_dir3, d_var3 = torch.zeros_like(d_dir), torch.zeros_like(d_var)
    tl = shard.tile_grid(W, H, 1, 1)
    items = (B.CdefItem * (2 * len(tl)))()
    for k, (x, y, w, h) in enumerate(tl + tl):
        it = items[k]
        it.inp, it.out = C.pointer(src), C.pointer(tiled2)
        it.d_skip8, it.d_dir, it.d_var = d_skip.data_ptr(), d_dir3.data_ptr(), d_var3.data_ptr()
        it.rx8, it.ry8, it.rw8, it.rh8 = x // 8, y // 8, w // 8, h // 8
    c.cdef_tiles_dev(items, 0, 0, 0, W, H, bd, 5, d_str)
    c.synchronize()
    assert torch.equal(d_dir, d_dir3) and torch.equal(d_var, d_var3)
    t2 = np.zeros_like(luma)
    import ctypes
    c.check(c.L.b200_plane_download(c.h, ctypes.byref(tiled2), t2.ctypes.data, t2.strides[0]))
    c.plane_free(tiled2)
    a, b = np.zeros_like(luma), np.zeros_like(luma)
    c.check(c.L.b200_plane_download(c.h, ctypes.byref(whole), a.ctypes.data, a.strides[0]))
    c.check(c.L.b200_plane_download(c.h, ctypes.byref(tiled), b.ctypes.data, b.strides[0]))
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(a, t2)
    for p in (src, whole, tiled):
        c.plane_free(p)
