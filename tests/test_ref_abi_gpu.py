"""The reference-signature symbols (rav1e_b200/csrc/ref_abi.cu): every family is called through the
reference's own argument list (asm/x86/mc.rs:17-76, asm/x86/predict.rs:21-234, asm/x86/cdef.rs:16-37,
:184-191) and compared with the oracle; plus re-entrancy: concurrent calls from several host threads
(one rayon worker per tile in the reference, encoder.rs:3253) each run on their own context."""
import ctypes as C
import threading

import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import gpu_util as G
from tests import oracle_lib as O

pytestmark = pytest.mark.gpu

FILTERS = [("8tap_regular", 0, 0), ("8tap_regular_smooth", 0, 1), ("8tap_regular_sharp", 0, 2),
           ("8tap_smooth_regular", 1, 0), ("8tap_smooth", 1, 1), ("8tap_smooth_sharp", 1, 2),
           ("8tap_sharp_regular", 2, 0), ("8tap_sharp_smooth", 2, 1), ("8tap_sharp", 2, 2), ("bilin", 3, 3)]
M = {n: i for i, n in enumerate(["DC_PRED", "V_PRED", "H_PRED", "D45_PRED", "D135_PRED", "D113_PRED", "D157_PRED",
                                 "D203_PRED", "D67_PRED", "SMOOTH_PRED", "SMOOTH_V_PRED", "SMOOTH_H_PRED",
                                 "PAETH_PRED", "UV_CFL_PRED"])}


def fn(name, restype=None):
    f = getattr(C.CDLL(B.LIB_PATH), name)
    f.restype = restype
    return f


vp, pd, i32 = C.c_void_p, C.c_ssize_t, C.c_int


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10), (np.uint16, 12)])
def test_mc_symbols(dtype, bd):
    rng = np.random.default_rng(bd)
    img = rng.integers(0, 1 << bd, (96, 128)).astype(dtype)
    op = O.Plane(128, 96, 16, dtype=dtype)
    op.fill_from(img)
    isz, bpc = img.itemsize, 8 if bd == 8 else 16
    hbd = [(1 << bd) - 1] if bd > 8 else []
    for name, mx, my in FILTERS:
        put, prep = fn(f"rav1e_put_{name}_{bpc}bpc_cuda"), fn(f"rav1e_prep_{name}_{bpc}bpc_cuda")
        put.argtypes = [vp, pd, vp, pd, i32, i32, i32, i32] + [i32] * len(hbd)
        prep.argtypes = [vp, vp, pd, i32, i32, i32, i32] + [i32] * len(hbd)
        for (w, h), (cf, rf) in (((8, 8), (0, 0)), ((16, 8), (5, 0)), ((4, 16), (0, 11)), ((32, 16), (7, 13))):
            got = np.zeros((h, w), dtype)
            put(got.ctypes.data, w * isz, op.at(24, 20), op.stride * isz, w, h, cf, rf, *hbd)
            np.testing.assert_array_equal(got, O.put_8tap(op, 24, 20, w, h, cf, rf, mx, my, bd), err_msg=name)
            gotp = np.zeros((h, w), np.int16)
            prep(gotp.ctypes.data, op.at(24, 20), op.stride * isz, w, h, cf, rf, *hbd)
            np.testing.assert_array_equal(gotp, O.prep_8tap(op, 24, 20, w, h, cf, rf, mx, my, bd), err_msg=name)
    t1 = O.prep_8tap(op, 20, 24, 16, 8, 3, 5, 0, 0, bd)
    t2 = O.prep_8tap(op, 33, 41, 16, 8, 12, 0, 1, 2, bd)
    avg = fn(f"rav1e_avg_{bpc}bpc_cuda")
    avg.argtypes = [vp, pd, vp, vp, i32, i32] + [i32] * len(hbd)
    got = np.zeros((8, 16), dtype)
    avg(got.ctypes.data, 16 * isz, t1.ctypes.data, t2.ctypes.data, 16, 8, *hbd)
    np.testing.assert_array_equal(got, O.mc_avg(t1, t2, bd))


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10)])
def test_ipred_symbols(dtype, bd):
    rng = np.random.default_rng(3 + bd)
    edge = rng.integers(0, 1 << bd, O.EDGE_LEN).astype(dtype)
    isz, bpc = edge.itemsize, 8 if bd == 8 else 16
    topleft = edge.ctypes.data + 128 * isz
    bdmax = (1 << bd) - 1
    plain = [("dc", "DC_PRED", 3), ("dc_128", "DC_PRED", 0), ("dc_left", "DC_PRED", 1), ("dc_top", "DC_PRED", 2),
             ("v", "V_PRED", 3), ("h", "H_PRED", 3), ("smooth", "SMOOTH_PRED", 3), ("smooth_v", "SMOOTH_V_PRED", 3),
             ("smooth_h", "SMOOTH_H_PRED", 3), ("paeth", "PAETH_PRED", 3)]
    for w, h in ((4, 4), (8, 8), (16, 8), (8, 32), (32, 32), (64, 64)):
        for name, mode, variant in plain:
            f = fn(f"rav1e_ipred_{name}_{bpc}bpc_cuda")
            f.argtypes = [vp, pd, vp, i32, i32, i32] + ([i32] * 3 if bd > 8 else [])
            angle = {"V_PRED": 90, "H_PRED": 180}.get(mode, 0)
            got = np.zeros((h, w), dtype)
            f(got.ctypes.data, w * isz, topleft, w, h, angle, *([0, 0, bdmax] if bd > 8 else []))
            want = O.predict_intra(M[mode], variant, edge, w, h, bd, angle=angle, left_len=h, above_len=w)
            np.testing.assert_array_equal(got, want, err_msg=f"{name} {w}x{h}")
        # directional zones: angle | enable_ief << 10 | smooth << 9
        for angle in (36, 45, 67, 87, 93, 113, 135, 157, 177, 183, 203, 225):
            for ief in (-1, 0, 1):
                zone = 1 if angle <= 90 else 2 if angle < 180 else 3
                arg = angle | ((ief >= 0) << 10) | ((ief > 0) << 9)
                f = fn(f"rav1e_ipred_z{zone}_{bpc}bpc_cuda")
                got = np.zeros((h, w), dtype)
                if zone == 2:
                    f.argtypes = [vp, pd, vp, i32, i32, i32, i32, i32] + ([i32] if bd > 8 else [])
                    f(got.ctypes.data, w * isz, topleft, w, h, arg, 200, 120, *([bdmax] if bd > 8 else []))
                    pw, ph = 200, 120
                else:
                    f.argtypes = [vp, pd, vp, i32, i32, i32] + ([i32] * 3 if bd > 8 else [])
                    f(got.ctypes.data, w * isz, topleft, w, h, arg, *([0, 0, bdmax] if bd > 8 else []))
                    pw = ph = 1 << 20
                ll = min(128, w + h) if zone == 3 else h
                al = min(128, w + h) if zone == 1 else w
                want = O.predict_intra(M["D45_PRED"], 3, edge, w, h, bd, angle=angle, ief=ief, left_len=ll,
                                       above_len=al, plane_w=pw, plane_h=ph, dst_x=0, dst_y=0)
                np.testing.assert_array_equal(got, want, err_msg=f"z{zone} {angle} ief{ief} {w}x{h}")
    # chroma from luma: ac through the cfl_ac symbols, prediction through the cfl symbols
    luma = rng.integers(0, 1 << bd, (64, 64)).astype(dtype)
    for lay, xdec, ydec in (("420", 1, 1), ("422", 1, 0), ("444", 0, 0)):
        for bw, bh, w_pad, h_pad in ((8, 8, 0, 0), (16, 16, 1, 0), (4, 4, 0, 0), (16, 8, 0, 1)):
            fac = fn(f"rav1e_ipred_cfl_ac_{lay}_{bpc}bpc_cuda")
            fac.argtypes = [vp, vp, pd, i32, i32, i32, i32]
            ac = np.zeros(bw * bh, np.int16)
            sub = np.ascontiguousarray(luma[:bh << ydec, :bw << xdec])
            fac(ac.ctypes.data, luma.ctypes.data, luma.strides[0], w_pad, h_pad, bw, bh)
            np.testing.assert_array_equal(ac, O.pred_cfl_ac(sub, bw, bh, w_pad, h_pad, xdec, ydec))
            for name, variant in (("cfl", 3), ("cfl_128", 0), ("cfl_left", 1), ("cfl_top", 2)):
                f = fn(f"rav1e_ipred_{name}_{bpc}bpc_cuda")
                f.argtypes = [vp, pd, vp, i32, i32, vp, i32] + ([i32] if bd > 8 else [])
                for alpha in (-9, 0, 5):
                    got = np.zeros((bh, bw), dtype)
                    f(got.ctypes.data, bw * isz, topleft, bw, bh, ac.ctypes.data, alpha, *([bdmax] if bd > 8 else []))
                    want = O.predict_intra(M["UV_CFL_PRED"], variant, edge, bw, bh, bd, angle=alpha, ac=ac,
                                           left_len=bh, above_len=bw)
                    np.testing.assert_array_equal(got, want, err_msg=f"{name} {lay} {bw}x{bh} alpha {alpha}")


@pytest.mark.parametrize("dtype,bd", [(np.uint8, 8), (np.uint16, 10)])
def test_cdef_symbols(dtype, bd):
    OL = O.lib()
    rng = np.random.default_rng(bd)
    isz = np.dtype(dtype).itemsize
    fdir = fn(f"rav1e_cdef_dir_{8 if bd == 8 else 16}bpc_cuda", i32)
    fdir.argtypes = [vp, pd, C.POINTER(C.c_uint32)] + ([i32] if bd > 8 else [])
    for t in range(10):
        img = (rng.integers(0, 256, (8, 24)) << (bd - 8)).astype(dtype)
        v = C.c_uint32()
        d = fdir(img.ctypes.data + 8 * isz, 24 * isz, C.byref(v), *([(1 << bd) - 1] if bd > 8 else []))
        assert (d, v.value) == O.cdef_find_dir(img[:, 8:16], bd)
    for size, xdec, ydec in (("4x4", 1, 1), ("4x8", 1, 0), ("8x8", 0, 0)):
        f = fn(f"rav1e_cdef_filter_{size}_cuda" if bd == 8 else f"rav1e_cdef_filter_{size}_16bpc_cuda")
        f.argtypes = [vp, pd, vp, pd, i32, i32, i32, i32] + ([i32] if bd > 8 else [])
        xs, ys = 8 >> xdec, 8 >> ydec
        for t in range(25):
            tmp = rng.integers(0, 1 << bd, (ys + 4, xs + 4)).astype(np.uint16)
            if t & 1:
                tmp[:, :2] = 0x8000
            if t & 2:
                tmp[-2:, :] = 0x8000
            pri, sec = int(rng.integers(0, 16)) << (bd - 8), int(rng.choice([0, 1, 2, 4])) << (bd - 8)
            d, damping = int(rng.integers(0, 8)), int(rng.integers(3, 7)) + bd - 8
            want = np.zeros((ys, xs), dtype)
            OL.orc_cdef_filter_block(O.ptr(want), xs, isz, O.ptr(tmp, 2 * (xs + 4) + 2), xs + 4, pri, sec, d,
                                     damping, bd, xdec, ydec, 15)
            got = np.zeros((ys, xs), dtype)
            f(got.ctypes.data, xs * isz, O.ptr(tmp, 2 * (xs + 4) + 2), (xs + 4) * 2, pri, sec, d, damping,
              *([(1 << bd) - 1] if bd > 8 else []))
            np.testing.assert_array_equal(got, want)


def test_percall_forms_are_reentrant():
    """8 host threads hammer the per-call symbols at once (ctypes releases the GIL): each thread gets
    its own context; every result must still be right."""
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (96, 128)).astype(np.uint8)
    op = O.Plane(128, 96, 16, dtype=np.uint8)
    op.fill_from(img)
    sad = fn("rav1e_sad16x16_cuda", C.c_uint32)
    sad.argtypes = [vp, pd, vp, pd]
    put = fn("rav1e_put_8tap_regular_8bpc_cuda")
    put.argtypes = [vp, pd, vp, pd, i32, i32, i32, i32]
    want_put = O.put_8tap(op, 24, 20, 16, 16, 5, 9, 0, 0, 8)
    errs = []

    def work(k):
        try:
            for it in range(40):
                x, y = 8 + (k * 7 + it) % 60, 4 + (k * 3 + it) % 40
                got = sad(op.at(16, 16), op.stride, op.at(x, y), op.stride)
                assert got == O.get_sad(op, 16, 16, op, x, y, 16, 16)
                g = np.zeros((16, 16), np.uint8)
                put(g.ctypes.data, 16, op.at(24, 20), op.stride, 16, 16, 5, 9)
                assert (g == want_put).all()
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs[:3]
