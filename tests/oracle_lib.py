"""ctypes loader for oracle/liboracle.so — the CPU checker (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


class Mv(C.Structure):
    _fields_ = [("row", C.c_int16), ("col", C.c_int16)]


class MeResult(C.Structure):
    _fields_ = [("cost", C.c_uint64), ("sad", C.c_uint32), ("mv", Mv)]


ME_RESULT_DTYPE = np.dtype(
    {"names": ["cost", "sad", "mv_row", "mv_col"],
     "formats": ["<u8", "<u4", "<i2", "<i2"], "offsets": [0, 8, 12, 14], "itemsize": 16})
BLOCK_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2")])
CAND_DTYPE = np.dtype([("block", "<u4"), ("mv_row", "<i2"), ("mv_col", "<i2")])


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def host_threads():
    """CPU threads this process may actually use: the affinity mask capped by the cgroup CPU
    quota (the GPU boxes expose 128 logical CPUs under a 16-CPU CFS quota; running more runnable
    threads than the quota only gets the whole process throttled)."""
    n = len(os.sched_getaffinity(0))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, -(-int(quota) // int(period))))
    except (OSError, ValueError):
        pass
    return n


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR)
            if f.endswith((".c", ".h"))]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        build()
    # idle OpenMP workers must sleep, not spin: spinning burns the cgroup's CPU quota
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    L = C.CDLL(path)
    vp, pd, i32, u32, u64, sz = C.c_void_p, C.c_ssize_t, C.c_int, C.c_uint32, C.c_uint64, C.c_size_t
    for name in ("orc_get_sad_u8", "orc_get_sad_u16", "orc_get_satd_u8", "orc_get_satd_u16"):
        f = getattr(L, name)
        f.restype = u32
        f.argtypes = [vp, pd, vp, pd, i32, i32]
    L.orc_get_mv_rate.restype = u32
    L.orc_get_mv_rate.argtypes = [Mv, Mv, i32]
    L.orc_mv_cost.restype = u64
    L.orc_mv_cost.argtypes = [u32, Mv, Mv, Mv, u32, i32]
    L.orc_get_mv_range.restype = None
    L.orc_get_mv_range.argtypes = [i32] * 6 + [C.POINTER(i32)] * 4
    L.orc_full_search.restype = MeResult
    L.orc_full_search.argtypes = [vp, pd, vp, pd, i32, i32, i32, i32, i32, i32, i32, i32, i32,
                                  i32, u32, Mv, Mv, i32]
    L.orc_fullpel_candidates.restype = None
    L.orc_fullpel_candidates.argtypes = [vp, pd, vp, pd, i32, i32, i32, vp, vp, sz, i32, i32, i32,
                                         u32, vp, i32, vp, vp, i32]
    L.orc_subpel_candidates.restype = None
    L.orc_subpel_candidates.argtypes = [vp, pd, vp, pd, i32, i32, i32, vp, vp, sz, i32, i32, i32,
                                        u32, vp, i32, i32, i32, vp, vp, i32]
    L.orc_full_search_blocks.restype = None
    L.orc_full_search_blocks.argtypes = [vp, pd, vp, pd, i32, i32, i32, vp, sz, i32, i32, i32, i32,
                                         i32, u32, i32, vp, i32]
    L.orc_num_threads.restype = i32
    L.orc_put_8tap.restype = None
    L.orc_put_8tap.argtypes = [vp, pd, vp, pd] + [i32] * 8
    L.orc_prep_8tap.restype = None
    L.orc_prep_8tap.argtypes = [vp, vp, pd] + [i32] * 8
    L.orc_mc_avg.restype = None
    L.orc_mc_avg.argtypes = [vp, pd, i32, vp, vp, i32, i32, i32]
    L.orc_get_filter.restype = None
    L.orc_get_filter.argtypes = [i32, i32, i32, C.POINTER(C.c_int32)]
    L.orc_mc_blocks.restype = None
    L.orc_mc_blocks.argtypes = [vp, pd, i32, vp, vp, sz] + [i32] * 8 + [vp, i32]
    L.orc_first_max_element.restype = i32
    L.orc_first_max_element.argtypes = [vp, i32, C.POINTER(C.c_int32)]
    L.orc_cdef_find_dir.restype = i32
    L.orc_cdef_find_dir.argtypes = [vp, pd, i32, C.POINTER(u32), i32]
    L.orc_cdef_filter_block.restype = None
    L.orc_cdef_filter_block.argtypes = [vp, pd, i32, vp, pd] + [i32] * 8
    L.orc_cdef_filter_block_px.restype = None
    L.orc_cdef_filter_block_px.argtypes = [vp, pd, vp, pd] + [i32] * 9
    L.orc_cdef_adjust_strength.restype = i32
    L.orc_cdef_adjust_strength.argtypes = [i32, i32]
    L.orc_cdef_analyze_frame.restype = None
    L.orc_cdef_analyze_frame.argtypes = [vp, pd, i32, i32, i32, i32, vp, vp, vp]
    L.orc_cdef_filter_plane.restype = None
    L.orc_cdef_filter_plane.argtypes = [vp, pd, vp, pd] + [i32] * 8 + [vp, vp, vp, vp]
    L.orc_predict_intra.restype = None
    L.orc_predict_intra.argtypes = [i32, i32, vp, pd, i32, i32, i32, i32, vp, i32, i32, vp, i32, i32,
                                    i32, i32, i32, i32]
    L.orc_pred_cfl_ac.restype = None
    L.orc_pred_cfl_ac.argtypes = [vp, vp, pd] + [i32] * 7
    L.orc_valid_av1_transform.restype = i32
    L.orc_valid_av1_transform.argtypes = [i32, i32]
    L.orc_tx_width.restype = i32
    L.orc_tx_width.argtypes = [i32]
    L.orc_tx_height.restype = i32
    L.orc_tx_height.argtypes = [i32]
    L.orc_forward_transform.restype = None
    L.orc_forward_transform.argtypes = [vp, vp, sz, i32, i32, i32, i32]
    L.orc_forward_transform_batch.restype = None
    L.orc_forward_transform_batch.argtypes = [vp, vp, sz, i32, i32, i32, i32, i32]
    _LIB = L
    return L


def ptr(a, offset_elems=0):
    """Address of element `offset_elems` of a C-contiguous numpy array."""
    return a.ctypes.data + offset_elems * a.itemsize


class Plane:
    """A padded plane: `data` is the whole allocation, pixel (0,0) sits at (pad, pad)."""

    def __init__(self, width, height, pad, dtype=np.uint8, stride=None):
        self.width, self.height, self.pad = width, height, pad
        self.stride = stride or (width + 2 * pad)
        self.data = np.zeros((height + 2 * pad, self.stride), dtype=dtype)

    @property
    def bpp(self):
        return self.data.itemsize

    def origin_ptr(self):
        return ptr(self.data, self.pad * self.stride + self.pad)

    def at(self, x, y):
        return ptr(self.data, (self.pad + y) * self.stride + self.pad + x)

    def view(self):
        return self.data[self.pad:self.pad + self.height, self.pad:self.pad + self.width]

    def fill_from(self, img):
        """Copy `img` (height x width) in and replicate edges into the padding (Plane::pad)."""
        p = self.pad
        self.data[p:p + self.height, p:p + self.width] = img
        full = np.pad(img, ((p, p), (p, self.stride - self.width - p)), mode="edge")
        self.data[:, :] = full


def get_sad(org: Plane, ox, oy, ref: Plane, rx, ry, w, h):
    L = lib()
    f = L.orc_get_sad_u8 if org.bpp == 1 else L.orc_get_sad_u16
    return f(org.at(ox, oy), org.stride, ref.at(rx, ry), ref.stride, w, h)


def get_satd(org: Plane, ox, oy, ref: Plane, rx, ry, w, h):
    L = lib()
    f = L.orc_get_satd_u8 if org.bpp == 1 else L.orc_get_satd_u16
    return f(org.at(ox, oy), org.stride, ref.at(rx, ry), ref.stride, w, h)


def fullpel_candidates(cur: Plane, ref: Plane, blocks, cands, w, h, use_satd=False, lambda_=0,
                       pmv=None, allow_hp=False, want_cost=True, threads=0):
    L = lib()
    n = len(cands)
    sad = np.empty(n, np.uint32)
    cost = np.empty(n, np.uint64) if want_cost else None
    w_in_b = 2 * ((cur.width + 7) >> 3)
    h_in_b = 2 * ((cur.height + 7) >> 3)
    L.orc_fullpel_candidates(cur.origin_ptr(), cur.stride, ref.origin_ptr(), ref.stride, cur.bpp,
                             w_in_b, h_in_b, ptr(blocks), ptr(cands), n, w, h, int(use_satd),
                             int(lambda_), ptr(pmv) if pmv is not None else None, int(allow_hp),
                             ptr(sad), ptr(cost) if want_cost else None, threads)
    return sad, cost


def subpel_candidates(cur: Plane, ref: Plane, blocks, cands, w, h, use_satd=True, lambda_=0, pmv=None,
                      allow_hp=False, filter_mode=0, bit_depth=8, threads=0):
    L = lib()
    n = len(cands)
    sad = np.empty(n, np.uint32)
    cost = np.empty(n, np.uint64)
    w_in_b = 2 * ((cur.width + 7) >> 3)
    h_in_b = 2 * ((cur.height + 7) >> 3)
    L.orc_subpel_candidates(cur.origin_ptr(), cur.stride, ref.origin_ptr(), ref.stride, cur.bpp,
                            w_in_b, h_in_b, ptr(blocks), ptr(cands), n, w, h, int(use_satd),
                            int(lambda_), ptr(pmv) if pmv is not None else None, int(allow_hp),
                            filter_mode, bit_depth, ptr(sad), ptr(cost), threads)
    return sad, cost


def full_search_blocks(cur: Plane, ref: Plane, blocks, w, h, range_x, range_y, step, lambda_,
                       allow_hp=False, threads=0):
    L = lib()
    out = np.zeros(len(blocks), ME_RESULT_DTYPE)
    w_in_b = 2 * ((cur.width + 7) >> 3)
    h_in_b = 2 * ((cur.height + 7) >> 3)
    L.orc_full_search_blocks(cur.origin_ptr(), cur.stride, ref.origin_ptr(), ref.stride, cur.bpp,
                             w_in_b, h_in_b, ptr(blocks), len(blocks), w, h, range_x, range_y,
                             step, int(lambda_), int(allow_hp), ptr(out), threads)
    return out


def full_pixel_me_blocks(cur: Plane, ref: Plane, blocks, preds, subset_offsets, nsubsets, w, h, lambda_,
                         pmv=None, thresh=None, umh_range=0, allow_hp=False, threads=0):
    """full_pixel_me's search stages per block (oracle/me.c: orc_full_pixel_me_blocks)."""
    L = lib()
    out = np.zeros(len(blocks), ME_RESULT_DTYPE)
    w_in_b = 2 * ((cur.width + 7) >> 3)
    h_in_b = 2 * ((cur.height + 7) >> 3)
    preds = np.ascontiguousarray(preds)
    subset_offsets = np.ascontiguousarray(subset_offsets, np.uint32)
    pm = None if pmv is None else np.ascontiguousarray(pmv, np.int16)
    th = None if thresh is None else np.ascontiguousarray(thresh, np.uint32)
    L.orc_full_pixel_me_blocks.restype = None
    L.orc_full_pixel_me_blocks(
        C.c_void_p(cur.origin_ptr()), C.c_ssize_t(cur.stride), C.c_void_p(ref.origin_ptr()),
        C.c_ssize_t(ref.stride), C.c_int(cur.bpp), C.c_int(w_in_b), C.c_int(h_in_b),
        C.c_void_p(ptr(blocks)), C.c_size_t(len(blocks)), C.c_void_p(ptr(preds) if len(preds) else None),
        C.c_void_p(ptr(subset_offsets)), C.c_int(nsubsets), C.c_void_p(ptr(pm) if pm is not None else None),
        C.c_void_p(ptr(th) if th is not None else None), C.c_int(w), C.c_int(h), C.c_uint32(int(lambda_)),
        C.c_int(int(allow_hp)), C.c_int(umh_range), C.c_void_p(ptr(out)), C.c_int(threads))
    return out


# ---------------------------------------------------------------- forward transform
TX_SIZES = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (4, 8), (8, 4), (8, 16), (16, 8),
            (16, 32), (32, 16), (32, 64), (64, 32), (4, 16), (16, 4), (8, 32), (32, 8),
            (16, 64), (64, 16)]                      # (w, h) in TxSize order, transform/mod.rs:101
TX_TYPE_NAMES = ["DCT_DCT", "ADST_DCT", "DCT_ADST", "ADST_ADST", "FLIPADST_DCT", "DCT_FLIPADST",
                 "FLIPADST_FLIPADST", "ADST_FLIPADST", "FLIPADST_ADST", "IDTX", "V_DCT", "H_DCT",
                 "V_ADST", "H_ADST", "V_FLIPADST", "H_FLIPADST", "WHT_WHT"]


def valid_txfm(tx_size, tx_type):
    return bool(lib().orc_valid_av1_transform(tx_size, tx_type))


def valid_txfm_combos():
    """The reference's sweep set (transform/mod.rs:420-467): 160 (size, type) pairs."""
    return [(s, t) for s in range(19) for t in range(17) if valid_txfm(s, t)]


def forward_transform_batch(residual, tx_size, tx_type, bd=8, coeff_i32=None, threads=0):
    """residual: int16 array (n, h, w) contiguous -> coefficients (n, w*h) in the reference's
    transposed / 32x32-chunked order; int16 for 8-bit pixels, int32 for HBD (T::Coeff)."""
    L = lib()
    w, h = TX_SIZES[tx_size]
    residual = np.ascontiguousarray(residual, dtype=np.int16)
    n = residual.shape[0]
    assert residual.shape == (n, h, w)
    if coeff_i32 is None:
        coeff_i32 = bd > 8
    out = np.empty((n, w * h), np.int32 if coeff_i32 else np.int16)
    L.orc_forward_transform_batch(ptr(residual), ptr(out), n, tx_size, tx_type, bd, int(coeff_i32),
                                  threads)
    return out


# ---------------------------------------------------------------- motion compensation
def put_8tap(src: Plane, x, y, w, h, col_frac, row_frac, mode_x, mode_y, bit_depth):
    dst = np.zeros((h, w), src.data.dtype)
    lib().orc_put_8tap(ptr(dst), w, src.at(x, y), src.stride, src.bpp, w, h, col_frac, row_frac,
                       mode_x, mode_y, bit_depth)
    return dst


def prep_8tap(src: Plane, x, y, w, h, col_frac, row_frac, mode_x, mode_y, bit_depth):
    tmp = np.zeros((h, w), np.int16)
    lib().orc_prep_8tap(ptr(tmp), src.at(x, y), src.stride, src.bpp, w, h, col_frac, row_frac,
                        mode_x, mode_y, bit_depth)
    return tmp


def mc_avg(t1, t2, bit_depth):
    h, w = t1.shape
    dst = np.zeros((h, w), np.uint8 if bit_depth == 8 else np.uint16)
    lib().orc_mc_avg(ptr(dst), w, dst.itemsize, ptr(np.ascontiguousarray(t1)),
                     ptr(np.ascontiguousarray(t2)), w, h, bit_depth)
    return dst


def mc_blocks(ref: Plane, blocks, mvs, w, h, mode_x, mode_y, bit_depth, xdec=0, ydec=0, kind=0,
              threads=0):
    n = len(blocks)
    out = np.zeros((n, h, w), ref.data.dtype if kind == 0 else np.int16)
    mvs = np.ascontiguousarray(mvs, dtype=np.int16)
    lib().orc_mc_blocks(ref.origin_ptr(), ref.stride, ref.bpp, ptr(blocks), ptr(mvs), n, w, h,
                        mode_x, mode_y, bit_depth, xdec, ydec, kind, ptr(out), threads)
    return out


# ---------------------------------------------------------------- CDEF
def cdef_find_dir(img8x8, bit_depth=8):
    img = np.ascontiguousarray(img8x8)
    v = C.c_uint32()
    d = lib().orc_cdef_find_dir(ptr(img), img.shape[1], img.itemsize, C.byref(v), bit_depth - 8)
    return d, v.value


def cdef_analyze_frame(luma, bit_depth, skip8=None):
    luma = np.ascontiguousarray(luma)
    h, w = luma.shape
    dirs = np.zeros((h // 8, w // 8), np.uint8)
    var = np.zeros((h // 8, w // 8), np.int32)
    lib().orc_cdef_analyze_frame(ptr(luma), w, luma.itemsize, w, h, bit_depth,
                                 ptr(skip8) if skip8 is not None else None, ptr(dirs), ptr(var))
    return dirs, var


def cdef_filter_plane(img, plane, xdec, ydec, luma_w, luma_h, bit_depth, damping, skip8, dirs, var,
                      strength_sb):
    img = np.ascontiguousarray(img)
    out = np.zeros_like(img)
    lib().orc_cdef_filter_plane(ptr(img), img.shape[1], ptr(out), img.shape[1], img.itemsize, plane,
                                xdec, ydec, luma_w, luma_h, bit_depth, damping,
                                ptr(skip8) if skip8 is not None else None, ptr(dirs), ptr(var),
                                ptr(strength_sb))
    return out


# ---------------------------------------------------------------- intra prediction
MODES = ["DC_PRED", "V_PRED", "H_PRED", "D45_PRED", "D135_PRED", "D113_PRED", "D157_PRED",
         "D203_PRED", "D67_PRED", "SMOOTH_PRED", "SMOOTH_V_PRED", "SMOOTH_H_PRED", "PAETH_PRED",
         "UV_CFL_PRED"]                       # predict.rs:73-87
MODE_ANGLE = {"V_PRED": 90, "H_PRED": 180, "D45_PRED": 45, "D135_PRED": 135, "D113_PRED": 113,
              "D157_PRED": 157, "D203_PRED": 203, "D67_PRED": 67}   # intra_mode_to_angle
EDGE_LEN = 4 * 64 + 1


def predict_intra(mode, variant, edge, w, h, bit_depth, angle=0, ief=-1, ac=None, left_len=128,
                  above_len=128, plane_w=4096, plane_h=4096, dst_x=64, dst_y=64):
    """edge: the 257-element IntraEdge buffer (top-left at [128]).  Returns the h x w block."""
    edge = np.ascontiguousarray(edge)
    assert edge.shape == (EDGE_LEN,)
    dst = np.zeros((h, w), edge.dtype)
    if ac is not None:
        ac = np.ascontiguousarray(ac, dtype=np.int16)
    lib().orc_predict_intra(mode, variant, ptr(dst), w, edge.itemsize, w, h, bit_depth,
                            ptr(ac) if ac is not None else None, angle, ief, ptr(edge), left_len,
                            above_len, plane_w, plane_h, dst_x, dst_y)
    return dst


def pred_cfl_ac(luma, bw, bh, w_pad, h_pad, xdec, ydec):
    luma = np.ascontiguousarray(luma)
    ac = np.zeros(bw * bh, np.int16)
    lib().orc_pred_cfl_ac(ptr(ac), ptr(luma), luma.shape[1], luma.itemsize, bw, bh, w_pad, h_pad,
                          xdec, ydec)
    return ac
