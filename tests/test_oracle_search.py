"""CPU tests of the oracle's search stages (oracle/me.c: orc_full_pixel_me_blocks) against an
independent pure-Python restatement of me.rs:692-1303 on small cases, plus invariants.

No stored vector in the reference pins these stages (they need rustc to run): the Python model
below was written from the reference text separately from the C code, so a slip in either shows
up as a mismatch.
"""
import numpy as np

from tests import oracle_lib as O
from tests.oracle_lib import ME_RESULT_DTYPE

EMPTY = (2**64 - 1, 2**32 - 1)


def i16(v):
    return ((int(v) + 32768) & 0xFFFF) - 32768


def tdiv8(v):                      # Rust `/ 8` on i16: truncation toward zero
    return int(abs(v) // 8) * (1 if v >= 0 else -1)


def ilog(v):
    return 0 if v <= 0 else int(v).bit_length()


def mv_rate(a, b, hp):
    def d2r(d):
        d = i16(d)
        d = d if hp else d >> 1
        return 2 * ilog(abs(d))
    return d2r(a[0] - b[0]) + d2r(a[1] - b[1])


class Model:
    """One block's search context; mv = (row, col) in 1/8 pel."""

    def __init__(self, cur, ref, pad, bx, by, w, h, rng, lam, pmv, hp=False):
        self.cur, self.ref, self.pad = cur, ref, pad
        self.bx, self.by, self.w, self.h = bx, by, w, h
        self.rng, self.lam, self.pmv, self.hp = rng, lam, pmv, hp
        self.evals = 0

    def rd(self, mv):
        row, col = mv
        x0, x1, y0, y1 = self.rng
        if col < x0 or col > x1 or row < y0 or row > y1:
            return (EMPTY[0], EMPTY[1], mv)
        p = self.pad
        o = self.cur[p + self.by:p + self.by + self.h, p + self.bx:p + self.bx + self.w].astype(np.int64)
        yy, xx = p + self.by + tdiv8(row), p + self.bx + tdiv8(col)
        r = self.ref[yy:yy + self.h, xx:xx + self.w].astype(np.int64)
        sad = int(np.abs(o - r).sum())
        r1 = mv_rate(mv, self.pmv[0], self.hp)
        r2 = mv_rate(mv, self.pmv[1], self.hp) + 1
        self.evals += 1
        return (256 * sad + min(r1, r2) * self.lam, sad, mv)

    def best_of(self, mvs):
        best = (EMPTY[0], EMPTY[1], (0, 0))
        idx = 0
        for i, mv in enumerate(mvs):
            r = self.rd(mv)
            if r[0] < best[0]:
                best, idx = r, i
        return best, idx


def add(a, b):
    return (i16(a[0] + b[0]), i16(a[1] + b[1]))


def mul(a, k):
    return (i16(a[0] * k), i16(a[1] * k))


def fp(cols, rows):
    return [(r * 8, c * 8) for c, r in zip(cols, rows)]


DIAMOND = fp([0, 1, 0, -1], [1, 0, -1, 0])
HEXAGON = fp([0, 2, 2, 0, -2, -2], [-2, -1, 1, 2, 1, -1])
SQUARE = fp([-1, 0, 1, -1, 1, -1, 0, 1], [1, 1, 1, 0, 0, -1, -1, -1])
UMH = fp([-2, -1, 0, 1, 2, 3, 4, 3, 2, 1, 0, -1, -2, 3, -4, -3], [4, 4, 4, 4, 4, 2, 0, -2, -4, -4, -4, -4, -4, -2, 0, 2])


def diamond(m, cur):
    radius = 1
    while True:
        best, _ = m.best_of([add(cur[2], mul(o, 1 << radius)) for o in DIAMOND])
        if cur[0] <= best[0]:
            if radius == 0:
                return cur
            radius -= 1
        else:
            cur = best


def hexagon(m, cur):
    best, bi = m.best_of([add(cur[2], o) for o in HEXAGON])
    while best[0] < cur[0]:
        cur = best
        center = bi
        idxs = [(center + k) % 6 for k in (5, 6, 7)]
        best, j = m.best_of([add(cur[2], HEXAGON[i]) for i in idxs])
        if best[0] != EMPTY[0]:
            bi = idxs[j]
    best, _ = m.best_of([add(cur[2], o) for o in SQUARE])
    return best if best[0] < cur[0] else cur


def umh(m, cur, me_range):
    center = cur[2]
    for i in range(1, me_range + 1, 2):
        for o in fp([0, 0], [-1, 1]):
            r = m.rd(add(center, mul(o, i)))
            if r[0] < cur[0]:
                cur = r
    for i in range(1, (me_range >> 1) + 1, 2):
        for o in fp([-1, 1], [0, 0]):
            r = m.rd(add(center, mul(o, i)))
            if r[0] < cur[0]:
                cur = r
    center = cur[2]
    for row in range(-2, 3):
        for col in range(-2, 3):
            if row == 0 and col == 0:
                continue
            r = m.rd(add(center, (row, col)))
            if r[0] < cur[0]:
                cur = r
    center = cur[2]
    for i in range(1, (me_range >> 2) + 1):
        for o in UMH:
            r = m.rd(add(center, mul(o, i)))
            if r[0] < cur[0]:
                cur = r
    return hexagon(m, cur)


def full_pixel_me(m, subsets, extensive, thresh, umh_range):
    best = (EMPTY[0], EMPTY[1], (0, 0))

    def try_cands(preds, best):
        r, _ = m.best_of(preds)
        r = diamond(m, r)
        return r if r[0] < best[0] else best
    if not extensive:
        return try_cands(subsets[0], best)
    if len(subsets[0]):
        best = try_cands(subsets[0], best)
        if best[1] < thresh:
            return best
    for s in subsets[1:]:
        best = try_cands(s, best)
        if best[1] < thresh:
            return best
    if umh_range and best[0] != EMPTY[0]:
        best = umh(m, best, umh_range)
    return best


def mv_range(w_in_b, h_in_b, bx, by, w, h):
    bw, bh = 128 + w * 8, 128 + h * 8
    x0 = -(bx // 4) * 32 - bw
    x1 = (w_in_b - bx // 4 - w // 4) * 32 + bw
    y0 = -(by // 4) * 32 - bh
    y1 = (h_in_b - by // 4 - h // 4) * 32 + bh
    lo, hi = -(1 << 14) + 1, (1 << 14) - 1
    return max(x0, lo), min(x1, hi), max(y0, lo), min(y1, hi)


def planes(W, H, PAD, seed, smooth=True, dtype=np.uint8, maxv=255):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, maxv + 1, (H + 2 * PAD + 16, W + 2 * PAD + 16)).astype(np.float64)
    if smooth:
        k = np.ones(7) / 7.0
        base = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, base)
        base = np.apply_along_axis(lambda c: np.convolve(c, k, mode="same"), 0, base)
        base = (base - base.min()) / (base.max() - base.min()) * maxv
    ref = np.rint(base[8:8 + H + 2 * PAD, 8:8 + W + 2 * PAD]).astype(dtype)
    cur = np.rint(base[8 - 3:8 - 3 + H + 2 * PAD, 8 + 6:8 + 6 + W + 2 * PAD] + rng.normal(0, 1.5, ref.shape))
    return np.clip(cur, 0, maxv).astype(dtype), ref


def run_case(w, h, extensive, umh_range, lam, smooth, seed, dtype=np.uint8, maxv=255):
    W, H, PAD = 96, 64, 160
    cur, ref = planes(W, H, PAD, seed, smooth, dtype, maxv)
    ocur, oref = O.Plane(W, H, PAD, dtype=dtype), O.Plane(W, H, PAD, dtype=dtype)
    ocur.data[:], oref.data[:] = cur, ref
    xs, ys = np.arange(0, W - w + 1, w), np.arange(0, H - h + 1, h)
    blocks = np.zeros(len(xs) * len(ys), O.BLOCK_DTYPE)
    blocks["x"], blocks["y"] = np.tile(xs, len(ys)), np.repeat(ys, len(xs))
    n = len(blocks)
    rng = np.random.default_rng(seed + 1)
    nsub = 3 if extensive else 1
    counts = rng.integers(0, 4, (n, nsub))
    if extensive:
        counts[:, 0] = rng.integers(0, 2, n)
    offs = np.zeros(n * nsub + 1, np.uint32)
    offs[1:] = np.cumsum(counts.reshape(-1))
    preds = np.zeros(int(offs[-1]), O.CAND_DTYPE)
    mv = rng.integers(-30, 31, (len(preds), 2)) * 8 + (rng.random((len(preds), 2)) < 0.2) * rng.integers(-7, 8, (len(preds), 2))
    preds["mv_row"], preds["mv_col"] = mv[:, 0], mv[:, 1]
    pmv = (rng.integers(-20, 21, (n, 4)) * 4).astype(np.int16)
    thresh = rng.integers(0, w * h * 6, n).astype(np.uint32)
    got = O.full_pixel_me_blocks(ocur, oref, blocks, preds, offs, nsub, w, h, lam, pmv,
                                 thresh if extensive else None, umh_range)
    w_in_b, h_in_b = 2 * ((W + 7) >> 3), 2 * ((H + 7) >> 3)
    for i, b in enumerate(blocks):
        bx, by = int(b["x"]), int(b["y"])
        m = Model(cur, ref, PAD, bx, by, w, h, mv_range(w_in_b, h_in_b, bx, by, w, h), lam,
                  [(int(pmv[i, 0]), int(pmv[i, 1])), (int(pmv[i, 2]), int(pmv[i, 3]))])
        subsets = [[(int(q["mv_row"]), int(q["mv_col"])) for q in preds[offs[i * nsub + k]:offs[i * nsub + k + 1]]]
                   for k in range(nsub)]
        want = full_pixel_me(m, subsets, extensive, int(thresh[i]), umh_range)
        assert (int(got[i]["cost"]), int(got[i]["sad"])) == (want[0], want[1]), (i, got[i], want)
        if want[0] != EMPTY[0]:
            assert (int(got[i]["mv_row"]), int(got[i]["mv_col"])) == want[2], (i, got[i], want)


def test_non_extensive_diamond_matches_python_model():
    run_case(16, 16, False, 0, 500, True, seed=1)


def test_extensive_ladder_and_umh_match_python_model():
    run_case(16, 16, True, 24, 500, True, seed=2)


def test_noise_planes_zero_lambda_ties():
    run_case(8, 8, True, 24, 0, False, seed=3)


def test_other_sizes_and_hbd():
    run_case(32, 16, True, 16, 900, True, seed=4)
    run_case(16, 16, True, 24, 2000, True, seed=5, dtype=np.uint16, maxv=1023)


def test_search_never_worse_than_its_predictors():
    W, H, PAD, w, h = 96, 64, 160, 16, 16
    cur, ref = planes(W, H, PAD, 7)
    ocur, oref = O.Plane(W, H, PAD), O.Plane(W, H, PAD)
    ocur.data[:], oref.data[:] = cur, ref
    blocks = np.zeros(4, O.BLOCK_DTYPE)
    blocks["x"], blocks["y"] = [0, 16, 32, 48], [16, 16, 32, 32]
    preds = np.zeros(4 * 3, O.CAND_DTYPE)
    preds["mv_row"] = np.tile([0, 16, -24], 4)
    preds["mv_col"] = np.tile([0, -32, 40], 4)
    offs = (np.arange(5) * 3).astype(np.uint32)
    res = O.full_pixel_me_blocks(ocur, oref, blocks, preds, offs, 1, w, h, 300)
    cands = preds.copy()
    cands["block"] = np.repeat(np.arange(4), 3)
    _, cost = O.fullpel_candidates(ocur, oref, blocks, cands, w, h, False, 300)
    for i in range(4):
        assert res[i]["cost"] <= cost[3 * i:3 * i + 3].min()
    # the synthetic shift is (+6, -3) px: the searches should land there from nearby predictors
    assert ((res["mv_col"] == 48) & (res["mv_row"] == -24)).sum() >= 3
