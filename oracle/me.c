/*
 * oracle/me.c — restatement of the motion-search cost and scan semantics of rav1e
 * src/me.rs (get_mv_range :339-362, get_fullpel_mv_rd :1386-1409, compute_mv_rd
 * :1445-1461, full_search :1464-1509, get_mv_rate :1512-1523).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * `ILog::ilog` comes from v_frame 0.3.9 (off disk): bits - leading_zeros of the value's own width (0 for x == 0).
 * No reference test pins it ("parity unpinned" for the rate term); SAD itself is pinned.
 */
#include "oracle.h"

#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MI_SIZE 4          /* context/superblock_unit.rs:12 */
#define MV_LOW (-(1 << 14)) /* context/mod.rs:130-132 */
#define MV_UPP (1 << 14)

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static inline uint32_t ilog_i16(int16_t v) { /* v_frame ILog::ilog for i16: 16 - leading_zeros */
  if (v == 0) return 0;
  /* bits of the i16 pattern: a negative value (only i16::MIN survives the wrapping abs) has no
   * leading zeros -> 16 */
  return 32 - (uint32_t)__builtin_clz((uint32_t)(uint16_t)v);
}

/* me.rs:1516-1519 diff_to_rate */
static inline uint32_t diff_to_rate(int16_t diff, int allow_hp) {
  int16_t d = allow_hp ? diff : (int16_t)(diff >> 1); /* arithmetic shift on i16 */
  int16_t a = (int16_t)(d < 0 ? -d : d);
  return 2 * ilog_i16(a);
}

/* me.rs:1512-1523 */
uint32_t orc_get_mv_rate(orc_mv a, orc_mv b, int allow_hp) {
  return diff_to_rate((int16_t)(a.row - b.row), allow_hp) +
         diff_to_rate((int16_t)(a.col - b.col), allow_hp);
}

/* me.rs:1455-1460 */
uint64_t orc_mv_cost(uint32_t sad, orc_mv cand, orc_mv pmv0, orc_mv pmv1, uint32_t lambda,
                     int allow_hp) {
  uint32_t rate1 = orc_get_mv_rate(cand, pmv0, allow_hp);
  uint32_t rate2 = orc_get_mv_rate(cand, pmv1, allow_hp);
  uint32_t rate = rate1 < rate2 + 1 ? rate1 : rate2 + 1;
  return 256 * (uint64_t)sad + (uint64_t)rate * (uint64_t)lambda;
}

/* me.rs:339-362 — bo in 4x4 block units, blk_w/h in pixels; result in 1/8 pel. */
void orc_get_mv_range(int w_in_b, int h_in_b, int bo_x, int bo_y, int blk_w, int blk_h,
                      int *mvx_min, int *mvx_max, int *mvy_min, int *mvy_max) {
  int border_w = 128 + blk_w * 8;
  int border_h = 128 + blk_h * 8;
  int x_min = -bo_x * (8 * MI_SIZE) - border_w;
  int x_max = ((w_in_b - bo_x) - (blk_w / MI_SIZE)) * (8 * MI_SIZE) + border_w;
  int y_min = -bo_y * (8 * MI_SIZE) - border_h;
  int y_max = ((h_in_b - bo_y) - (blk_h / MI_SIZE)) * (8 * MI_SIZE) + border_h;
  *mvx_min = x_min > MV_LOW + 1 ? x_min : MV_LOW + 1;
  *mvx_max = x_max < MV_UPP - 1 ? x_max : MV_UPP - 1;
  *mvy_min = y_min > MV_LOW + 1 ? y_min : MV_LOW + 1;
  *mvy_max = y_max < MV_UPP - 1 ? y_max : MV_UPP - 1;
}

static inline uint32_t dist_any(const void *org, ptrdiff_t os, const void *ref, ptrdiff_t rs,
                                int bpp, int w, int h, int use_satd) {
  if (bpp == 1)
    return use_satd ? orc_get_satd_u8((const uint8_t *)org, os, (const uint8_t *)ref, rs, w, h)
                    : orc_get_sad_u8((const uint8_t *)org, os, (const uint8_t *)ref, rs, w, h);
  return use_satd ? orc_get_satd_u16((const uint16_t *)org, os, (const uint16_t *)ref, rs, w, h)
                  : orc_get_sad_u16((const uint16_t *)org, os, (const uint16_t *)ref, rs, w, h);
}

static inline const void *px_at(const void *p0, ptrdiff_t stride, int bpp, int x, int y) {
  return (const uint8_t *)p0 + ((ptrdiff_t)y * stride + x) * bpp;
}

/* me.rs:1464-1509.  vert_windows(h).step_by(step) x horz_windows(w).step_by(step): rows
 * y_lo, y_lo+step, ... <= y_hi outer; columns x_lo ... <= x_hi inner; strict `<` keeps
 * the first minimum in scan order. */
orc_me_result orc_full_search(const void *org, ptrdiff_t org_stride, const void *ref0,
                              ptrdiff_t ref_stride, int bpp, int x_lo, int x_hi, int y_lo,
                              int y_hi, int w, int h, int po_x, int po_y, int step,
                              uint32_t lambda, orc_mv pmv0, orc_mv pmv1, int allow_hp) {
  orc_me_result best;
  best.cost = UINT64_MAX;
  best.sad = UINT32_MAX;
  best.mv.row = 0;
  best.mv.col = 0;
  for (int y = y_lo; y <= y_hi; y += step) {
    for (int x = x_lo; x <= x_hi; x += step) {
      orc_mv mv;
      mv.row = (int16_t)(8 * (int16_t)((int16_t)y - (int16_t)po_y));
      mv.col = (int16_t)(8 * (int16_t)((int16_t)x - (int16_t)po_x));
      uint32_t sad = dist_any(org, org_stride, px_at(ref0, ref_stride, bpp, x, y), ref_stride,
                              bpp, w, h, 0);
      uint64_t cost = orc_mv_cost(sad, mv, pmv0, pmv1, lambda, allow_hp);
      if (cost < best.cost) {
        best.cost = cost;
        best.sad = sad;
        best.mv = mv;
      }
    }
  }
  return best;
}

/* me.rs:1386-1409 applied to a list.  The block offset bo = (x/4, y/4). */
void orc_fullpel_candidates(const void *cur0, ptrdiff_t cur_stride, const void *ref0,
                            ptrdiff_t ref_stride, int bpp, int frame_w_in_b, int frame_h_in_b,
                            const orc_block *blocks, const orc_cand *cands, size_t n, int w,
                            int h, int use_satd, uint32_t lambda, const orc_mv *pmv,
                            int allow_hp, uint32_t *out_sad, uint64_t *out_cost, int threads) {
  (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : orc_num_threads())
  for (ptrdiff_t i = 0; i < (ptrdiff_t)n; i++) {
    const orc_cand c = cands[i];
    const orc_block b = blocks[c.block];
    int mvx_min, mvx_max, mvy_min, mvy_max;
    orc_get_mv_range(frame_w_in_b, frame_h_in_b, b.x / MI_SIZE, b.y / MI_SIZE, w, h, &mvx_min,
                     &mvx_max, &mvy_min, &mvy_max);
    uint32_t sad = UINT32_MAX;
    uint64_t cost = UINT64_MAX;
    if (!(c.mv_col < mvx_min || c.mv_col > mvx_max || c.mv_row < mvy_min || c.mv_row > mvy_max)) {
      /* Rust `/` on i16 truncates toward zero, as C does. */
      int rx = b.x + c.mv_col / 8, ry = b.y + c.mv_row / 8;
      sad = dist_any(px_at(cur0, cur_stride, bpp, b.x, b.y), cur_stride,
                     px_at(ref0, ref_stride, bpp, rx, ry), ref_stride, bpp, w, h, use_satd);
      orc_mv z = {0, 0};
      orc_mv cm = {c.mv_row, c.mv_col};
      cost = orc_mv_cost(sad, cm, pmv ? pmv[2 * c.block] : z, pmv ? pmv[2 * c.block + 1] : z,
                         lambda, allow_hp);
    }
    if (out_sad) out_sad[i] = sad;
    if (out_cost) out_cost[i] = cost;
  }
}

/* me.rs:822-846 (ssdec = 0 form): window = po +- range clamped to the mv range, step. */
void orc_full_search_blocks(const void *cur0, ptrdiff_t cur_stride, const void *ref0,
                            ptrdiff_t ref_stride, int bpp, int frame_w_in_b, int frame_h_in_b,
                            const orc_block *blocks, size_t nblocks, int w, int h, int range_x,
                            int range_y, int step, uint32_t lambda, int allow_hp,
                            orc_me_result *out, int threads) {
  (void)threads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : orc_num_threads())
  for (ptrdiff_t i = 0; i < (ptrdiff_t)nblocks; i++) {
    const orc_block b = blocks[i];
    int mvx_min, mvx_max, mvy_min, mvy_max;
    orc_get_mv_range(frame_w_in_b, frame_h_in_b, b.x / MI_SIZE, b.y / MI_SIZE, w, h, &mvx_min,
                     &mvx_max, &mvy_min, &mvy_max);
    int lo, hi;
    lo = -range_x > mvx_min / 8 ? -range_x : mvx_min / 8;
    hi = range_x < mvx_max / 8 ? range_x : mvx_max / 8;
    int x_lo = b.x + lo, x_hi = b.x + hi;
    lo = -range_y > mvy_min / 8 ? -range_y : mvy_min / 8;
    hi = range_y < mvy_max / 8 ? range_y : mvy_max / 8;
    int y_lo = b.y + lo, y_hi = b.y + hi;
    orc_mv z = {0, 0};
    out[i] = orc_full_search(px_at(cur0, cur_stride, bpp, b.x, b.y), cur_stride, ref0, ref_stride,
                             bpp, x_lo, x_hi, y_lo, y_hi, w, h, b.x, b.y, step, lambda, z, z,
                             allow_hp);
  }
}

/* get_subpel_mv_rd (me.rs:1411-1442) over a candidate list: range check, predict_inter_single
 * (predict.rs:304-336 -> put_8tap with fi.default_filter on the luma plane) into a scratch of
 * mc_w x mc_h = next_power_of_two(w) x ((h+1)&!1) (me.rs:1322-1324), then compute_mv_rd against it. */
void orc_subpel_candidates(const void *cur0, ptrdiff_t cur_stride, const void *ref0,
                           ptrdiff_t ref_stride, int bpp, int frame_w_in_b, int frame_h_in_b,
                           const orc_block *blocks, const orc_cand *cands, size_t n, int w, int h,
                           int use_satd, uint32_t lambda, const orc_mv *pmv, int allow_hp,
                           int filter_mode, int bit_depth, uint32_t *out_sad, uint64_t *out_cost,
                           int threads) {
  int mc_w = 1;
  while (mc_w < w) mc_w <<= 1;
  const int mc_h = (h + 1) & ~1;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : orc_num_threads())
  for (ptrdiff_t i = 0; i < (ptrdiff_t)n; i++) {
    const orc_cand c = cands[i];
    const orc_block b = blocks[c.block];
    int mvx_min, mvx_max, mvy_min, mvy_max;
    orc_get_mv_range(frame_w_in_b, frame_h_in_b, b.x / MI_SIZE, b.y / MI_SIZE, w, h, &mvx_min,
                     &mvx_max, &mvy_min, &mvy_max);
    uint32_t sad = UINT32_MAX;
    uint64_t cost = UINT64_MAX;
    if (!(c.mv_col < mvx_min || c.mv_col > mvx_max || c.mv_row < mvy_min || c.mv_row > mvy_max)) {
      uint16_t tmp[128 * 128];
      int ro, co, rf, cf;
      orc_get_mv_params(c.mv_row, c.mv_col, 0, 0, &ro, &co, &rf, &cf);
      orc_put_8tap(tmp, mc_w, px_at(ref0, ref_stride, bpp, b.x + co, b.y + ro), ref_stride, bpp, mc_w,
                   mc_h, cf, rf, filter_mode, filter_mode, bit_depth);
      sad = dist_any(px_at(cur0, cur_stride, bpp, b.x, b.y), cur_stride, tmp, mc_w, bpp, w, h, use_satd);
      orc_mv z = {0, 0};
      orc_mv cm = {c.mv_row, c.mv_col};
      cost = orc_mv_cost(sad, cm, pmv ? pmv[2 * c.block] : z, pmv ? pmv[2 * c.block + 1] : z, lambda,
                         allow_hp);
    }
    if (out_sad) out_sad[i] = sad;
    if (out_cost) out_cost[i] = cost;
  }
}

/* ------------------------------------------------------------------------------------------
 * Search stages of full_pixel_me (me.rs:692-856): get_best_predictor :884-909,
 * fullpel_diamond_search :955-998, hexagon_search :1055-1135, uneven_multi_hex_search
 * :1170-1303, driven per block like the `try_cands` closure (:722-762) and its callers.
 * MotionVector arithmetic is i16 (mc.rs:57-100; release builds wrap).  No stored vector in the
 * reference pins these stages ("parity unpinned"); they are compositions of get_fullpel_mv_rd,
 * whose SAD is pinned, with the strict-`<` update rule restated line by line.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void *org; /* block pixel (0,0) in cur */
  ptrdiff_t org_stride;
  const void *ref0;
  ptrdiff_t ref_stride;
  int bpp, w, h, po_x, po_y;
  int mvx_min, mvx_max, mvy_min, mvy_max;
  uint32_t lambda;
  orc_mv pmv0, pmv1;
  int allow_hp;
} srch;

static inline orc_mv mv_add(orc_mv a, orc_mv b) {
  orc_mv r = {(int16_t)(a.row + b.row), (int16_t)(a.col + b.col)};
  return r;
}
static inline orc_mv mv_mul(orc_mv a, int16_t k) {
  orc_mv r = {(int16_t)(a.row * k), (int16_t)(a.col * k)};
  return r;
}
static inline orc_mv mv_shl(orc_mv a, int s) {
  orc_mv r = {(int16_t)((int)a.row * (1 << s)), (int16_t)((int)a.col * (1 << s))};
  return r;
}
static inline orc_me_result res_empty(void) { /* MotionSearchResult::empty(), me.rs:111-116 */
  orc_me_result r;
  r.cost = UINT64_MAX;
  r.sad = UINT32_MAX;
  r.mv.row = 0;
  r.mv.col = 0;
  return r;
}

/* get_fullpel_mv_rd, me.rs:1386-1409 (use_satd = false in every full-pel stage) */
static orc_me_result fullpel_rd(const srch *s, orc_mv mv) {
  orc_me_result r = res_empty();
  r.mv = mv;
  if (mv.col < s->mvx_min || mv.col > s->mvx_max || mv.row < s->mvy_min || mv.row > s->mvy_max)
    return r;
  r.sad = dist_any(s->org, s->org_stride,
                   px_at(s->ref0, s->ref_stride, s->bpp, s->po_x + mv.col / 8, s->po_y + mv.row / 8),
                   s->ref_stride, s->bpp, s->w, s->h, 0);
  r.cost = orc_mv_cost(r.sad, mv, s->pmv0, s->pmv1, s->lambda, s->allow_hp);
  return r;
}

/* me.rs:884-909 */
static orc_me_result best_predictor(const srch *s, const orc_cand *preds, size_t n) {
  orc_me_result best = res_empty();
  for (size_t i = 0; i < n; i++) {
    orc_mv mv = {preds[i].mv_row, preds[i].mv_col};
    orc_me_result rd = fullpel_rd(s, mv);
    if (rd.cost < best.cost) best = rd;
  }
  return best;
}

#define FP(c, r) {(int16_t)((r) * 8), (int16_t)((c) * 8)} /* search_pattern!: {row, col} << 3 */
static const orc_mv DIAMOND_R1[4] = {FP(0, 1), FP(1, 0), FP(0, -1), FP(-1, 0)}; /* me.rs:944-947 */
static const orc_mv HEXAGON[6] = {FP(0, -2), FP(2, -1), FP(2, 1), FP(0, 2), FP(-2, 1), FP(-2, -1)}; /* :1022-1025 */
static const orc_mv SQUARE_REFINE[8] = {FP(-1, 1), FP(0, 1),  FP(1, 1),  FP(-1, 0),
                                        FP(1, 0),  FP(-1, -1), FP(0, -1), FP(1, -1)}; /* :1035-1038 */
static const orc_mv UMH[16] = {FP(-2, 4), FP(-1, 4), FP(0, 4),  FP(1, 4),  FP(2, 4),  FP(3, 2),
                               FP(4, 0),  FP(3, -2), FP(2, -4), FP(1, -4), FP(0, -4), FP(-1, -4),
                               FP(-2, -4), FP(3, -2), FP(-4, 0), FP(-3, 2)}; /* :1153-1156 */

/* me.rs:955-998 */
static void fullpel_diamond(const srch *s, orc_me_result *current) {
  int radius_log2 = 1;
  const int end_log2 = 0;
  for (;;) {
    orc_me_result best_cand = res_empty();
    for (int k = 0; k < 4; k++) {
      orc_me_result rd = fullpel_rd(s, mv_add(current->mv, mv_shl(DIAMOND_R1[k], radius_log2)));
      if (rd.cost < best_cand.cost) best_cand = rd;
    }
    if (current->cost <= best_cand.cost) {
      if (radius_log2 == end_log2) break;
      radius_log2--;
    } else {
      *current = best_cand;
    }
  }
}

/* me.rs:1055-1135 */
static void hexagon(const srch *s, orc_me_result *current) {
  int best_idx = 0;
  orc_me_result best_cand = res_empty();
  for (int i = 0; i < 6; i++) {
    orc_me_result rd = fullpel_rd(s, mv_add(current->mv, HEXAGON[i]));
    if (rd.cost < best_cand.cost) {
      best_idx = i;
      best_cand = rd;
    }
  }
  while (best_cand.cost < current->cost) {
    *current = best_cand;
    best_cand = res_empty();
    const int center_idx = best_idx;
    for (int off = 5; off <= 7; off++) {
      const int i = (center_idx + off) % 6;
      orc_me_result rd = fullpel_rd(s, mv_add(current->mv, HEXAGON[i]));
      if (rd.cost < best_cand.cost) {
        best_idx = i;
        best_cand = rd;
      }
    }
  }
  best_cand = res_empty();
  for (int k = 0; k < 8; k++) {
    orc_me_result rd = fullpel_rd(s, mv_add(current->mv, SQUARE_REFINE[k]));
    if (rd.cost < best_cand.cost) best_cand = rd;
  }
  if (best_cand.cost < current->cost) *current = best_cand;
}

/* me.rs:1170-1303.  Quirks kept as they are: the "horizontal" line steps the ROW component
 * (:1195-1198), and the 5x5 stage adds raw {row, col} in 1/8 pel, not full pixels (:1240-1246). */
static void umh(const srch *s, orc_me_result *current, int me_range) {
  orc_mv center = current->mv;
  for (int i = 1; i <= me_range; i += 2) {
    static const orc_mv line[2] = {FP(0, -1), FP(0, 1)};
    for (int k = 0; k < 2; k++) {
      orc_me_result rd = fullpel_rd(s, mv_add(center, mv_mul(line[k], (int16_t)i)));
      if (rd.cost < current->cost) *current = rd;
    }
  }
  for (int i = 1; i <= (me_range >> 1); i += 2) {
    static const orc_mv line[2] = {FP(-1, 0), FP(1, 0)};
    for (int k = 0; k < 2; k++) {
      orc_me_result rd = fullpel_rd(s, mv_add(center, mv_mul(line[k], (int16_t)i)));
      if (rd.cost < current->cost) *current = rd;
    }
  }
  center = current->mv;
  for (int row = -2; row <= 2; row++)
    for (int col = -2; col <= 2; col++) {
      if (row == 0 && col == 0) continue;
      orc_mv off = {(int16_t)row, (int16_t)col};
      orc_me_result rd = fullpel_rd(s, mv_add(center, off));
      if (rd.cost < current->cost) *current = rd;
    }
  center = current->mv;
  const int iterations = me_range >> 2;
  for (int i = 1; i <= iterations; i++)
    for (int k = 0; k < 16; k++) {
      orc_me_result rd = fullpel_rd(s, mv_add(center, mv_mul(UMH[k], (int16_t)i)));
      if (rd.cost < current->cost) *current = rd;
    }
  hexagon(s, current);
}

/* full_pixel_me (me.rs:692-856) for every block, full resolution (ssdec = 0), without the final
 * exhaustive grid (orc_full_search_blocks).  Predictor subsets come from the caller
 * (get_subset_predictors reads neighbouring blocks' results): nsubsets = 1 -> `all_mvs`
 * (non-extensive, :757-759); nsubsets = 3 -> {median (skipped when empty, :774), subset_b,
 * subset_c} with the early exits `best.rd.sad < thresh` (:777-790) and the UMH stage (:794-812).
 * subset_offsets: nblocks * nsubsets + 1 offsets into preds.  Where the reference would hit
 * `assert!(!current.is_empty())` (no candidate in range at all) the result stays empty. */
void orc_full_pixel_me_blocks(const void *cur0, ptrdiff_t cur_stride, const void *ref0,
                              ptrdiff_t ref_stride, int bpp, int frame_w_in_b, int frame_h_in_b,
                              const orc_block *blocks, size_t nblocks, const orc_cand *preds,
                              const uint32_t *subset_offsets, int nsubsets, const orc_mv *pmv,
                              const uint32_t *thresh, int w, int h, uint32_t lambda, int allow_hp,
                              int umh_range, orc_me_result *out, int threads) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 0 ? threads : orc_num_threads())
  for (ptrdiff_t i = 0; i < (ptrdiff_t)nblocks; i++) {
    const orc_block b = blocks[i];
    srch s;
    s.org = px_at(cur0, cur_stride, bpp, b.x, b.y);
    s.org_stride = cur_stride;
    s.ref0 = ref0;
    s.ref_stride = ref_stride;
    s.bpp = bpp;
    s.w = w;
    s.h = h;
    s.po_x = b.x;
    s.po_y = b.y;
    orc_get_mv_range(frame_w_in_b, frame_h_in_b, b.x / MI_SIZE, b.y / MI_SIZE, w, h, &s.mvx_min,
                     &s.mvx_max, &s.mvy_min, &s.mvy_max);
    s.lambda = lambda;
    orc_mv z = {0, 0};
    s.pmv0 = pmv ? pmv[2 * i] : z;
    s.pmv1 = pmv ? pmv[2 * i + 1] : z;
    s.allow_hp = allow_hp;
    orc_me_result best = res_empty();
    int done = 0;
    for (int k = 0; k < nsubsets && !done; k++) {
      const uint32_t lo = subset_offsets[i * nsubsets + k], hi = subset_offsets[i * nsubsets + k + 1];
      if (nsubsets == 3 && k == 0 && lo == hi) continue; /* `if let Some(median)` */
      orc_me_result r = best_predictor(&s, preds + lo, hi - lo);
      fullpel_diamond(&s, &r);
      if (r.cost < best.cost) best = r;
      if (nsubsets == 3 && best.sad < thresh[i]) done = 1;
    }
    if (nsubsets == 3 && !done && umh_range > 0 && best.cost != UINT64_MAX) umh(&s, &best, umh_range);
    out[i] = best;
  }
}

/* subpel_diamond_search (me.rs:1311-1383) for every block: `current` starts as the caller's result
 * (the full-pel stage's winner with its cost); radius 1/2 pel down to 1/4 pel (1/8 with
 * allow_high_precision_mv); per radius the four DIAMOND_R1_PATTERN_SUBPEL candidates (me.rs:931-934:
 * row+1, col+1, row-1, col-1 in that order, first minimum kept) are evaluated with get_subpel_mv_rd and
 * the centre moves while a candidate is strictly better.  results: in / out, one per block. */
void orc_subpel_diamond_search_blocks(const void *cur0, ptrdiff_t cur_stride, const void *ref0, ptrdiff_t ref_stride,
                                      int bpp, int frame_w_in_b, int frame_h_in_b, const orc_block *blocks, size_t n,
                                      int w, int h, int use_satd, uint32_t lambda, const orc_mv *pmv, int allow_hp,
                                      int filter_mode, int bit_depth, orc_me_result *results, int threads) {
  static const int8_t pat[4][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}}; /* {row, col} */
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : orc_num_threads())
  for (ptrdiff_t i = 0; i < (ptrdiff_t)n; i++) {
    orc_me_result cur = results[i];
    int radius_log2 = 2;
    const int end_log2 = allow_hp ? 0 : 1;
    for (;;) {
      orc_me_result best;
      best.cost = UINT64_MAX;
      best.sad = UINT32_MAX;
      best.mv.row = best.mv.col = 0;
      for (int k = 0; k < 4; k++) {
        orc_cand c;
        c.block = (uint32_t)i;
        c.mv_row = (int16_t)(cur.mv.row + (int16_t)(pat[k][0] << radius_log2)); /* i16 wrapping add */
        c.mv_col = (int16_t)(cur.mv.col + (int16_t)(pat[k][1] << radius_log2));
        uint32_t sad;
        uint64_t cost;
        orc_subpel_candidates(cur0, cur_stride, ref0, ref_stride, bpp, frame_w_in_b, frame_h_in_b, blocks, &c, 1, w, h,
                              use_satd, lambda, pmv, allow_hp, filter_mode, bit_depth, &sad, &cost, 1);
        if (cost < best.cost) {
          best.cost = cost;
          best.sad = sad;
          best.mv.row = c.mv_row;
          best.mv.col = c.mv_col;
        }
      }
      if (cur.cost <= best.cost) {
        if (radius_log2 == end_log2) break;
        radius_log2--;
      } else {
        cur = best;
      }
    }
    results[i] = cur;
  }
}
