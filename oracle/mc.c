/*
 * oracle/mc.c — restatement of rav1e src/mc.rs `rust::put_8tap` (:250-353), `prep_8tap`
 * (:360-451), `mc_avg` (:454-479), `get_filter` (:238-247) and SUBPEL_FILTERS (:110-219),
 * plus get_mv_params (src/predict.rs:284-297).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Pinning: the reference holds no stored vectors for MC (only asm==rust random tests that need
 * rustc, asm/x86/mc.rs:624-833): "parity unpinned".  Cross-checks in tests/test_oracle_mc.py:
 * the filters are the AV1-normative tables (each row sums to 128), integer positions copy,
 * the H-only path's double rounding differs from single rounding exactly as stated
 * (mc.rs:299-307), compound average of two identical preps reproduces put (AV1 spec 7.11.3.1).
 */
#include "oracle.h"

/* mc.rs:110-219 (AV1 spec "Subpel_Filters"): [bank][phase][tap]; banks 0 REGULAR, 1 SMOOTH,
 * 2 SHARP, 3 BILINEAR, 4/5 the 4-tap REGULAR/SMOOTH variants used when a dimension is <= 4. */
static const int16_t SUBPEL_FILTERS[6][16][8] = {
  {
    {0, 0, 0, 128, 0, 0, 0, 0},
    {0, 2, -6, 126, 8, -2, 0, 0},
    {0, 2, -10, 122, 18, -4, 0, 0},
    {0, 2, -12, 116, 28, -8, 2, 0},
    {0, 2, -14, 110, 38, -10, 2, 0},
    {0, 2, -14, 102, 48, -12, 2, 0},
    {0, 2, -16, 94, 58, -12, 2, 0},
    {0, 2, -14, 84, 66, -12, 2, 0},
    {0, 2, -14, 76, 76, -14, 2, 0},
    {0, 2, -12, 66, 84, -14, 2, 0},
    {0, 2, -12, 58, 94, -16, 2, 0},
    {0, 2, -12, 48, 102, -14, 2, 0},
    {0, 2, -10, 38, 110, -14, 2, 0},
    {0, 2, -8, 28, 116, -12, 2, 0},
    {0, 0, -4, 18, 122, -10, 2, 0},
    {0, 0, -2, 8, 126, -6, 2, 0},
  },
  {
    {0, 0, 0, 128, 0, 0, 0, 0},
    {0, 2, 28, 62, 34, 2, 0, 0},
    {0, 0, 26, 62, 36, 4, 0, 0},
    {0, 0, 22, 62, 40, 4, 0, 0},
    {0, 0, 20, 60, 42, 6, 0, 0},
    {0, 0, 18, 58, 44, 8, 0, 0},
    {0, 0, 16, 56, 46, 10, 0, 0},
    {0, -2, 16, 54, 48, 12, 0, 0},
    {0, -2, 14, 52, 52, 14, -2, 0},
    {0, 0, 12, 48, 54, 16, -2, 0},
    {0, 0, 10, 46, 56, 16, 0, 0},
    {0, 0, 8, 44, 58, 18, 0, 0},
    {0, 0, 6, 42, 60, 20, 0, 0},
    {0, 0, 4, 40, 62, 22, 0, 0},
    {0, 0, 4, 36, 62, 26, 0, 0},
    {0, 0, 2, 34, 62, 28, 2, 0},
  },
  {
    {0, 0, 0, 128, 0, 0, 0, 0},
    {-2, 2, -6, 126, 8, -2, 2, 0},
    {-2, 6, -12, 124, 16, -6, 4, -2},
    {-2, 8, -18, 120, 26, -10, 6, -2},
    {-4, 10, -22, 116, 38, -14, 6, -2},
    {-4, 10, -22, 108, 48, -18, 8, -2},
    {-4, 10, -24, 100, 60, -20, 8, -2},
    {-4, 10, -24, 90, 70, -22, 10, -2},
    {-4, 12, -24, 80, 80, -24, 12, -4},
    {-2, 10, -22, 70, 90, -24, 10, -4},
    {-2, 8, -20, 60, 100, -24, 10, -4},
    {-2, 8, -18, 48, 108, -22, 10, -4},
    {-2, 6, -14, 38, 116, -22, 10, -4},
    {-2, 6, -10, 26, 120, -18, 8, -2},
    {-2, 4, -6, 16, 124, -12, 6, -2},
    {0, 2, -2, 8, 126, -6, 2, -2},
  },
  {
    {0, 0, 0, 128, 0, 0, 0, 0},
    {0, 0, 0, 120, 8, 0, 0, 0},
    {0, 0, 0, 112, 16, 0, 0, 0},
    {0, 0, 0, 104, 24, 0, 0, 0},
    {0, 0, 0, 96, 32, 0, 0, 0},
    {0, 0, 0, 88, 40, 0, 0, 0},
    {0, 0, 0, 80, 48, 0, 0, 0},
    {0, 0, 0, 72, 56, 0, 0, 0},
    {0, 0, 0, 64, 64, 0, 0, 0},
    {0, 0, 0, 56, 72, 0, 0, 0},
    {0, 0, 0, 48, 80, 0, 0, 0},
    {0, 0, 0, 40, 88, 0, 0, 0},
    {0, 0, 0, 32, 96, 0, 0, 0},
    {0, 0, 0, 24, 104, 0, 0, 0},
    {0, 0, 0, 16, 112, 0, 0, 0},
    {0, 0, 0, 8, 120, 0, 0, 0},
  },
  {
    {0, 0, 0, 128, 0, 0, 0, 0},
    {0, 0, -4, 126, 8, -2, 0, 0},
    {0, 0, -8, 122, 18, -4, 0, 0},
    {0, 0, -10, 116, 28, -6, 0, 0},
    {0, 0, -12, 110, 38, -8, 0, 0},
    {0, 0, -12, 102, 48, -10, 0, 0},
    {0, 0, -14, 94, 58, -10, 0, 0},
    {0, 0, -12, 84, 66, -10, 0, 0},
    {0, 0, -12, 76, 76, -12, 0, 0},
    {0, 0, -10, 66, 84, -12, 0, 0},
    {0, 0, -10, 58, 94, -14, 0, 0},
    {0, 0, -10, 48, 102, -12, 0, 0},
    {0, 0, -8, 38, 110, -12, 0, 0},
    {0, 0, -6, 28, 116, -10, 0, 0},
    {0, 0, -4, 18, 122, -8, 0, 0},
    {0, 0, -2, 8, 126, -4, 0, 0},
  },
  {
    {0, 0, 0, 128, 0, 0, 0, 0},
    {0, 0, 30, 62, 34, 2, 0, 0},
    {0, 0, 26, 62, 36, 4, 0, 0},
    {0, 0, 22, 62, 40, 4, 0, 0},
    {0, 0, 20, 60, 42, 6, 0, 0},
    {0, 0, 18, 58, 44, 8, 0, 0},
    {0, 0, 16, 56, 46, 10, 0, 0},
    {0, 0, 14, 54, 48, 12, 0, 0},
    {0, 0, 12, 52, 52, 12, 0, 0},
    {0, 0, 12, 48, 54, 14, 0, 0},
    {0, 0, 10, 46, 56, 16, 0, 0},
    {0, 0, 8, 44, 58, 18, 0, 0},
    {0, 0, 6, 42, 60, 20, 0, 0},
    {0, 0, 4, 40, 62, 22, 0, 0},
    {0, 0, 4, 36, 62, 26, 0, 0},
    {0, 0, 2, 34, 62, 30, 0, 0},
  },
};

/* mc.rs:238-247 get_filter */
void orc_get_filter(int mode, int frac, int length, int32_t out[8]) {
  int idx = (mode == 3 || length > 4) ? mode : (mode < 1 ? mode : 1) + 4;
  for (int k = 0; k < 8; k++) out[k] = SUBPEL_FILTERS[idx][frac][k];
}

static inline int32_t round_shift_i32(int32_t v, int bit) { /* v_frame round_shift */
  return (v + ((1 << bit) >> 1)) >> bit;
}
static inline int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : v > hi ? hi : v; }

#define PIXEL uint8_t
#define SFX(name) name##_u8
#include "mc_impl.h"
#undef PIXEL
#undef SFX
#define PIXEL uint16_t
#define SFX(name) name##_u16
#include "mc_impl.h"
#undef PIXEL
#undef SFX

void orc_put_8tap(void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride, int bpp,
                  int w, int h, int col_frac, int row_frac, int mode_x, int mode_y, int bit_depth) {
  if (bpp == 1)
    put_8tap_u8((uint8_t *)dst, dst_stride, (const uint8_t *)src, src_stride, w, h, col_frac,
                row_frac, mode_x, mode_y, bit_depth);
  else
    put_8tap_u16((uint16_t *)dst, dst_stride, (const uint16_t *)src, src_stride, w, h, col_frac,
                 row_frac, mode_x, mode_y, bit_depth);
}

void orc_prep_8tap(int16_t *tmp, const void *src, ptrdiff_t src_stride, int bpp, int w, int h,
                   int col_frac, int row_frac, int mode_x, int mode_y, int bit_depth) {
  if (bpp == 1)
    prep_8tap_u8(tmp, (const uint8_t *)src, src_stride, w, h, col_frac, row_frac, mode_x, mode_y,
                 bit_depth);
  else
    prep_8tap_u16(tmp, (const uint16_t *)src, src_stride, w, h, col_frac, row_frac, mode_x, mode_y,
                  bit_depth);
}

void orc_mc_avg(void *dst, ptrdiff_t dst_stride, int bpp, const int16_t *tmp1, const int16_t *tmp2,
                int w, int h, int bit_depth) {
  if (bpp == 1)
    mc_avg_u8((uint8_t *)dst, dst_stride, tmp1, tmp2, w, h, bit_depth);
  else
    mc_avg_u16((uint16_t *)dst, dst_stride, tmp1, tmp2, w, h, bit_depth);
}

/* predict.rs:284-297 get_mv_params for a luma/chroma plane with decimation (xdec, ydec). */
void orc_get_mv_params(int mv_row, int mv_col, int xdec, int ydec, int *row_off, int *col_off,
                       int *row_frac, int *col_frac) {
  *row_off = mv_row >> (3 + ydec);
  *col_off = mv_col >> (3 + xdec);
  *row_frac = (int)(((uint32_t)mv_row << (1 - ydec)) & 0xf);
  *col_frac = (int)(((uint32_t)mv_col << (1 - xdec)) & 0xf);
}

/* Batched predict_inter_single (predict.rs:304-336) into packed w x h blocks.
 * kind: 0 put (dst pixels), 1 prep (int16). */
void orc_mc_blocks(const void *ref0, ptrdiff_t ref_stride, int bpp, const orc_block *blocks,
                   const orc_mv *mvs, size_t n, int w, int h, int mode_x, int mode_y, int bit_depth,
                   int xdec, int ydec, int kind, void *out, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : orc_num_threads())
  for (ptrdiff_t i = 0; i < (ptrdiff_t)n; i++) {
    int ro, co, rf, cf;
    orc_get_mv_params(mvs[i].row, mvs[i].col, xdec, ydec, &ro, &co, &rf, &cf);
    const uint8_t *src = (const uint8_t *)ref0 +
                         ((ptrdiff_t)(blocks[i].y + ro) * ref_stride + blocks[i].x + co) * bpp;
    if (kind == 0)
      orc_put_8tap((uint8_t *)out + (size_t)i * w * h * bpp, w, src, ref_stride, bpp, w, h, cf, rf,
                   mode_x, mode_y, bit_depth);
    else
      orc_prep_8tap((int16_t *)out + (size_t)i * w * h, src, ref_stride, bpp, w, h, cf, rf, mode_x,
                    mode_y, bit_depth);
  }
}
