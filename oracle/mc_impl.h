/* Included twice by mc.c with PIXEL = uint8_t / uint16_t.  Strides in elements. */

/* mc.rs:224-236 run_filter */
static inline int32_t SFX(run_filter)(const PIXEL *src, ptrdiff_t stride, const int32_t f[8]) {
  int32_t s = 0;
  for (int i = 0; i < 8; i++) s += f[i] * (int32_t)src[i * stride];
  return s;
}
static inline int32_t SFX(run_filter_i16)(const int16_t *src, ptrdiff_t stride, const int32_t f[8]) {
  int32_t s = 0;
  for (int i = 0; i < 8; i++) s += f[i] * (int32_t)src[i * stride];
  return s;
}

/* mc.rs:250-353 */
static void SFX(put_8tap)(PIXEL *dst, ptrdiff_t dst_stride, const PIXEL *src, ptrdiff_t ref_stride,
                          int width, int height, int col_frac, int row_frac, int mode_x, int mode_y,
                          int bit_depth) {
  int32_t y_filter[8], x_filter[8];
  orc_get_filter(mode_y, row_frac, height, y_filter);
  orc_get_filter(mode_x, col_frac, width, x_filter);
  const int32_t max_sample_val = (1 << bit_depth) - 1;
  const int ib = 4 - (bit_depth == 12 ? 2 : 0); /* intermediate_bits */
  if (col_frac == 0 && row_frac == 0) {
    for (int r = 0; r < height; r++)
      for (int c = 0; c < width; c++) dst[r * dst_stride + c] = src[r * ref_stride + c];
  } else if (col_frac == 0) {
    const PIXEL *s = src - 3 * ref_stride; /* go_up(3) */
    for (int r = 0; r < height; r++)
      for (int c = 0; c < width; c++)
        dst[r * dst_stride + c] = (PIXEL)clampi(
            round_shift_i32(SFX(run_filter)(s + r * ref_stride + c, ref_stride, y_filter), 7), 0,
            max_sample_val);
  } else if (row_frac == 0) {
    const PIXEL *s = src - 3; /* go_left(3) */
    for (int r = 0; r < height; r++)
      for (int c = 0; c < width; c++)
        dst[r * dst_stride + c] = (PIXEL)clampi(
            round_shift_i32(round_shift_i32(SFX(run_filter)(s + r * ref_stride + c, 1, x_filter), 7 - ib), ib),
            0, max_sample_val);
  } else {
    int16_t intermediate[8 * (128 + 7)];
    const PIXEL *s = src - 3 - 3 * ref_stride;
    for (int cg = 0; cg < width; cg += 8) {
      int cend = cg + 8 < width ? cg + 8 : width;
      for (int r = 0; r < height + 7; r++)
        for (int c = cg; c < cend; c++)
          intermediate[8 * r + (c - cg)] =
              (int16_t)round_shift_i32(SFX(run_filter)(s + r * ref_stride + c, 1, x_filter), 7 - ib);
      for (int r = 0; r < height; r++)
        for (int c = cg; c < cend; c++)
          dst[r * dst_stride + c] = (PIXEL)clampi(
              round_shift_i32(SFX(run_filter_i16)(intermediate + 8 * r + c - cg, 8, y_filter), 7 + ib),
              0, max_sample_val);
    }
  }
}

/* mc.rs:360-451 */
static void SFX(prep_8tap)(int16_t *tmp, const PIXEL *src, ptrdiff_t ref_stride, int width,
                           int height, int col_frac, int row_frac, int mode_x, int mode_y,
                           int bit_depth) {
  int32_t y_filter[8], x_filter[8];
  orc_get_filter(mode_y, row_frac, height, y_filter);
  orc_get_filter(mode_x, col_frac, width, x_filter);
  const int ib = 4 - (bit_depth == 12 ? 2 : 0);
  const int32_t prep_bias = bit_depth == 8 ? 0 : 8192; /* PREP_BIAS mc.rs:357 */
  if (col_frac == 0 && row_frac == 0) {
    for (int r = 0; r < height; r++)
      for (int c = 0; c < width; c++)
        tmp[r * width + c] =
            (int16_t)((int16_t)((int16_t)src[r * ref_stride + c] << ib) - (int16_t)prep_bias);
  } else if (col_frac == 0) {
    const PIXEL *s = src - 3 * ref_stride;
    for (int r = 0; r < height; r++)
      for (int c = 0; c < width; c++)
        tmp[r * width + c] = (int16_t)(
            round_shift_i32(SFX(run_filter)(s + r * ref_stride + c, ref_stride, y_filter), 7 - ib) - prep_bias);
  } else if (row_frac == 0) {
    const PIXEL *s = src - 3;
    for (int r = 0; r < height; r++)
      for (int c = 0; c < width; c++)
        tmp[r * width + c] = (int16_t)(
            round_shift_i32(SFX(run_filter)(s + r * ref_stride + c, 1, x_filter), 7 - ib) - prep_bias);
  } else {
    int16_t intermediate[8 * (128 + 7)];
    const PIXEL *s = src - 3 - 3 * ref_stride;
    for (int cg = 0; cg < width; cg += 8) {
      int cend = cg + 8 < width ? cg + 8 : width;
      for (int r = 0; r < height + 7; r++)
        for (int c = cg; c < cend; c++)
          intermediate[8 * r + (c - cg)] =
              (int16_t)round_shift_i32(SFX(run_filter)(s + r * ref_stride + c, 1, x_filter), 7 - ib);
      for (int r = 0; r < height; r++)
        for (int c = cg; c < cend; c++)
          tmp[r * width + c] = (int16_t)(
              round_shift_i32(SFX(run_filter_i16)(intermediate + 8 * r + c - cg, 8, y_filter), 7) - prep_bias);
    }
  }
}

/* mc.rs:454-479 */
static void SFX(mc_avg)(PIXEL *dst, ptrdiff_t dst_stride, const int16_t *tmp1, const int16_t *tmp2,
                        int width, int height, int bit_depth) {
  const int32_t max_sample_val = (1 << bit_depth) - 1;
  const int ib = 4 - (bit_depth == 12 ? 2 : 0);
  const int32_t prep_bias = bit_depth == 8 ? 0 : 8192 * 2;
  for (int r = 0; r < height; r++)
    for (int c = 0; c < width; c++)
      dst[r * dst_stride + c] = (PIXEL)clampi(
          round_shift_i32((int32_t)tmp1[r * width + c] + (int32_t)tmp2[r * width + c] + prep_bias, ib + 1),
          0, max_sample_val);
}
