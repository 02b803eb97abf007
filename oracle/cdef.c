/*
 * oracle/cdef.c — restatement of rav1e src/cdef.rs: first_max_element (:54-76), cdef_find_dir
 * (:84-143), constrain (:146-159), pad_into_tmp16 (:161-194), cdef_filter_block (:198-298),
 * adjust_strength (:315-322) and the per-superblock driver cdef_analyze_superblock /
 * cdef_filter_superblock / cdef_filter_tile (:340-374, :401-570, :597-625) flattened to a
 * whole-frame pass (tile rect = frame).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Pinning: first_max_element is pinned by the reference KAT (:304-309, tests/test_oracle_cdef.py).
 * The direction search and the filter have no stored vectors upstream ("parity unpinned"); they
 * are AV1-normative (spec 7.15) and cross-checked on constructed inputs: pure directional
 * ramps give the expected direction index, a flat block is a fixed point of the filter, zero
 * strengths copy, and the sentinel (CDEF_VERY_LARGE) never leaks into the output.
 */
#include "oracle.h"

#include <string.h>

#define CDEF_VERY_LARGE 0x8000
#define CDEF_HAVE_LEFT 1
#define CDEF_HAVE_RIGHT 2
#define CDEF_HAVE_TOP 4
#define CDEF_HAVE_BOTTOM 8
#define CDEF_HAVE_ALL 15

static const int32_t CDEF_DIV_TABLE[9] = {0, 840, 420, 280, 210, 168, 140, 120, 105}; /* :54 */

static inline int msb32(int32_t x) { return 31 - __builtin_clz((uint32_t)x); }

/* cdef.rs:66-76: first instance of the maximum */
int orc_first_max_element(const int32_t *elems, int n, int32_t *max_out) {
  int best = 0;
  for (int i = 1; i < n; i++)
    if (elems[i] > elems[best]) best = i;
  if (max_out) *max_out = elems[best];
  return best;
}

static inline int32_t px_at(const void *img, ptrdiff_t stride, int bpp, int y, int x) {
  return bpp == 1 ? (int32_t)((const uint8_t *)img)[y * stride + x]
                  : (int32_t)((const uint16_t *)img)[y * stride + x];
}

/* cdef.rs:84-143 */
int orc_cdef_find_dir(const void *img, ptrdiff_t stride, int bpp, uint32_t *var, int coeff_shift) {
  int32_t cost[8] = {0};
  int32_t partial[8][15];
  memset(partial, 0, sizeof partial);
  for (int i = 0; i < 8; i++) {
    for (int j = 0; j < 8; j++) {
      int32_t p = px_at(img, stride, bpp, i, j);
      int32_t x = (p >> coeff_shift) - 128;
      partial[0][i + j] += x;
      partial[1][i + j / 2] += x;
      partial[2][i] += x;
      partial[3][3 + i - j / 2] += x;
      partial[4][7 + i - j] += x;
      partial[5][3 - i / 2 + j] += x;
      partial[6][j] += x;
      partial[7][i / 2 + j] += x;
    }
  }
  for (int i = 0; i < 8; i++) {
    cost[2] += partial[2][i] * partial[2][i];
    cost[6] += partial[6][i] * partial[6][i];
  }
  cost[2] *= CDEF_DIV_TABLE[8];
  cost[6] *= CDEF_DIV_TABLE[8];
  for (int i = 0; i < 7; i++) {
    cost[0] += (partial[0][i] * partial[0][i] + partial[0][14 - i] * partial[0][14 - i]) *
               CDEF_DIV_TABLE[i + 1];
    cost[4] += (partial[4][i] * partial[4][i] + partial[4][14 - i] * partial[4][14 - i]) *
               CDEF_DIV_TABLE[i + 1];
  }
  cost[0] += partial[0][7] * partial[0][7] * CDEF_DIV_TABLE[8];
  cost[4] += partial[4][7] * partial[4][7] * CDEF_DIV_TABLE[8];
  for (int i = 1; i < 8; i += 2) {
    for (int j = 0; j < 5; j++) cost[i] += partial[i][3 + j] * partial[i][3 + j];
    cost[i] *= CDEF_DIV_TABLE[8];
    for (int j = 0; j < 3; j++)
      cost[i] += (partial[i][j] * partial[i][j] + partial[i][10 - j] * partial[i][10 - j]) *
                 CDEF_DIV_TABLE[2 * j + 2];
  }
  int32_t best_cost;
  int best_dir = orc_first_max_element(cost, 8, &best_cost);
  *var = (uint32_t)((best_cost - cost[(best_dir + 4) & 7]) >> 10);
  return best_dir;
}

/* cdef.rs:146-159 */
static inline int32_t constrain(int32_t diff, int32_t threshold, int32_t damping) {
  if (threshold == 0) return 0;
  int32_t shift = damping - msb32(threshold);
  if (shift < 0) shift = 0;
  int32_t ad = diff < 0 ? -diff : diff;
  int32_t mag = threshold - (ad >> shift);
  if (mag < 0) mag = 0;
  if (mag > ad) mag = ad; /* clamp(0, |diff|) */
  return diff < 0 ? -mag : mag;
}

/* cdef.rs:198-298 with edges == CDEF_HAVE_ALL: `in` is a u16 image that already carries the
 * sentinel where pixels are unavailable. */
static void filter_block_all(void *dst, ptrdiff_t dst_stride, int dst_bpp, const uint16_t *in,
                             ptrdiff_t istride, int pri_strength, int sec_strength, int dir,
                             int damping, int bit_depth, int xdec, int ydec) {
  const int xsize = 8 >> xdec, ysize = 8 >> ydec;
  const int coeff_shift = bit_depth - 8;
  static const int pri_taps_tab[2][2] = {{4, 2}, {3, 3}};
  static const int sec_taps_tab[2][2] = {{2, 1}, {2, 1}};
  const int *pri_taps = pri_taps_tab[(pri_strength >> coeff_shift) & 1];
  const int *sec_taps = sec_taps_tab[(pri_strength >> coeff_shift) & 1];
  const ptrdiff_t dirs[8][2] = {
      {-1 * istride + 1, -2 * istride + 2}, {0 * istride + 1, -1 * istride + 2},
      {0 * istride + 1, 0 * istride + 2},   {0 * istride + 1, 1 * istride + 2},
      {1 * istride + 1, 2 * istride + 2},   {1 * istride + 0, 2 * istride + 1},
      {1 * istride + 0, 2 * istride + 0},   {1 * istride + 0, 2 * istride - 1}};
  for (int i = 0; i < ysize; i++) {
    for (int j = 0; j < xsize; j++) {
      const uint16_t *ptr_in = in + i * istride + j;
      int32_t x = *ptr_in, sum = 0, max = x, min = x;
      for (int k = 0; k < 2; k++) {
        ptrdiff_t d0 = dirs[dir][k], d1 = dirs[(dir + 2) & 7][k], d2 = dirs[(dir + 6) & 7][k];
        int32_t p[2] = {ptr_in[d0], ptr_in[-d0]};
        for (int t = 0; t < 2; t++) {
          sum += pri_taps[k] * constrain(p[t] - x, pri_strength, damping);
          if (p[t] != CDEF_VERY_LARGE && p[t] > max) max = p[t];
          if (p[t] < min) min = p[t];
        }
        int32_t s[4] = {ptr_in[d1], ptr_in[-d1], ptr_in[d2], ptr_in[-d2]};
        for (int t = 0; t < 4; t++) {
          if (s[t] != CDEF_VERY_LARGE && s[t] > max) max = s[t];
          if (s[t] < min) min = s[t];
          sum += sec_taps[k] * constrain(s[t] - x, sec_strength, damping);
        }
      }
      int32_t v = x + ((8 + sum - (sum < 0)) >> 4);
      v = v < min ? min : v > max ? max : v;
      if (dst_bpp == 1)
        ((uint8_t *)dst)[i * dst_stride + j] = (uint8_t)v;
      else
        ((uint16_t *)dst)[i * dst_stride + j] = (uint16_t)v;
    }
  }
}

/* cdef.rs:161-194 + :205-231: build the 12x12 u16 tile (sentinel where edges are missing),
 * then filter.  `in` points at the block's top-left pixel in an image of `in_bpp` bytes/px. */
void orc_cdef_filter_block_px(void *dst, ptrdiff_t dst_stride, const void *in, ptrdiff_t istride,
                              int bpp, int pri_strength, int sec_strength, int dir, int damping,
                              int bit_depth, int xdec, int ydec, int edges) {
  const int bw = 8 >> xdec, bh = 8 >> ydec;
  const int tmpstride = 2 + bw + 2;
  uint16_t tmp[12 * 12];
  for (int i = 0; i < 144; i++) tmp[i] = CDEF_VERY_LARGE;
  int x0 = (edges & CDEF_HAVE_LEFT) ? -2 : 0, x1 = bw + ((edges & CDEF_HAVE_RIGHT) ? 2 : 0);
  int y0 = (edges & CDEF_HAVE_TOP) ? -2 : 0, y1 = bh + ((edges & CDEF_HAVE_BOTTOM) ? 2 : 0);
  for (int y = y0; y < y1; y++)
    for (int x = x0; x < x1; x++)
      tmp[(y + 2) * tmpstride + (x + 2)] =
          (uint16_t)(bpp == 1 ? ((const uint8_t *)in)[y * istride + x]
                              : ((const uint16_t *)in)[y * istride + x]);
  filter_block_all(dst, dst_stride, bpp, tmp + 2 * tmpstride + 2, tmpstride, pri_strength,
                   sec_strength, dir, damping, bit_depth, xdec, ydec);
}

/* The asm-facing form (asm/x86/cdef.rs:16-37): caller supplies the padded u16 tile. */
void orc_cdef_filter_block(void *dst, ptrdiff_t dst_stride, int dst_bpp, const uint16_t *in,
                           ptrdiff_t in_stride, int pri_strength, int sec_strength, int dir,
                           int damping, int bit_depth, int xdec, int ydec, int edges) {
  (void)edges;
  filter_block_all(dst, dst_stride, dst_bpp, in, in_stride, pri_strength, sec_strength, dir,
                   damping, bit_depth, xdec, ydec);
}

/* cdef.rs:315-322 */
int orc_cdef_adjust_strength(int strength, int var) {
  int i = (var >> 6) != 0 ? (msb32(var >> 6) < 12 ? msb32(var >> 6) : 12) : 0;
  return var != 0 ? (strength * (4 + i) + 8) >> 4 : 0;
}

/* cdef_analyze_superblock (:340-374) over the whole luma plane: dir/var per 8x8 block, 0 where
 * skipped.  skip8: one byte per 8x8 block (the AND of its four 4x4 `skip` flags), row-major,
 * stride w8; NULL = nothing skipped. */
void orc_cdef_analyze_frame(const void *luma, ptrdiff_t stride, int bpp, int width, int height,
                            int bit_depth, const uint8_t *skip8, uint8_t *dir, int32_t *var) {
  const int w8 = width >> 3, h8 = height >> 3;
#pragma omp parallel for schedule(static)
  for (int by = 0; by < h8; by++)
    for (int bx = 0; bx < w8; bx++) {
      dir[by * w8 + bx] = 0;
      var[by * w8 + bx] = 0;
      if (skip8 && skip8[by * w8 + bx]) continue;
      uint32_t v = 0;
      const uint8_t *p = (const uint8_t *)luma + ((ptrdiff_t)(8 * by) * stride + 8 * bx) * bpp;
      dir[by * w8 + bx] = (uint8_t)orc_cdef_find_dir(p, stride, bpp, &v, bit_depth - 8);
      var[by * w8 + bx] = (int32_t)v;
    }
}

/* cdef_filter_superblock (:401-570) flattened over a frame plane (tile rect = frame; width and
 * height are the LUMA dimensions, multiples of 8).  plane 0 = luma, >0 = chroma with (xdec, ydec).
 * strength_sb: the 6-bit cdef strength (pri*4 + sec, fi.cdef_y_strengths / cdef_uv_strengths
 * [cdef_index]) per 64x64 superblock, row-major, stride ceil(width/64). */
void orc_cdef_filter_plane(const void *in, ptrdiff_t in_stride, void *out, ptrdiff_t out_stride,
                           int bpp, int plane, int xdec, int ydec, int width, int height,
                           int bit_depth, int damping, const uint8_t *skip8, const uint8_t *dir,
                           const int32_t *var, const uint8_t *strength_sb) {
  const int w8 = width >> 3, h8 = height >> 3, sbw = (width + 63) >> 6;
  const int coeff_shift = bit_depth - 8;
  const int xsize = 8 >> xdec, ysize = 8 >> ydec;
  static const uint8_t uv_dir_422[8] = {7, 0, 2, 4, 5, 6, 6, 6}; /* :505-509 */
#pragma omp parallel for schedule(static)
  for (int gy = 0; gy < h8; gy++)
    for (int gx = 0; gx < w8; gx++) {
      int edges = 0;
      if (gy > 0) edges |= CDEF_HAVE_TOP;
      if (gx > 0) edges |= CDEF_HAVE_LEFT;
      if (gy + 1 < h8) edges |= CDEF_HAVE_BOTTOM;
      if (gx + 1 < w8) edges |= CDEF_HAVE_RIGHT;
      const uint8_t *pin = (const uint8_t *)in + ((ptrdiff_t)(gy * ysize) * in_stride + gx * xsize) * bpp;
      uint8_t *pout = (uint8_t *)out + ((ptrdiff_t)(gy * ysize) * out_stride + gx * xsize) * bpp;
      if (skip8 && skip8[gy * w8 + gx]) { /* :557-564 copy */
        for (int i = 0; i < ysize; i++) memcpy(pout + (ptrdiff_t)i * out_stride * bpp, pin + (ptrdiff_t)i * in_stride * bpp, (size_t)xsize * bpp);
        continue;
      }
      const int strength = strength_sb[(gy >> 3) * sbw + (gx >> 3)];
      const int pri = strength / 4;
      int sec = strength % 4;
      if (sec == 3) sec += 1; /* :421-426 */
      int local_pri, local_sec = sec << coeff_shift, local_damping = damping + coeff_shift, local_dir;
      const int d = dir[gy * w8 + gx];
      if (plane == 0) {
        local_pri = orc_cdef_adjust_strength(pri << coeff_shift, var[gy * w8 + gx]);
        local_dir = pri != 0 ? d : 0;
      } else {
        local_pri = pri << coeff_shift;
        local_damping -= 1;
        local_dir = pri != 0 ? (xdec != ydec ? uv_dir_422[d] : d) : 0;
      }
      orc_cdef_filter_block_px(pout, out_stride, pin, in_stride, bpp, local_pri, local_sec,
                               local_dir, local_damping, bit_depth, xdec, ydec, edges);
    }
}
