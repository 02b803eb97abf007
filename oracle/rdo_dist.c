/*
 * oracle/rdo_dist.c — restatement of rav1e's RDO distortion kernels:
 *   get_weighted_sse   src/dist.rs:234-283   (GET_WEIGHTED_SSE_SHIFT :223; DistortionScale::new
 *                                              src/rdo.rs:579-583, SHIFT = 14 :571)
 *   cdef_dist_kernel   src/dist.rs:302-372   (AREA_DIVISORS :288-297)
 *   apply_ssim_boost   src/activity.rs:159-186, ssim_boost_rsqrt :107-143
 *   DistortionScale::mul_u64  src/rdo.rs:613-615 (RawDistortion * DistortionScale)
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * No stored vector in the reference pins these ("parity unpinned"): the reference tests are
 * asm==rust on random input plus the float cross-checks of activity.rs:204-252, which
 * tests/test_oracle_rdo_dist.py restates.  u32 sums wrap like Rust release builds.
 */
#include "oracle.h"

#define WSSE_SHIFT 8 /* dist.rs:223 */

#define DEF_WSSE(NAME, PIXEL)                                                                     \
  uint64_t NAME(const PIXEL *src1, ptrdiff_t stride1, const PIXEL *src2, ptrdiff_t stride2,       \
                const uint32_t *scale, size_t scale_stride, int w, int h) {                       \
    uint64_t sse = 0;                                                                             \
    /* chunk_size = IMPORTANCE_BLOCK_SIZE >> 1 = 4 (dist.rs:244); vert/horz windows of 4 stepped  \
       by 4 only yield whole chunks */                                                            \
    for (int cy = 0; cy + 4 <= h; cy += 4)                                                        \
      for (int cx = 0; cx + 4 <= w; cx += 4) {                                                    \
        uint32_t sum = 0;                                                                         \
        for (int y = 0; y < 4; y++)                                                               \
          for (int x = 0; x < 4; x++) {                                                           \
            int32_t c = (int32_t)src1[(cy + y) * stride1 + cx + x] -                              \
                        (int32_t)src2[(cy + y) * stride2 + cx + x];                               \
            sum += (uint32_t)(c * c);                                                             \
          }                                                                                       \
        uint64_t s = scale[(size_t)(cy / 4) * scale_stride + (size_t)(cx / 4)];                   \
        sse += ((uint64_t)sum * s + ((1u << WSSE_SHIFT) >> 1)) >> WSSE_SHIFT;                     \
      }                                                                                           \
    /* den = DistortionScale::new(1, 1 << 8).0 = ((1 << 14) + 128) / 256 = 64 */                  \
    const uint64_t den = (((uint64_t)1 << 14) + ((1u << WSSE_SHIFT) / 2)) / (1u << WSSE_SHIFT);   \
    return (sse + (den >> 1)) / den;                                                              \
  }
DEF_WSSE(orc_weighted_sse_u8, uint8_t)
DEF_WSSE(orc_weighted_sse_u16, uint16_t)

/* activity.rs:107-143 */
static void ssim_boost_rsqrt(uint64_t x, uint16_t *norm, uint8_t *shift) {
  const int INSHIFT = 16, OUTSHIFT = 14;
  const int ilog2 = 63 - __builtin_clzll(x);
  const int16_t k = (int16_t)(ilog2 >> 1);
  const int16_t s = (int16_t)(2 * k - (INSHIFT - 2));
  const uint16_t t = (uint16_t)(s > 0 ? x >> s : x << -s);
  *shift = (uint8_t)(OUTSHIFT + ((s + INSHIFT) >> 1));
  const int32_t n = (int32_t)t - 32768;
  const int32_t inner = -13490 + ((n * 6711) >> 15);
  const int32_t rsqrt = 23557 + ((n * inner) >> 15);
  *norm = (uint16_t)rsqrt;
}

/* activity.rs:159-186 */
uint32_t orc_apply_ssim_boost(uint32_t input, uint32_t svar, uint32_t dvar, int bit_depth) {
  const int coeff_shift = bit_depth - 8;
  const uint64_t sv = svar >> (2 * coeff_shift), dv = dvar >> (2 * coeff_shift);
  const uint64_t C1 = 3355, C2 = 16128, C3 = 12338;
  const int RATIO_SHIFT = 14;
  const uint64_t RATIO = (((C1 << (RATIO_SHIFT + 1)) / C3) + 1) >> 1;
  uint16_t norm;
  uint8_t shift;
  ssim_boost_rsqrt(C1 * C1 + sv * dv, &norm, &shift);
  return (uint32_t)(((uint64_t)input * (((RATIO * (sv + dv + C2)) * (uint64_t)norm) >> RATIO_SHIFT)) >> shift);
}

static const uint16_t AREA_DIVISORS[64] = { /* round(2^14 / (1 + x)), dist.rs:288-297 */
    16384, 8192, 5461, 4096, 3277, 2731, 2341, 2048, 1820, 1638, 1489, 1365, 1260, 1170, 1092, 1024,
    964,   910,  862,  819,  780,  745,  712,  683,  655,  630,  607,  585,  565,  546,  529,  512,
    496,   482,  468,  455,  443,  431,  420,  410,  400,  390,  381,  372,  364,  356,  349,  341,
    334,   328,  321,  315,  309,  303,  298,  293,  287,  282,  278,  273,  269,  264,  260,  256};

static inline uint32_t sat_sub_u32(uint32_t a, uint32_t b) { return a > b ? a - b : 0; }

#define DEF_CDEF_DIST(NAME, PIXEL)                                                                \
  uint32_t NAME(const PIXEL *src, ptrdiff_t ss, const PIXEL *dst, ptrdiff_t ds, int w, int h,     \
                int bit_depth, uint32_t raw[3]) {                                                 \
    uint32_t sum_s = 0, sum_d = 0, sum_s2 = 0, sum_d2 = 0, sum_sd = 0;                            \
    for (int y = 0; y < h; y++)                                                                   \
      for (int x = 0; x < w; x++) {                                                               \
        uint32_t s = src[y * ss + x], d = dst[y * ds + x];                                        \
        sum_s += s;                                                                               \
        sum_d += d;                                                                               \
        sum_s2 += s * s;                                                                          \
        sum_d2 += d * d;                                                                          \
        sum_sd += s * d;                                                                          \
      }                                                                                           \
    const uint32_t sse = sum_d2 + sum_s2 - 2 * sum_sd;                                            \
    const uint64_t S = sum_s, D = sum_d, div = AREA_DIVISORS[w * h - 1];                          \
    const int div_shift = 14;                                                                     \
    uint32_t svar = sat_sub_u32(sum_s2, (uint32_t)((S * S * div + ((1u << div_shift) >> 1)) >> div_shift)); \
    uint32_t dvar = sat_sub_u32(sum_d2, (uint32_t)((D * D * div + ((1u << div_shift) >> 1)) >> div_shift)); \
    const int scale_shift = 14 - 6;                                                               \
    svar = (uint32_t)(((uint64_t)svar * div + ((1u << scale_shift) >> 1)) >> scale_shift);        \
    dvar = (uint32_t)(((uint64_t)dvar * div + ((1u << scale_shift) >> 1)) >> scale_shift);        \
    if (raw) {                                                                                    \
      raw[0] = svar;                                                                              \
      raw[1] = dvar;                                                                              \
      raw[2] = sse;                                                                               \
    }                                                                                             \
    return orc_apply_ssim_boost(sse, svar, dvar, bit_depth);                                      \
  }
DEF_CDEF_DIST(orc_cdef_dist_kernel_u8, uint8_t)
DEF_CDEF_DIST(orc_cdef_dist_kernel_u16, uint16_t)

/* rdo.rs:613-615 */
uint64_t orc_distortion_scale_mul(uint32_t scale, uint64_t dist) {
  return ((uint64_t)scale * dist + ((1u << 14) >> 1)) >> 14;
}

/* ------------------------------------------------------------------------------------------
 * ActivityMask (src/activity.rs:21-69): variance_8x8 (:71-100) of every 8x8 luma block of the
 * plane rounded up to whole blocks (the region reads into the plane's padding), and fill_scales
 * (:58-68): ssim_boost(var, var, bit_depth) = apply_ssim_boost(1 << 14, ..) (:147-154).
 * ------------------------------------------------------------------------------------------ */
#define DEF_VAR8(NAME, PIXEL)                                                                 \
  uint32_t NAME(const PIXEL *src, ptrdiff_t stride) {                                         \
    uint16_t sum_s_cols[8] = {0};                                                             \
    uint32_t sum_s2_cols[8] = {0};                                                            \
    for (int j = 0; j < 8; j++)                                                               \
      for (int i = 0; i < 8; i++) {                                                           \
        const uint16_t s = (uint16_t)src[j * stride + i];                                     \
        sum_s_cols[i] = (uint16_t)(sum_s_cols[i] + s);                                        \
        sum_s2_cols[i] += (uint32_t)s * (uint32_t)s;                                          \
      }                                                                                       \
    uint64_t sum_s = 0, sum_s2 = 0;                                                           \
    for (int i = 0; i < 8; i++) {                                                             \
      sum_s += sum_s_cols[i];                                                                 \
      sum_s2 += sum_s2_cols[i];                                                               \
    }                                                                                         \
    const uint64_t v = sum_s2 - ((sum_s * sum_s + 32) >> 6);                                  \
    return v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v; /* u32::try_from(..).unwrap_or(MAX) */ \
  }
DEF_VAR8(orc_variance_8x8_u8, uint8_t)
DEF_VAR8(orc_variance_8x8_u16, uint16_t)

/* luma: pixel (0,0), readable up to the next multiple of 8 in both directions.
 * variances / scales: ceil(w/8) * ceil(h/8) entries, row-major; scales may be NULL. */
void orc_activity_mask(const void *luma, ptrdiff_t stride, int bpp, int width, int height, int bit_depth,
                       uint32_t *variances, uint32_t *scales) {
  const int wb = (width + 7) >> 3, hb = (height + 7) >> 3;
  for (int y = 0; y < hb; y++)
    for (int x = 0; x < wb; x++) {
      const uint32_t v = bpp == 1 ? orc_variance_8x8_u8((const uint8_t *)luma + (ptrdiff_t)(y * 8) * stride + x * 8, stride)
                                  : orc_variance_8x8_u16((const uint16_t *)luma + (ptrdiff_t)(y * 8) * stride + x * 8, stride);
      variances[y * wb + x] = v;
      if (scales) scales[y * wb + x] = orc_apply_ssim_boost(1u << 14, v, v, bit_depth);
    }
}
