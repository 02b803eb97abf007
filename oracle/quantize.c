/*
 * oracle/quantize.c — restatement of rav1e's quantize -> dequantize -> transform-domain distortion
 * chain, the steps of encode_tx_block that follow the forward transform (src/encoder.rs:1556-1655):
 *   get_log_tx_scale            src/quantize/mod.rs:29-34      (KAT: test_tx_log_scale :186-215)
 *   divu_gen / divu_pair        src/quantize/mod.rs:129-157    (KAT: test_divu_pair :173-181)
 *   QuantizationContext::update src/quantize/mod.rs:219-267    (the rounding offsets)
 *   QuantizationContext::quantize :269-361
 *   rust::dequantize            :368-392
 *   scan orders                 src/scan_order.rs (tables regenerated from their rule, below)
 *   av1_get_coded_tx_size       src/transform/mod.rs (64-point dimensions code 32)
 *   raw tx-domain distortion    src/encoder.rs:1611-1640
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The quantizer step sizes (dc_q / ac_q of a qindex, quantize/tables.rs) are inputs: looking them
 * up is the caller's control-plane business.  Scan tables are not copied: rav1e stores the forward
 * transform's output transposed (index = col * H + row), and in that layout `mcol` is the identity,
 * `mrow` walks rows, and `default` walks anti-diagonals — bottom-up for wide blocks, top-down for
 * tall ones, alternating (odd diagonals top-down) for square ones.  tests/test_oracle_quantize.py
 * checks the generated tables against the digests of all 42 reference tables.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

/* quantize/mod.rs:29-34 */
int orc_get_log_tx_scale(int tx_size) {
  const int n = orc_tx_width(tx_size) * orc_tx_height(tx_size);
  return (n > 256) + (n > 1024);
}

/* quantize/mod.rs:129-146 (d: NonZeroU32) */
void orc_divu_gen(uint32_t d, uint32_t out[3]) {
  const uint64_t nbits = 32;
  const uint64_t m = nbits - (uint64_t)__builtin_clz(d) - 1;
  if ((d & (d - 1)) == 0) {
    out[0] = 0xFFFFFFFFu;
    out[1] = 0xFFFFFFFFu;
    out[2] = (uint32_t)m;
  } else {
    const uint64_t t = ((uint64_t)1 << (m + nbits)) / d;
    const uint64_t r = (t * d + d) & (((uint64_t)1 << nbits) - 1);
    if (r <= (uint64_t)1 << m) {
      out[0] = (uint32_t)t + 1;
      out[1] = 0;
    } else {
      out[0] = (uint32_t)t;
      out[1] = (uint32_t)t;
    }
    out[2] = (uint32_t)m;
  }
}

/* quantize/mod.rs:148-157 */
uint32_t orc_divu_pair(uint32_t x, const uint32_t d[3]) {
  return (uint32_t)((((uint64_t)d[0] * x + d[1]) >> 32) >> d[2]);
}

/* coded dimensions: av1_get_coded_tx_size clamps 64 to 32 */
static int coded_w(int tx_size) { return orc_tx_width(tx_size) > 32 ? 32 : orc_tx_width(tx_size); }
static int coded_h(int tx_size) { return orc_tx_height(tx_size) > 32 ? 32 : orc_tx_height(tx_size); }
int orc_coded_tx_area(int tx_size) { return coded_w(tx_size) * coded_h(tx_size); }

/* av1_scan_orders[tx_size][tx_type] (scan_order.rs:949-1321): types 0..9 default, then
 * V_DCT/V_ADST/V_FLIPADST (10, 12, 14) mrow and H_* (11, 13, 15) mcol. */
int orc_scan_kind(int tx_type) { return tx_type < 10 ? 0 : ((tx_type & 1) ? 2 : 1); }

void orc_scan_order(int tx_size, int tx_type, uint16_t *scan, uint16_t *iscan) {
  const int W = coded_w(tx_size), H = coded_h(tx_size), kind = orc_scan_kind(tx_type);
  int n = 0;
  if (kind == 2) {
    for (int i = 0; i < W * H; i++) scan[n++] = (uint16_t)i;
  } else if (kind == 1) {
    for (int r = 0; r < H; r++)
      for (int c = 0; c < W; c++) scan[n++] = (uint16_t)(c * H + r);
  } else {
    for (int d = 0; d < W + H - 1; d++) {
      const int r_lo = d - (W - 1) > 0 ? d - (W - 1) : 0, r_hi = d < H - 1 ? d : H - 1;
      const int down = W < H || (W == H && (d & 1));  /* rows increasing along the diagonal */
      for (int k = 0; k <= r_hi - r_lo; k++) {
        const int r = down ? r_lo + k : r_hi - k;
        scan[n++] = (uint16_t)((d - r) * H + r);
      }
    }
  }
  if (iscan)
    for (int i = 0; i < n; i++) iscan[scan[i]] = (uint16_t)i;
}

typedef struct {
  int log_tx_scale;
  uint32_t dc_quant, ac_quant;
  uint32_t dc_offset, ac_offset0, ac_offset1, ac_offset_eob;
  uint32_t dc_mul_add[3], ac_mul_add[3];
} qctx;

/* quantize/mod.rs:219-267 with the step sizes given */
static void qctx_update(qctx *q, int tx_size, int is_intra, uint32_t dc_quant, uint32_t ac_quant) {
  q->log_tx_scale = orc_get_log_tx_scale(tx_size);
  q->dc_quant = dc_quant;
  q->ac_quant = ac_quant;
  orc_divu_gen(dc_quant, q->dc_mul_add);
  orc_divu_gen(ac_quant, q->ac_mul_add);
  q->dc_offset = dc_quant * (is_intra ? 109u : 108u) / 256;
  q->ac_offset0 = ac_quant * (is_intra ? 98u : 97u) / 256;
  q->ac_offset1 = ac_quant * (is_intra ? 109u : 108u) / 256;
  q->ac_offset_eob = ac_quant * (is_intra ? 88u : 44u) / 256;
}

static inline int32_t copysign_i32(uint32_t v, int32_t s) { return s < 0 ? -(int32_t)v : (int32_t)v; }

#define DEF_CHAIN(SFX, T)                                                                          \
  /* quantize/mod.rs:269-361; qcoeffs must be zero-filled (coded area entries) */                  \
  static uint16_t quantize_##SFX(const qctx *q, const T *coeffs, T *qcoeffs, const uint16_t *scan, \
                                 const uint16_t *iscan, int coded) {                               \
    {                                                                                              \
      const int32_t coeff = (int32_t)((uint32_t)(int32_t)coeffs[0] << q->log_tx_scale);            \
      const uint32_t abs_coeff = coeff < 0 ? 0u - (uint32_t)coeff : (uint32_t)coeff;               \
      qcoeffs[0] = (T)copysign_i32(orc_divu_pair(abs_coeff + q->dc_offset, q->dc_mul_add), coeff); \
    }                                                                                              \
    const size_t dzv = (size_t)q->ac_quant - (size_t)q->ac_offset_eob;                             \
    const T deadzone = (T)((dzv + ((size_t)1 << q->log_tx_scale) - 1) >> q->log_tx_scale);         \
    uint16_t eob_minus_one = 0;                                                                    \
    for (int i = 0; i < coded; i++) {                                                              \
      const T c = coeffs[i];                                                                       \
      const T a = (T)(c < 0 ? -c : c);                                                             \
      if (a >= deadzone && iscan[i] > eob_minus_one) eob_minus_one = iscan[i];                     \
    }                                                                                              \
    const uint16_t eob = eob_minus_one > 0 ? (uint16_t)(eob_minus_one + 1) : (uint16_t)(qcoeffs[0] != 0); \
    uint32_t level_mode = 1;                                                                       \
    for (int j = 1; j < (int)eob; j++) {                                                           \
      const int pos = scan[j];                                                                     \
      const int32_t coeff = (int32_t)((uint32_t)(int32_t)coeffs[pos] << q->log_tx_scale);          \
      const uint32_t abs_coeff = coeff < 0 ? 0u - (uint32_t)coeff : (uint32_t)coeff;               \
      const uint32_t level0 = orc_divu_pair(abs_coeff, q->ac_mul_add);                             \
      const uint32_t offset = level0 > 1 - level_mode ? q->ac_offset1 : q->ac_offset0;             \
      const uint32_t abs_q = level0 + (abs_coeff + offset >= (level0 + 1) * q->ac_quant);          \
      if (level_mode != 0 && abs_q == 0)                                                           \
        level_mode = 0;                                                                            \
      else if (abs_q > 1)                                                                          \
        level_mode = 1;                                                                            \
      qcoeffs[pos] = (T)copysign_i32(abs_q, coeff);                                                \
    }                                                                                              \
    return eob;                                                                                    \
  }                                                                                                \
  /* quantize/mod.rs:368-392 */                                                                    \
  static void dequantize_##SFX(const qctx *q, const T *qcoeffs, T *rcoeffs, int coded) {           \
    const int32_t offset = (1 << q->log_tx_scale) - 1;                                             \
    for (int i = 0; i < coded; i++) {                                                              \
      const int32_t c = (int32_t)qcoeffs[i];                                                       \
      const int32_t quant = (int32_t)(i == 0 ? q->dc_quant : q->ac_quant);                         \
      rcoeffs[i] = (T)((int32_t)((uint32_t)c * (uint32_t)quant + (uint32_t)((c >> 31) & offset)) >> q->log_tx_scale); \
    }                                                                                              \
  }                                                                                                \
  /* encoder.rs:1611-1640 (before estimate_rate / the bias multiplications) */                     \
  static uint64_t tx_dist_##SFX(const qctx *q, const T *coeffs, const T *rcoeffs, int area, int coded) { \
    uint64_t raw = 0;                                                                              \
    for (int i = 0; i < coded; i++) {                                                              \
      const int32_t c = (int32_t)coeffs[i] - (int32_t)rcoeffs[i];                                  \
      raw += (uint64_t)(int64_t)(int32_t)((uint32_t)c * (uint32_t)c);                              \
    }                                                                                              \
    for (int i = coded; i < area; i++) {                                                           \
      const int32_t c = (int32_t)coeffs[i];                                                        \
      raw += (uint64_t)(int64_t)(int32_t)((uint32_t)c * (uint32_t)c);                              \
    }                                                                                              \
    const int bits = 2 * (3 - q->log_tx_scale);                                                    \
    return (raw + ((uint64_t)1 << (bits - 1))) >> bits;                                            \
  }
DEF_CHAIN(i16, int16_t)
DEF_CHAIN(i32, int32_t)

/* The whole chain for nblocks blocks of one (tx_size, tx_type): coeffs[n][w*h] in,
 * qcoeffs[n][coded] / rcoeffs[n][coded] (may be NULL) / eob[n] / tx_dist[n] (may be NULL) out. */
void orc_quantize_chain_batch(const void *coeffs, size_t nblocks, int tx_size, int tx_type,
                              uint32_t dc_quant, uint32_t ac_quant, int is_intra, int coeff_is_i32,
                              void *qcoeffs, void *rcoeffs, uint16_t *eob, uint64_t *tx_dist,
                              int threads) {
  const int area = orc_tx_width(tx_size) * orc_tx_height(tx_size), coded = orc_coded_tx_area(tx_size);
  uint16_t *scan = (uint16_t *)malloc(2 * (size_t)coded * sizeof(uint16_t)), *iscan = scan + coded;
  orc_scan_order(tx_size, tx_type, scan, iscan);
  qctx q;
  qctx_update(&q, tx_size, is_intra, dc_quant, ac_quant);
  const size_t esz = coeff_is_i32 ? 4 : 2;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : orc_num_threads())
  for (ptrdiff_t i = 0; i < (ptrdiff_t)nblocks; i++) {
    const uint8_t *c = (const uint8_t *)coeffs + (size_t)i * area * esz;
    uint8_t *qc = (uint8_t *)qcoeffs + (size_t)i * coded * esz;
    int32_t rtmp[1024];
    uint8_t *rc = rcoeffs ? (uint8_t *)rcoeffs + (size_t)i * coded * esz : (uint8_t *)rtmp;
    memset(qc, 0, (size_t)coded * esz);
    if (coeff_is_i32) {
      const uint16_t e = quantize_i32(&q, (const int32_t *)c, (int32_t *)qc, scan, iscan, coded);
      if (eob) eob[i] = e;
      dequantize_i32(&q, (const int32_t *)qc, (int32_t *)rc, coded);
      if (tx_dist) tx_dist[i] = tx_dist_i32(&q, (const int32_t *)c, (const int32_t *)rc, area, coded);
    } else {
      const uint16_t e = quantize_i16(&q, (const int16_t *)c, (int16_t *)qc, scan, iscan, coded);
      if (eob) eob[i] = e;
      dequantize_i16(&q, (const int16_t *)qc, (int16_t *)rc, coded);
      if (tx_dist) tx_dist[i] = tx_dist_i16(&q, (const int16_t *)c, (const int16_t *)rc, area, coded);
    }
  }
  free(scan);
}
