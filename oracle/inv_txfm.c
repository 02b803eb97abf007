/*
 * oracle/inv_txfm.c — restatement of rav1e's inverse transform + reconstruction
 * (src/transform/inverse.rs): rust::inverse_transform_add :1637-1704, INV_TXFM_FNS :1593-1623,
 * INV_INTERMEDIATE_SHIFTS :1710-1711, av1_iwht4 :35-53, the identity transforms :150-157,
 * :299-304, :579-584, :886-891, the flipped ADSTs :93-96, :209-212, :403-407, and the primitives
 * half_btf / clamp_value (src/transform/mod.rs:296-315).  The 1-D DCT / ADST butterfly networks
 * are generated mechanically into inv_txfm_networks.h by tools/gen_inv_txfm.py.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Pinned jointly with the forward transform by the reference's own round-trip test
 * (transform/mod.rs:479-617: forward_transform -> inverse_transform_add must reproduce the source
 * within a per-(size, type) tolerance of 0..2), restated in tests/test_oracle_inv_txfm.py.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

/* i32 arithmetic wraps in Rust release builds */
#define WADD(a, b) ((int32_t)((uint32_t)(a) + (uint32_t)(b)))
#define WSUB(a, b) ((int32_t)((uint32_t)(a) - (uint32_t)(b)))
#define WMUL(a, b) ((int32_t)((uint32_t)(a) * (uint32_t)(b)))
#define WNEG(a) ((int32_t)(0u - (uint32_t)(a)))
#define INV_COS_BIT 12

/* transform/mod.rs:296-307 */
static inline int32_t half_btf(int32_t w0, int32_t in0, int32_t w1, int32_t in1, int bit) {
  const int32_t result = WADD(WMUL(w0, in0), WMUL(w1, in1));
  if (bit == 0) return result;
  return WADD(result, 1 << (bit - 1)) >> bit;
}
#define HALF_BTF(w0, in0, w1, in1) half_btf((w0), (in0), (w1), (in1), INV_COS_BIT)

/* transform/mod.rs:310-315 */
static inline int32_t clamp_value(int32_t value, int bit) {
  const int32_t max_value = (int32_t)(((int64_t)1 << (bit - 1)) - 1);
  const int32_t min_value = (int32_t)(-((int64_t)1 << (bit - 1)));
  return value < min_value ? min_value : value > max_value ? max_value : value;
}
#define CLAMP_VALUE(v, range) clamp_value((v), (range))

/* v_frame round_shift: (v + (1 << b >> 1)) >> b */
static inline int32_t round_shift_i32(int32_t v, int b) { return WADD(v, (1 << b) >> 1) >> b; }
#define ROUND_SHIFT(v, b) round_shift_i32((v), (b))

#include "inv_txfm_networks.h"

#define SQRT2_BITS 12 /* transform/mod.rs:47-49 */
#define SQRT2 5793
#define INV_SQRT2 2896

static void reverse_n(int32_t *v, int n) {
  for (int i = 0; i < n / 2; i++) {
    const int32_t t = v[i];
    v[i] = v[n - 1 - i];
    v[n - 1 - i] = t;
  }
}

/* inverse.rs:93-96, :209-212, :403-407 */
static void av1_iflipadst4(const int32_t *in, int32_t *out, int range) {
  av1_iadst4(in, out, range);
  reverse_n(out, 4);
}
static void av1_iflipadst8(const int32_t *in, int32_t *out, int range) {
  av1_iadst8(in, out, range);
  reverse_n(out, 8);
}
static void av1_iflipadst16(const int32_t *in, int32_t *out, int range) {
  av1_iadst16(in, out, range);
  reverse_n(out, 16);
}
/* inverse.rs:150-157, :299-304, :579-584, :886-891 */
static void av1_iidentity4(const int32_t *in, int32_t *out, int range) {
  (void)range;
  for (int i = 0; i < 4; i++) out[i] = round_shift_i32(WMUL(SQRT2, in[i]), 12);
}
static void av1_iidentity8(const int32_t *in, int32_t *out, int range) {
  (void)range;
  for (int i = 0; i < 8; i++) out[i] = WMUL(2, in[i]);
}
static void av1_iidentity16(const int32_t *in, int32_t *out, int range) {
  (void)range;
  for (int i = 0; i < 16; i++) out[i] = round_shift_i32(WMUL(WMUL(SQRT2, 2), in[i]), 12);
}
static void av1_iidentity32(const int32_t *in, int32_t *out, int range) {
  (void)range;
  for (int i = 0; i < 32; i++) out[i] = WMUL(4, in[i]);
}
/* inverse.rs:35-53 */
static void av1_iwht4(const int32_t *in, int32_t *out, int range) {
  (void)range;
  const int32_t x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
  const int32_t s0 = WADD(x0, x1);
  const int32_t s2 = WSUB(x2, x3);
  const int32_t s4 = WSUB(s0, s2) >> 1;
  const int32_t s3 = WSUB(s4, x3);
  const int32_t s1 = WSUB(s4, x1);
  out[0] = WSUB(s0, s3);
  out[1] = s3;
  out[2] = s1;
  out[3] = WADD(s2, s1);
}

typedef void (*inv_fn)(const int32_t *, int32_t *, int);
/* INV_TXFM_FNS[TxType1D][log2(n) - 2], inverse.rs:1593-1623 (NULL = unimplemented!()) */
static const inv_fn INV_FNS[5][5] = {
    {av1_idct4, av1_idct8, av1_idct16, av1_idct32, av1_idct64},
    {av1_iadst4, av1_iadst8, av1_iadst16, NULL, NULL},
    {av1_iflipadst4, av1_iflipadst8, av1_iflipadst16, NULL, NULL},
    {av1_iidentity4, av1_iidentity8, av1_iidentity16, av1_iidentity32, NULL},
    {av1_iwht4, NULL, NULL, NULL, NULL},
};

/* get_1d_tx_types (transform/mod.rs:342-362): {column (vertical), row (horizontal)} 1-D types;
 * TxType1D order DCT, ADST, FLIPADST, IDTX, WHT. */
static const int8_t TX1D[17][2] = {
    {0, 0}, {1, 0}, {0, 1}, {1, 1}, {2, 0}, {0, 2}, {2, 2}, {1, 2}, {2, 1},
    {3, 3}, {0, 3}, {3, 0}, {1, 3}, {3, 1}, {2, 3}, {3, 2}, {4, 4}};

/* inverse.rs:1710-1711 */
static const int8_t INV_INTERMEDIATE_SHIFTS[19] = {0, 1, 2, 2, 2, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2};

static int ilog2(int v) { return 31 - __builtin_clz((unsigned)v); }

/* 1-D entry for tests: kind = TxType1D, n in {4, 8, 16, 32, 64}; returns 0 when unimplemented */
int orc_inv_txfm_1d(int kind, int n, const int32_t *in, int32_t *out, int range) {
  if (kind < 0 || kind > 4 || n < 4 || n > 64 || (n & (n - 1))) return 0;
  const inv_fn f = INV_FNS[kind][ilog2(n) - 2];
  if (!f) return 0;
  f(in, out, range);
  return 1;
}

/* rust::inverse_transform_add, inverse.rs:1637-1704.  input: coded coefficients (i16 or i32,
 * the forward transform's layout: index = col * min(h, 32) + row); dst: pixels, stride in elements. */
void orc_inverse_transform_add(const void *input, int coeff_is_i32, void *dst, ptrdiff_t dst_stride,
                               int bpp, int tx_size, int tx_type, int bd) {
  const int width = orc_tx_width(tx_size), height = orc_tx_height(tx_size);
  const int h32 = height < 32 ? height : 32, w32 = width < 32 ? width : 32;
  int32_t *buffer = (int32_t *)calloc((size_t)width * height, sizeof(int32_t));
  const int rect_type = ilog2(width) - ilog2(height); /* rect_ratio_log2, transform/mod.rs:274-276 */
  const int lossless = tx_type == 16;
  /* rows */
  int range = bd + 8;
  inv_fn f = INV_FNS[TX1D[tx_type][1]][ilog2(width) - 2];
  for (int r = 0; r < h32; r++) {
    int32_t temp_in[64] = {0};
    /* input[r..].step_by(min(h, 32)) zipped with 64 slots: column c of row r, c < min(w, 32) */
    for (int c = 0; c < w32; c++) {
      const size_t idx = (size_t)r + (size_t)c * h32;
      const int32_t raw = coeff_is_i32 ? ((const int32_t *)input)[idx] : (int32_t)((const int16_t *)input)[idx];
      int32_t val;
      if (rect_type == 1 || rect_type == -1)
        val = round_shift_i32(WMUL(raw, INV_SQRT2), SQRT2_BITS);
      else if (lossless)
        val = raw >> 2;
      else
        val = raw;
      temp_in[c] = clamp_value(val, range);
    }
    f(temp_in, buffer + (size_t)r * width, range);
  }
  /* columns */
  range = bd + 6 > 16 ? bd + 6 : 16;
  f = INV_FNS[TX1D[tx_type][0]][ilog2(height) - 2];
  const int maxv = (1 << bd) - 1;
  for (int c = 0; c < width; c++) {
    int32_t temp_in[64] = {0}, temp_out[64] = {0};
    for (int r = 0; r < height; r++)
      temp_in[r] = clamp_value(round_shift_i32(buffer[(size_t)r * width + c], INV_INTERMEDIATE_SHIFTS[tx_size]), range);
    f(temp_in, temp_out, range);
    for (int r = 0; r < height; r++) {
      const int32_t rr = lossless ? temp_out[r] : round_shift_i32(temp_out[r], 4);
      if (bpp == 1) {
        uint8_t *p = (uint8_t *)dst + (size_t)r * dst_stride + c;
        const int32_t v = WADD((int32_t)*p, rr);
        *p = (uint8_t)(v < 0 ? 0 : v > maxv ? maxv : v);
      } else {
        uint16_t *p = (uint16_t *)dst + (size_t)r * dst_stride + c;
        const int32_t v = WADD((int32_t)*p, rr);
        *p = (uint16_t)(v < 0 ? 0 : v > maxv ? maxv : v);
      }
    }
  }
  free(buffer);
}
