/*
 * oracle/fwd_txfm.c — restatement of rav1e's 2-D forward transform driver
 *   src/transform/forward.rs:71-161 (rust::forward_transform)
 *   src/transform/forward_shared.rs:22-165 (shift tables, Txfm2DFlipCfg::fwd, flips)
 *   src/transform/mod.rs:56-123 (TxType / TxSize), :320-336 (av1_round_shift_array),
 *   :364-417 (VTX_TAB / HTX_TAB / valid_av1_transform).
 * The 1-D networks live in txfm_networks.h (generated restatement, see its header).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Pinning: the reference stores no golden coefficients ("parity unpinned"); this file is
 * cross-checked in tests/test_oracle_txfm.py against double-precision orthonormal
 * DCT-II / DST-IV / DST-VII (each 1-D network within a few LSB) and the 2-D scale law
 * 2^(shift0+shift1+shift2) stated in SURVEY §8c.
 */
#include "oracle.h"
#include "txfm_networks.h"

#include <string.h>

enum { T1_DCT = 0, T1_ADST = 1, T1_FLIPADST = 2, T1_IDTX = 3, T1_WHT = 4 };
enum { TX_DCT_DCT = 0, TX_IDTX = 9, TX_WHT_WHT = 16 };

/* transform/mod.rs:101-123 declaration order */
static const uint8_t TX_W[19] = {4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64};
static const uint8_t TX_H[19] = {4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16};

/* transform/mod.rs:364-402 */
static const uint8_t VTX_TAB[17] = {T1_DCT, T1_ADST, T1_DCT, T1_ADST, T1_FLIPADST, T1_DCT,
                                    T1_FLIPADST, T1_ADST, T1_FLIPADST, T1_IDTX, T1_DCT, T1_IDTX,
                                    T1_ADST, T1_IDTX, T1_FLIPADST, T1_IDTX, T1_WHT};
static const uint8_t HTX_TAB[17] = {T1_DCT, T1_DCT, T1_ADST, T1_ADST, T1_DCT, T1_FLIPADST,
                                    T1_FLIPADST, T1_FLIPADST, T1_ADST, T1_IDTX, T1_IDTX, T1_DCT,
                                    T1_IDTX, T1_ADST, T1_IDTX, T1_FLIPADST, T1_WHT};

/* forward_shared.rs:22-64: [tx_size][(bd-8)/2][3] */
#define S3(a, b, c, d, e, f, g, h, i) {{a, b, c}, {d, e, f}, {g, h, i}}
static const int8_t FWD_SHIFT[19][3][3] = {
    S3(3, 0, 0, 2, 0, 1, 0, 0, 3),    /* 4x4 */
    S3(4, -1, 0, 2, 0, 1, 0, 0, 3),   /* 8x8 */
    S3(4, -1, 0, 2, 0, 1, 0, 0, 3),   /* 16x16 */
    S3(4, -2, 0, 2, 0, 0, 0, 0, 2),   /* 32x32 */
    S3(4, -1, -2, 2, 0, -1, 0, 0, 1), /* 64x64 */
    S3(4, -1, 0, 2, 0, 1, 0, 0, 3),   /* 4x8 */
    S3(4, -1, 0, 2, 0, 1, 0, 0, 3),   /* 8x4 */
    S3(4, -1, 0, 2, 0, 1, 0, 0, 3),   /* 8x16 */
    S3(4, -1, 0, 2, 0, 1, 0, 0, 3),   /* 16x8 */
    S3(4, -2, 0, 2, 0, 0, 0, 0, 2),   /* 16x32 */
    S3(4, -2, 0, 2, 0, 0, 0, 0, 2),   /* 32x16 */
    S3(4, -1, -2, 2, 0, -1, 0, 0, 1), /* 32x64 */
    S3(4, -1, -2, 2, 0, -1, 0, 0, 1), /* 64x32 */
    S3(4, -1, 0, 2, 0, 1, 0, 0, 3),   /* 4x16 */
    S3(4, -1, 0, 2, 0, 1, 0, 0, 3),   /* 16x4 */
    S3(4, -1, 0, 2, 0, 1, 0, 0, 3),   /* 8x32 */
    S3(4, -1, 0, 2, 0, 1, 0, 0, 3),   /* 32x8 */
    S3(4, -2, 0, 2, 0, 0, 0, 0, 2),   /* 16x64 */
    S3(4, -2, 0, 2, 0, 0, 0, 0, 2),   /* 64x16 */
};
static const int8_t FWD_SHIFT_WHT[3] = {0, 0, 2}; /* forward_shared.rs:42 */

int orc_tx_width(int tx_size) { return TX_W[tx_size]; }
int orc_tx_height(int tx_size) { return TX_H[tx_size]; }

static int size_index(int n) { /* width_index/height_index: log2(n) - 2 */
  return n == 4 ? 0 : n == 8 ? 1 : n == 16 ? 2 : n == 32 ? 3 : 4;
}

/* mod.rs:405-417 valid_av1_transform AND the `.unwrap()` in Txfm2DFlipCfg::fwd
 * (forward_shared.rs:131-134): a 1-D type must exist for both dimensions. */
int orc_valid_av1_transform(int tx_size, int tx_type) {
  if (tx_size < 0 || tx_size >= 19 || tx_type < 0 || tx_type > 16) return 0;
  int w = TX_W[tx_size], h = TX_H[tx_size];
  int m = w > h ? w : h;
  if (m == 64 && tx_type != TX_DCT_DCT) return 0;
  if (m == 32 && tx_type != TX_DCT_DCT && tx_type != TX_IDTX) return 0;
  /* AV1_TXFM_TYPE_LS (forward_shared.rs:85-109) */
  int t1[2] = {VTX_TAB[tx_type], HTX_TAB[tx_type]};
  int n[2] = {h, w};
  for (int k = 0; k < 2; k++) {
    int idx = size_index(n[k]);
    if (t1[k] == T1_WHT && idx != 0) return 0;
    if ((t1[k] == T1_ADST || t1[k] == T1_FLIPADST) && idx > 2) return 0;
    if (t1[k] == T1_IDTX && idx > 3) return 0;
  }
  return 1;
}

static void txfm_1d(int type1d, int n, int32_t *c) {
  switch (type1d) {
    case T1_DCT:
      if (n == 4) orc_fdct4(c);
      else if (n == 8) orc_fdct8(c);
      else if (n == 16) orc_fdct16(c);
      else if (n == 32) orc_fdct32(c);
      else orc_fdct64(c);
      break;
    case T1_ADST:
    case T1_FLIPADST:
      if (n == 4) orc_fdst_vii_4(c);
      else if (n == 8) orc_fdst8(c);
      else orc_fdst16(c);
      break;
    case T1_IDTX: /* fidentity: no-op, forward_shared.rs:1775 */
      break;
    case T1_WHT:
      orc_fwht4(c);
      break;
  }
}

/* mod.rs:320-336 av1_round_shift_array: bit > 0 rounds down-shift, bit < 0 shifts up */
static void round_shift_array(int32_t *a, int n, int bit) {
  if (bit == 0) return;
  if (bit > 0)
    for (int i = 0; i < n; i++) a[i] = (a[i] + ((1 << bit) >> 1)) >> bit;
  else
    for (int i = 0; i < n; i++) a[i] = (int32_t)((uint32_t)a[i] << (-bit));
}

/* forward.rs:71-161 */
void orc_forward_transform(const int16_t *input, void *output, size_t stride, int tx_size,
                           int tx_type, int bd, int coeff_is_i32) {
  const int w = TX_W[tx_size], h = TX_H[tx_size];
  const int8_t *shift = tx_type == TX_WHT_WHT ? FWD_SHIFT_WHT : FWD_SHIFT[tx_size][(bd - 8) / 2];
  const int col_type = VTX_TAB[tx_type], row_type = HTX_TAB[tx_type];
  /* forward_shared.rs:155-164 get_flip_cfg */
  const int ud_flip = col_type == T1_FLIPADST, lr_flip = row_type == T1_FLIPADST;
  static _Thread_local int32_t buf[64 * 64];

  for (int c = 0; c < w; c++) { /* columns, forward.rs:95-126 */
    int32_t col[64];
    for (int r = 0; r < h; r++)
      col[r] = ud_flip ? input[(size_t)(h - r - 1) * stride + c] : input[(size_t)r * stride + c];
    round_shift_array(col, h, -shift[0]);
    txfm_1d(col_type, h, col);
    round_shift_array(col, h, -shift[1]);
    for (int r = 0; r < h; r++) buf[r * w + (lr_flip ? w - c - 1 : c)] = col[r];
  }
  for (int r = 0; r < h; r++) { /* rows, forward.rs:131-160 */
    int32_t *row = buf + r * w;
    txfm_1d(row_type, w, row);
    round_shift_array(row, w, -shift[2]);
    const int output_stride = h < 32 ? h : 32;
    const int wc = w < 32 ? w : 32;
    size_t base = (r >= 32) ? (size_t)output_stride * wc : 0;
    for (int cg = 0; cg < w; cg += 32) {
      size_t b2 = base + (size_t)h * cg;
      for (int c = 0; c < wc; c++) {
        size_t idx = b2 + (size_t)c * output_stride + (r & 31);
        if (coeff_is_i32)
          ((int32_t *)output)[idx] = row[c + cg];
        else
          ((int16_t *)output)[idx] = (int16_t)row[c + cg]; /* `as i16`: truncating */
      }
    }
  }
}

void orc_forward_transform_batch(const int16_t *input, void *output, size_t nblocks, int tx_size,
                                 int tx_type, int bd, int coeff_is_i32, int threads) {
  const size_t area = (size_t)TX_W[tx_size] * TX_H[tx_size];
  const size_t osz = coeff_is_i32 ? 4 : 2;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : orc_num_threads())
  for (ptrdiff_t i = 0; i < (ptrdiff_t)nblocks; i++)
    orc_forward_transform(input + (size_t)i * area, (uint8_t *)output + (size_t)i * area * osz,
                          TX_W[tx_size], tx_size, tx_type, bd, coeff_is_i32);
}
