/*
 * oracle/predict.c — restatement of rav1e's intra predictors, src/predict.rs:
 *   dispatch_predict_intra (:705-784), pred_dc/_128/_left/_top (:786-840), pred_h (:842),
 *   pred_v (:851), pred_paeth (:860-887), pred_smooth/_h/_v (:889-1018), pred_cfl_ac
 *   (:1020-1063), pred_cfl_inner + pred_cfl* (:1065-1123), select_ief_strength (:1125-1186),
 *   select_ief_upsample (:1188-1201), filter_edge (:1203-1232), upsample_edge (:1234-1266),
 *   dr_intra_derivative (:1268-1299), pred_directional (:1301-1505), sm_weight_arrays
 *   (:603-624), get_scaled_luma_q0 (:626-635); IntraEdge layout src/partition.rs:600-637.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Pinned by the reference KATs src/predict.rs:1514-1619 (4x4 DC/DC_TOP/DC_LEFT/DC_128/V/H/
 * Paeth/smooth x3 + 27 directional angles) and :1622-1693 (12-bit saturation):
 * tests/test_oracle_predict.py.  Edge filter / upsample / CfL paths have no stored vectors
 * upstream (AV1-normative; "parity unpinned" for those sub-paths).
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

#define MAX_TX_SIZE 64

enum { /* predict.rs:73-87 */
  DC_PRED = 0, V_PRED, H_PRED, D45_PRED, D135_PRED, D113_PRED, D157_PRED, D203_PRED, D67_PRED,
  SMOOTH_PRED, SMOOTH_V_PRED, SMOOTH_H_PRED, PAETH_PRED, UV_CFL_PRED
};
enum { VAR_NONE = 0, VAR_LEFT, VAR_TOP, VAR_BOTH }; /* predict.rs:112-118 */

/* predict.rs:603-624 (AV1 spec Sm_Weights_Tx_*), indexed [size + i] */
static const uint8_t sm_weight_arrays[2 * MAX_TX_SIZE] = {
    0, 0,
    255, 128,
    255, 149, 85, 64,
    255, 197, 146, 105, 73, 50, 37, 32,
    255, 225, 196, 170, 145, 123, 102, 84, 68, 54, 43, 33, 26, 20, 17, 16,
    255, 240, 225, 210, 196, 182, 169, 157, 145, 133, 122, 111, 101, 92, 83, 74,
    66, 59, 52, 45, 39, 34, 29, 25, 21, 17, 14, 12, 10, 9, 8, 8,
    255, 248, 240, 233, 225, 218, 210, 203, 196, 189, 182, 176, 169, 163, 156,
    150, 144, 138, 133, 127, 121, 116, 111, 106, 101, 96, 91, 86, 82, 77, 73, 69,
    65, 61, 57, 54, 50, 47, 44, 41, 38, 35, 32, 29, 27, 25, 22, 20, 18, 16, 15,
    13, 12, 10, 9, 8, 7, 6, 6, 5, 5, 4, 4, 4,
};

/* predict.rs:626-635 */
static inline int32_t get_scaled_luma_q0(int16_t alpha_q3, int16_t ac_pred_q3) {
  int32_t scaled_luma_q6 = (int32_t)alpha_q3 * (int32_t)ac_pred_q3;
  int32_t a = (abs(scaled_luma_q6) + 32) >> 6;
  return scaled_luma_q6 < 0 ? -a : a;
}

/* predict.rs:1125-1186 */
static int select_ief_strength(int width, int height, int smooth_filter, int angle_delta) {
  int block_wh = width + height;
  int abs_delta = abs(angle_delta);
  if (smooth_filter) {
    if (block_wh <= 8) {
      if (abs_delta >= 64) return 2;
      if (abs_delta >= 40) return 1;
    } else if (block_wh <= 16) {
      if (abs_delta >= 48) return 2;
      if (abs_delta >= 20) return 1;
    } else if (block_wh <= 24) {
      if (abs_delta >= 4) return 3;
    } else {
      return 3;
    }
  } else {
    if (block_wh <= 8) {
      if (abs_delta >= 56) return 1;
    } else if (block_wh <= 16) {
      if (abs_delta >= 40) return 1;
    } else if (block_wh <= 24) {
      if (abs_delta >= 32) return 3;
      if (abs_delta >= 16) return 2;
      if (abs_delta >= 8) return 1;
    } else if (block_wh <= 32) {
      if (abs_delta >= 32) return 3;
      if (abs_delta >= 4) return 2;
      return 1;
    } else {
      return 3;
    }
  }
  return 0;
}

/* predict.rs:1188-1201 */
static int select_ief_upsample(int width, int height, int smooth_filter, int angle_delta) {
  int block_wh = width + height;
  int abs_delta = abs(angle_delta);
  if (abs_delta == 0 || abs_delta >= 40) return 0;
  return smooth_filter ? block_wh <= 8 : block_wh <= 16;
}

/* predict.rs:1268-1299 */
static int dr_intra_derivative(int p_angle) {
  switch (p_angle) {
    case 3: return 1023; case 6: return 547; case 9: return 372; case 14: return 273;
    case 17: return 215; case 20: return 178; case 23: return 151; case 26: return 132;
    case 29: return 116; case 32: return 102; case 36: return 90; case 39: return 80;
    case 42: return 71; case 45: return 64; case 48: return 57; case 51: return 51;
    case 54: return 45; case 58: return 40; case 61: return 35; case 64: return 31;
    case 67: return 27; case 70: return 23; case 73: return 19; case 76: return 15;
    case 81: return 11; case 84: return 7; case 87: return 3;
    default: return 0;
  }
}

#define PIXEL uint8_t
#define SFX(name) name##_u8
#include "predict_impl.h"
#undef PIXEL
#undef SFX
#define PIXEL uint16_t
#define SFX(name) name##_u16
#include "predict_impl.h"
#undef PIXEL
#undef SFX

void orc_predict_intra(int mode, int variant, void *dst, ptrdiff_t dst_stride, int bpp, int w,
                       int h, int bit_depth, const int16_t *ac, int angle, int ief,
                       const void *edge, int left_len, int above_len, int plane_w, int plane_h,
                       int dst_x, int dst_y) {
  if (bpp == 1)
    dispatch_u8(mode, variant, (uint8_t *)dst, dst_stride, w, h, bit_depth, ac, angle, ief,
                (const uint8_t *)edge, left_len, above_len, plane_w, plane_h, dst_x, dst_y);
  else
    dispatch_u16(mode, variant, (uint16_t *)dst, dst_stride, w, h, bit_depth, ac, angle, ief,
                 (const uint16_t *)edge, left_len, above_len, plane_w, plane_h, dst_x, dst_y);
}

void orc_pred_cfl_ac(int16_t *ac, const void *luma, ptrdiff_t luma_stride, int bpp, int bw, int bh,
                     int w_pad, int h_pad, int xdec, int ydec) {
  if (bpp == 1)
    pred_cfl_ac_u8(ac, (const uint8_t *)luma, luma_stride, bw, bh, w_pad, h_pad, xdec, ydec);
  else
    pred_cfl_ac_u16(ac, (const uint16_t *)luma, luma_stride, bw, bh, w_pad, h_pad, xdec, ydec);
}

/* Batched form for the CPU baseline: items mirror b200_intra_item (include/b200rdo.h). */
typedef struct {
  uint32_t edge, ac;
  int16_t x, y, angle;
  uint8_t mode, variant;
  int8_t ief;
  uint8_t left_len, above_len, pad_;
} orc_intra_item;

void orc_predict_intra_batch(const void *edges, int bpp, const orc_intra_item *items, size_t n,
                             const int16_t *ac, int w, int h, int bit_depth, int plane_w, int plane_h,
                             void *out) {
#pragma omp parallel for schedule(static)
  for (ptrdiff_t i = 0; i < (ptrdiff_t)n; i++) {
    const orc_intra_item it = items[i];
    orc_predict_intra(it.mode, it.variant, (uint8_t *)out + (size_t)i * w * h * bpp, w, bpp, w, h,
                      bit_depth, ac ? ac + (size_t)it.ac * w * h : 0, it.angle, it.ief,
                      (const uint8_t *)edges + (size_t)it.edge * 257 * bpp, it.left_len, it.above_len,
                      plane_w, plane_h, it.x, it.y);
  }
}
