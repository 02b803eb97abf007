/* lookahead.c — CPU restatement of the frame-wide, dependence-free consumers of the SATD / intra
 * kernels in rav1e's lookahead (TEST INFRASTRUCTURE ONLY, see oracle.h).
 *
 *   Plane::downsampled + Plane::pad        v_frame 0.3.9 plane.rs (off disk; used at src/encoder.rs:476-477
 *                                          to build the half / quarter resolution ME pyramid)
 *   estimate_intra_costs                   src/api/lookahead.rs:30-128
 *   estimate_importance_block_difference   src/api/lookahead.rs:131-180
 *   estimate_inter_costs (cost part)       src/api/lookahead.rs:238-270 (the motion vectors it reads come
 *                                          from estimate_tile_motion: an input here)
 *
 * v_frame is not on disk: downsampled() / pad() are restated from the crate's published source
 * (2x2 box filter `(a + b + c + d + 2) >> 2` over the source rows 2r, 2r+1 and columns 2c, 2c+1, new
 * size ((w + 1) / 2, (h + 1) / 2); pad(w, h) replicates column (w + xdec >> xdec) - 1 / row
 * (h + ydec >> ydec) - 1 outwards and the first column / row into the leading padding).
 * "parity unpinned": no test under /root/reference fixes these outputs.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include "oracle.h"

#define LPX(p, stride, bpp, y, x)                                      \
  ((bpp) == 1 ? (uint32_t)((const uint8_t *)(p))[(ptrdiff_t)(y) * (stride) + (x)] \
              : (uint32_t)((const uint16_t *)(p))[(ptrdiff_t)(y) * (stride) + (x)])

/* src: pixel (0,0) of the source plane (readable one pixel past width / height when they are odd:
 * the source is padded).  dst: pixel (0,0) of a plane with `dst_pad` readable pixels on every side;
 * its visible size is ((src_w + 1) / 2, (src_h + 1) / 2); (pad_w, pad_h) = the size Plane::pad
 * replicates from: ((frame_w + xdec) >> xdec, (frame_h + ydec) >> ydec) of the NEW plane. */
void orc_plane_downsample(const void *src, ptrdiff_t src_stride, int src_w, int src_h, void *dst,
                          ptrdiff_t dst_stride, int dst_pad, int bpp, int pad_w, int pad_h) {
  const int w = (src_w + 1) / 2, h = (src_h + 1) / 2;
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) {
      const uint32_t sum = LPX(src, src_stride, bpp, 2 * r, 2 * c) + LPX(src, src_stride, bpp, 2 * r, 2 * c + 1) +
                           LPX(src, src_stride, bpp, 2 * r + 1, 2 * c) + LPX(src, src_stride, bpp, 2 * r + 1, 2 * c + 1);
      const uint32_t avg = (sum + 2) >> 2;
      if (bpp == 1) ((uint8_t *)dst)[(ptrdiff_t)r * dst_stride + c] = (uint8_t)avg;
      else ((uint16_t *)dst)[(ptrdiff_t)r * dst_stride + c] = (uint16_t)avg;
    }
  /* Plane::pad(pad_w, pad_h): left / right of the first pad_h rows, then whole rows up / down */
  for (int r = 0; r < pad_h; r++)
    for (int c = -dst_pad; c < w + dst_pad; c++) {
      if (c >= 0 && c < pad_w) continue;
      const int sc = c < 0 ? 0 : pad_w - 1;
      if (bpp == 1) ((uint8_t *)dst)[(ptrdiff_t)r * dst_stride + c] = ((uint8_t *)dst)[(ptrdiff_t)r * dst_stride + sc];
      else ((uint16_t *)dst)[(ptrdiff_t)r * dst_stride + c] = ((uint16_t *)dst)[(ptrdiff_t)r * dst_stride + sc];
    }
  for (int r = -dst_pad; r < h + dst_pad; r++) {
    if (r >= 0 && r < pad_h) continue;
    const int sr = r < 0 ? 0 : pad_h - 1;
    for (int c = -dst_pad; c < w + dst_pad; c++) {
      if (bpp == 1) ((uint8_t *)dst)[(ptrdiff_t)r * dst_stride + c] = ((uint8_t *)dst)[(ptrdiff_t)sr * dst_stride + c];
      else ((uint16_t *)dst)[(ptrdiff_t)r * dst_stride + c] = ((uint16_t *)dst)[(ptrdiff_t)sr * dst_stride + c];
    }
  }
}

/* lookahead.rs:30-128: per 8x8 importance block get_intra_edges(DC_PRED, TX_8X8) on the SOURCE luma,
 * DC prediction (variant from the block's position, predict.rs:126-135), get_satd against the source.
 * costs: (height / 8) x (width / 8) u32, row-major. */
void orc_estimate_intra_costs(const void *luma, ptrdiff_t stride, int width, int height, int bpp, int bit_depth,
                              uint32_t *costs) {
  const int wb = width / 8, hb = height / 8;
  const int bs8 = orc_block_size_index(8, 8);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < hb; y++)
    for (int x = 0; x < wb; x++) {
      uint16_t edge16[257];
      uint8_t edge8[257];
      void *edge = bpp == 1 ? (void *)edge8 : (void *)edge16;
      int il, ia;
      orc_get_intra_edges(edge, luma, stride, bpp, width, height, 0, 0, width, height, 0, 0, x, y, 0, 0, bs8, 8 * x,
                          8 * y, 8, 8, bit_depth, 0 /* DC_PRED */, 0, 0, &il, &ia);
      const int variant = (x != 0 && y != 0) ? 3 : x != 0 ? 1 : y != 0 ? 2 : 0;
      uint16_t pred16[64];
      uint8_t pred8[64];
      void *pred = bpp == 1 ? (void *)pred8 : (void *)pred16;
      orc_predict_intra(0, variant, pred, 8, bpp, 8, 8, bit_depth, NULL, 0, -1, edge, il, ia, width, height, 8 * x, 8 * y);
      const void *org = (const uint8_t *)luma + ((ptrdiff_t)(8 * y) * stride + 8 * x) * bpp;
      costs[y * wb + x] = bpp == 1 ? orc_get_satd_u8((const uint8_t *)org, stride, pred8, 8, 8, 8)
                                   : orc_get_satd_u16((const uint16_t *)org, stride, pred16, 8, 8, 8);
    }
}

/* lookahead.rs:131-180 */
double orc_importance_block_difference(const void *org, ptrdiff_t org_stride, const void *ref, ptrdiff_t ref_stride,
                                       int width, int height, int bpp) {
  const int wb = width / 8, hb = height / 8;
  uint64_t total = 0;
  for (int y = 0; y < hb; y++)
    for (int x = 0; x < wb; x++) {
      int64_t so = 0, sr = 0;
      for (int r = 0; r < 8; r++) {
        uint16_t ro = 0, rr = 0; /* u16 row sums, :157-160 */
        for (int c = 0; c < 8; c++) {
          ro = (uint16_t)(ro + LPX(org, org_stride, bpp, 8 * y + r, 8 * x + c));
          rr = (uint16_t)(rr + LPX(ref, ref_stride, bpp, 8 * y + r, 8 * x + c));
        }
        so += ro;
        sr += rr;
      }
      const int64_t count = 64;
      int64_t mean = (so + count / 2) / count - (sr + count / 2) / count;
      if (mean < 0) mean = -mean;
      total += (uint64_t)mean;
    }
  return (double)total / (double)((size_t)wb * (size_t)hb);
}

/* lookahead.rs:238-270: per importance block the SATD between the source block and the reference block
 * displaced by the block's motion vector (mvs: (row, col) int16 per importance block, row-major; the
 * reference reads stats[y * 2][x * 2].mv); region origin = (8 x * 8 + mv.col) / 8 with the division
 * truncating toward zero (`as isize / 8`).  costs (may be NULL) receives the per-block SATDs. */
double orc_estimate_inter_costs(const void *org, ptrdiff_t org_stride, const void *ref, ptrdiff_t ref_stride,
                                int width, int height, int bpp, const int16_t *mvs, uint32_t *costs) {
  const int wb = width / 8, hb = height / 8;
  uint64_t total = 0;
  for (int y = 0; y < hb; y++)
    for (int x = 0; x < wb; x++) {
      const int16_t mr = mvs[2 * (y * wb + x)], mc = mvs[2 * (y * wb + x) + 1];
      const int64_t reference_x = (int64_t)x * 64 + mc, reference_y = (int64_t)y * 64 + mr;
      const ptrdiff_t rx = (ptrdiff_t)(reference_x / 8), ry = (ptrdiff_t)(reference_y / 8); /* C division truncates */
      const void *o = (const uint8_t *)org + ((ptrdiff_t)(8 * y) * org_stride + 8 * x) * bpp;
      const void *r = (const uint8_t *)ref + (ry * ref_stride + rx) * bpp;
      const uint32_t c = bpp == 1 ? orc_get_satd_u8((const uint8_t *)o, org_stride, (const uint8_t *)r, ref_stride, 8, 8)
                                  : orc_get_satd_u16((const uint16_t *)o, org_stride, (const uint16_t *)r, ref_stride, 8, 8);
      if (costs) costs[y * wb + x] = c;
      total += c;
    }
  return (double)total / (double)((size_t)wb * (size_t)hb);
}
