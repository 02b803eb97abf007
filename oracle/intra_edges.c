/* intra_edges.c — CPU restatement of rav1e's intra edge gathering (TEST INFRASTRUCTURE ONLY, see
 * oracle.h).
 *
 *   get_intra_edges                      src/partition.rs:639-898
 *   has_top_right / has_bottom_left      src/recon_intra.rs:174-255, :374-452
 *   supersample_chroma_bsize             src/partition.rs:559-598
 *   intra_mode_to_angle, ANGLE_STEP      src/predict.rs:138-150, :38
 *
 * The 22 + 22 availability bitmaps (recon_intra.rs:30-136, :258-354) are not copied: they are
 * regenerated from the rule they encode — "is the 4x4 unit diagonally above-right (below-left) of
 * the block coded before the block in the recursive Z-order partition walk of a 128x128 superblock"
 * — and tests/test_oracle_intra_edges.py checks every regenerated table against the length, first
 * bytes and sha256 of the reference's (tests/golden/reference_kats.json, intra_avail_tables).
 *
 * Parity pinning: the tables are pinned by those digests; get_intra_edges itself has no stored
 * vectors upstream ("parity unpinned"); it is cross-checked against an independent numpy model for
 * the cases the lookahead uses and by invariants (tests/test_oracle_intra_edges.py).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "oracle.h"

/* BlockSize in enum order, partition.rs:130-153 */
static const uint8_t BS_W[22] = {4, 4, 8, 8, 8, 16, 16, 16, 32, 32, 32, 64, 64, 64, 128, 128, 4, 16, 8, 32, 16, 64};
static const uint8_t BS_H[22] = {4, 8, 4, 8, 16, 8, 16, 32, 16, 32, 64, 32, 64, 128, 64, 128, 16, 4, 32, 8, 64, 16};

int orc_block_size_index(int w, int h) {
  for (int i = 0; i < 22; i++)
    if (BS_W[i] == w && BS_H[i] == h) return i;
  return -1;
}

static unsigned morton(unsigned x, unsigned y) {
  unsigned m = 0;
  for (int b = 0; b < 6; b++) m |= ((x >> b) & 1u) << (2 * b) | ((y >> b) & 1u) << (2 * b + 1);
  return m;
}

/* One bit per block of `bsize` in a 128x128 superblock, raster order, LSB first.  kind 0: top-right,
 * 1: bottom-left.  Returns the table length in bytes (out holds up to 128). */
int orc_intra_avail_table(int kind, int bsize, uint8_t *out) {
  const int bw = BS_W[bsize] / 4, bh = BS_H[bsize] / 4; /* in 4x4 units */
  const int S = bw > bh ? bw : bh;                      /* the square the block was split from */
  const int cols = 32 / bw, rows = 32 / bh;
  const int nbits = cols * rows, n = (nbits + 7) / 8;
  memset(out, 0, (size_t)n);
  for (int br = 0; br < rows; br++)
    for (int bc = 0; bc < cols; bc++) {
      const int x = bc * bw, y = br * bh, ox = x % S, oy = y % S, tx = x - ox, ty = y - oy;
      int v;
      if (kind == 0) {
        if (oy > 0) v = 0;                  /* lower part of a horizontal split: the right neighbour is later */
        else if (ox + bw < S) v = 1;        /* left part of a vertical split: directly above */
        else if (ty == 0) v = 1;            /* top row of the superblock: the row above is coded */
        else if (tx + S >= 32) v = 0;       /* right edge: the next superblock is later */
        else v = morton((unsigned)(tx + S), (unsigned)(ty - 1)) < morton((unsigned)tx, (unsigned)ty);
      } else {
        if (x == 0) v = 0;                  /* left column: decided by the caller (recon_intra.rs:421-429) */
        else if (ox > 0) v = 0;             /* right part of a vertical split: the block below is later */
        else if (oy + bh < S) v = 1;        /* upper part of a horizontal split: directly left */
        else if (ty + S >= 32) v = 0;       /* bottom row: the next superblock row is later */
        else v = morton((unsigned)(tx - 1), (unsigned)(ty + S)) < morton((unsigned)tx, (unsigned)ty);
      }
      const int i = br * cols + bc;
      out[i >> 3] |= (uint8_t)(v << (i & 7));
    }
  return n;
}

static int log2i(int v) {
  int l = 0;
  while ((1 << l) < v) l++;
  return l;
}

/* recon_intra.rs:174-255 (bsize as enum index; tx dims in pixels) */
int orc_has_top_right(int bsize, int mi_col, int mi_row, int top_available, int right_available, int tx_w,
                      int row_off, int col_off, int ss_x, int ss_y) {
  if (!top_available || !right_available) return 0;
  const int bw_unit = BS_W[bsize] >> 2;
  int plane_bw_unit = bw_unit >> ss_x;
  if (plane_bw_unit < 1) plane_bw_unit = 1;
  const int top_right_count_unit = tx_w >> 2;
  if (row_off > 0) {
    if (BS_W[bsize] > 64) {
      if (row_off == (16 >> ss_y) && col_off + top_right_count_unit == (16 >> ss_x)) return 0;
      const int plane_bw_unit_64 = 16 >> ss_x;
      return col_off % plane_bw_unit_64 + top_right_count_unit < plane_bw_unit_64;
    }
    return col_off + top_right_count_unit < plane_bw_unit;
  }
  if (col_off + top_right_count_unit < plane_bw_unit) return 1;
  const int bw_in_mi_log2 = log2i(BS_W[bsize]) - 2, bh_in_mi_log2 = log2i(BS_H[bsize]) - 2;
  const int sb_mi_size = 16;
  const int blk_row_in_sb = (mi_row & (sb_mi_size - 1)) >> bh_in_mi_log2;
  const int blk_col_in_sb = (mi_col & (sb_mi_size - 1)) >> bw_in_mi_log2;
  if (blk_row_in_sb == 0) return 1;
  if (((blk_col_in_sb + 1) << bw_in_mi_log2) >= sb_mi_size) return 0;
  const int idx = (blk_row_in_sb << (5 - bw_in_mi_log2)) + blk_col_in_sb; /* MAX_MIB_SIZE_LOG2 = 5 */
  uint8_t tab[128];
  orc_intra_avail_table(0, bsize, tab);
  return (tab[idx >> 3] >> (idx & 7)) & 1;
}

/* recon_intra.rs:374-452 */
int orc_has_bottom_left(int bsize, int mi_col, int mi_row, int bottom_available, int left_available, int tx_h,
                        int row_off, int col_off, int ss_x, int ss_y) {
  if (!bottom_available || !left_available) return 0;
  if (BS_W[bsize] > 64 && col_off > 0) {
    const int plane_bw_unit_64 = 16 >> ss_x;
    if (col_off % plane_bw_unit_64 == 0) {
      const int plane_bh_unit_64 = 16 >> ss_y;
      int plane_bh_unit = (BS_H[bsize] >> 2) >> ss_y;
      if (plane_bh_unit > plane_bh_unit_64) plane_bh_unit = plane_bh_unit_64;
      return row_off % plane_bh_unit_64 + (tx_h >> 2) < plane_bh_unit;
    }
  }
  if (col_off > 0) return 0;
  int plane_bh_unit = (BS_H[bsize] >> 2) >> ss_y;
  if (plane_bh_unit < 1) plane_bh_unit = 1;
  const int bottom_left_count_unit = tx_h >> 2;
  if (row_off + bottom_left_count_unit < plane_bh_unit) return 1;
  const int bw_in_mi_log2 = log2i(BS_W[bsize]) - 2, bh_in_mi_log2 = log2i(BS_H[bsize]) - 2;
  const int sb_mi_size = 16;
  const int blk_row_in_sb = (mi_row & (sb_mi_size - 1)) >> bh_in_mi_log2;
  const int blk_col_in_sb = (mi_col & (sb_mi_size - 1)) >> bw_in_mi_log2;
  if (blk_col_in_sb == 0) {
    const int blk_start_row_off = (blk_row_in_sb << bh_in_mi_log2) >> ss_y;
    return blk_start_row_off + row_off + bottom_left_count_unit < (sb_mi_size >> ss_y);
  }
  if (((blk_row_in_sb + 1) << bh_in_mi_log2) >= sb_mi_size) return 0;
  const int idx = (blk_row_in_sb << (5 - bw_in_mi_log2)) + blk_col_in_sb;
  uint8_t tab[128];
  orc_intra_avail_table(1, bsize, tab);
  return (tab[idx >> 3] >> (idx & 7)) & 1;
}

/* partition.rs:559-598 */
static int supersample_chroma_bsize(int bsize, int ss_x, int ss_y) {
  const int w = BS_W[bsize], h = BS_H[bsize];
  int nw = w, nh = h;
  if (w == 4 && h == 4) {
    nw = ss_x ? 8 : 4, nh = ss_y ? 8 : 4;
  } else if (w == 4 && h == 8) {
    nw = ss_x ? 8 : 4;
  } else if (w == 8 && h == 4) {
    nh = ss_y ? 8 : 4;
  } else if (w == 4 && h == 16) {
    nw = ss_x ? 8 : 4;
  } else if (w == 16 && h == 4) {
    nh = ss_y ? 8 : 4;
  }
  return orc_block_size_index(nw, nh);
}

static int mode_angle(int mode) { /* predict.rs:138-150 */
  static const int a[9] = {0, 90, 180, 45, 135, 113, 157, 203, 67};
  return mode >= 1 && mode <= 8 ? a[mode] : 0;
}

#define PX(yy, xx) (bpp == 1 ? (uint32_t)((const uint8_t *)region)[(ptrdiff_t)(yy) * stride + (xx)] \
                             : (uint32_t)((const uint16_t *)region)[(ptrdiff_t)(yy) * stride + (xx)])
#define PUT(i, v)                                   \
  do {                                              \
    if (bpp == 1) ((uint8_t *)edge)[i] = (uint8_t)(v); \
    else ((uint16_t *)edge)[i] = (uint16_t)(v);     \
  } while (0)
#define GET(i) (bpp == 1 ? (uint32_t)((uint8_t *)edge)[i] : (uint32_t)((uint16_t *)edge)[i])

/* get_intra_edges, partition.rs:639-898.  `region` addresses pixel (0,0) of `dst` (a PlaneRegion whose
 * rect is (rect_x, rect_y, rect_w, rect_h) inside a plane of plane_w x plane_h pixels, decimated by
 * xdec / ydec); stride in elements.  mode < 0 = None.  Entries outside
 * [128 - init_left, 129 + init_above) are left untouched (MaybeUninit upstream). */
void orc_get_intra_edges(void *edge, const void *region, ptrdiff_t stride, int bpp, int plane_w, int plane_h,
                         int rect_x, int rect_y, int rect_w, int rect_h, int xdec, int ydec, int part_bo_x,
                         int part_bo_y, int bx, int by, int partition_bsize, int po_x, int po_y, int tx_w,
                         int tx_h, int bit_depth, int mode, int enable_intra_edge_filter, int angle_delta,
                         int *out_init_left, int *out_init_above) {
  enum { MAXTX = 64, L0 = 2 * MAXTX };  /* left ends at index 127, top-left at 128, above from 129 */
  int init_left = 0, init_above = 0;
  const uint32_t base = 128u << (bit_depth - 8);
  const int x = po_x, y = po_y;
  int needs_left = 1, needs_topleft = 1, needs_top = 1, needs_topright = 1, needs_bottomleft = 1,
      needs_topleft_filter = 0;
  if (mode >= 0) {
    int m = mode;
    if (m == 12) m = (x == 0 && y == 0) ? 0 : x == 0 ? 1 : y == 0 ? 2 : 12; /* PAETH, :670-679 */
    const int p_angle = mode_angle(m) + angle_delta * 3;
    const int dc_or_cfl = m == 0 || m == 13;
    const int directional = m >= 1 && m <= 8;
    needs_left = (!dc_or_cfl || x != 0) || (p_angle > 90 && p_angle != 180);
    needs_topleft = m == 12 || (directional && p_angle != 90 && p_angle != 180);
    needs_top = (!dc_or_cfl || y != 0) || (p_angle != 90 && p_angle < 180);
    needs_topright = directional && p_angle < 90;
    needs_bottomleft = directional && p_angle > 180;
    needs_topleft_filter = enable_intra_edge_filter && p_angle > 90 && p_angle < 180;
  }
  const int rw = rect_w < plane_w - rect_x ? rect_w : plane_w - rect_x; /* :705-708 */
  const int rh = rect_h < plane_h - rect_y ? rect_h : plane_h - rect_y;
  if (needs_left) { /* :711-733 */
    const int txh = y + tx_h > rh ? rh - y : tx_h;
    if (x != 0) {
      for (int i = 0; i < txh; i++) PUT(L0 - 1 - i, PX(y + i, x - 1));
      if (txh < tx_h) {
        const uint32_t val = PX(y + txh - 1, x - 1);
        for (int i = txh; i < tx_h; i++) PUT(L0 - 1 - i, val);
      }
    } else {
      const uint32_t val = y != 0 ? PX(y - 1, 0) : base + 1;
      for (int i = 0; i < tx_h; i++) PUT(L0 - 1 - i, val);
    }
    init_left += tx_h;
  }
  if (needs_top) { /* :736-762 */
    const int txw = x + tx_w > rw ? rw - x : tx_w;
    if (y != 0) {
      for (int i = 0; i < txw; i++) PUT(L0 + 1 + i, PX(y - 1, x + i));
      if (txw < tx_w) {
        const uint32_t val = PX(y - 1, x + txw - 1);
        for (int i = txw; i < tx_w; i++) PUT(L0 + 1 + i, val);
      }
    } else {
      const uint32_t val = x != 0 ? PX(0, x - 1) : base - 1;
      for (int i = 0; i < tx_w; i++) PUT(L0 + 1 + i, val);
    }
    init_above += tx_w;
  }
  const int bx4 = bx * (tx_w >> 2), by4 = by * (tx_h >> 2);
  const int have_top = by4 != 0 || (ydec ? part_bo_y > 1 : part_bo_y > 0);
  const int have_left = bx4 != 0 || (xdec ? part_bo_x > 1 : part_bo_x > 0);
  const int right_available = x + tx_w < rw, bottom_available = y + tx_h < rh;
  const int scaled = supersample_chroma_bsize(partition_bsize, xdec, ydec);
  if (needs_topright) { /* :789-829 */
    int num_avail = 0;
    if (y != 0 && orc_has_top_right(scaled, part_bo_x, part_bo_y, have_top, right_available, tx_w, by4, bx4, xdec, ydec)) {
      num_avail = rw - x - tx_w;
      if (num_avail > tx_w) num_avail = tx_w;
    }
    for (int i = 0; i < num_avail; i++) PUT(L0 + 1 + tx_w + i, PX(y - 1, x + tx_w + i));
    if (num_avail < tx_h) {
      const uint32_t val = GET(L0 + 1 + tx_w + num_avail - 1);
      for (int i = tx_w + num_avail; i < tx_w + tx_h; i++) PUT(L0 + 1 + i, val);
    }
    init_above += tx_h;
  }
  if (needs_bottomleft) { /* :835-870 */
    int num_avail = 0;
    if (x != 0 && orc_has_bottom_left(scaled, part_bo_x, part_bo_y, bottom_available, have_left, tx_h, by4, bx4, xdec, ydec)) {
      num_avail = rh - y - tx_h;
      if (num_avail > tx_h) num_avail = tx_h;
    }
    for (int i = 0; i < num_avail; i++) PUT(L0 - tx_h - 1 - i, PX(y + tx_h + i, x - 1));
    if (num_avail < tx_w) {
      const uint32_t val = GET(L0 - tx_h - num_avail);
      for (int i = L0 - tx_h - tx_w; i < L0 - tx_h - num_avail; i++) PUT(i, val);
    }
    init_left += tx_w;
  }
  if (needs_topleft) { /* :878-894 */
    uint32_t tl = (x == 0 && y == 0) ? base : y == 0 ? PX(0, x - 1) : x == 0 ? PX(y - 1, 0) : PX(y - 1, x - 1);
    if (needs_topleft_filter && tx_w + tx_h >= 24) {
      const uint32_t l = GET(L0 - 1), a = GET(L0 + 1);
      tl = (l * 5 + tl * 6 + a * 5 + (1u << 3)) >> 4;
    }
    PUT(L0, tl);
  } else {
    PUT(L0, base);
  }
  *out_init_left = init_left;
  *out_init_above = init_above;
}
