/* Included twice by predict.c with PIXEL = uint8_t / uint16_t.  Strides in elements.
 * Slices follow the reference: `left` is ordered bottom -> top (left[len-1] touches the top-left
 * pixel), `above` left -> right (partition.rs:600-637). */

static void SFX(fill)(PIXEL *out, ptrdiff_t stride, int w, int h, PIXEL v) {
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) out[r * stride + c] = v;
}

/* predict.rs:786-803 */
static void SFX(pred_dc)(PIXEL *out, ptrdiff_t stride, const PIXEL *above, const PIXEL *left, int w, int h) {
  uint32_t sum = 0, len = (uint32_t)(w + h);
  for (int i = 0; i < h; i++) sum += left[i];
  for (int i = 0; i < w; i++) sum += above[i];
  SFX(fill)(out, stride, w, h, (PIXEL)((sum + (len >> 1)) / len));
}
/* predict.rs:815-827: sums the whole `left` slice (its length is min(left_len, height)) */
static void SFX(pred_dc_left)(PIXEL *out, ptrdiff_t stride, const PIXEL *left, int left_n, int w, int h) {
  uint32_t sum = 0;
  for (int i = 0; i < left_n; i++) sum += left[i];
  SFX(fill)(out, stride, w, h, (PIXEL)((sum + (uint32_t)(h >> 1)) / (uint32_t)h));
}
/* predict.rs:829-840 */
static void SFX(pred_dc_top)(PIXEL *out, ptrdiff_t stride, const PIXEL *above, int w, int h) {
  uint32_t sum = 0;
  for (int i = 0; i < w; i++) sum += above[i];
  SFX(fill)(out, stride, w, h, (PIXEL)((sum + (uint32_t)(w >> 1)) / (uint32_t)w));
}
/* predict.rs:842-858 */
static void SFX(pred_h)(PIXEL *out, ptrdiff_t stride, const PIXEL *left, int w, int h) {
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) out[r * stride + c] = left[h - 1 - r];
}
static void SFX(pred_v)(PIXEL *out, ptrdiff_t stride, const PIXEL *above, int w, int h) {
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) out[r * stride + c] = above[c];
}
/* predict.rs:860-887 */
static void SFX(pred_paeth)(PIXEL *out, ptrdiff_t stride, const PIXEL *above, const PIXEL *left,
                            PIXEL above_left, int w, int h) {
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) {
      int32_t tl = above_left, l = left[h - 1 - r], t = above[c];
      int32_t p_base = t + l - tl;
      int32_t p_left = abs(p_base - l), p_top = abs(p_base - t), p_tl = abs(p_base - tl);
      if (p_left <= p_top && p_left <= p_tl)
        out[r * stride + c] = (PIXEL)l;
      else if (p_top <= p_tl)
        out[r * stride + c] = (PIXEL)t;
      else
        out[r * stride + c] = (PIXEL)tl;
    }
}
/* predict.rs:889-944 */
static void SFX(pred_smooth)(PIXEL *out, ptrdiff_t stride, const PIXEL *above, const PIXEL *left, int w, int h) {
  const uint32_t below_pred = left[0], right_pred = above[w - 1];
  const uint8_t *ww = sm_weight_arrays + w, *wh = sm_weight_arrays + h;
  const int log2_scale = 1 + 8;
  const uint32_t scale = 1u << 8;
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) {
      uint32_t p = wh[r] * (uint32_t)above[c] + (scale - wh[r]) * below_pred +
                   ww[c] * (uint32_t)left[h - 1 - r] + (scale - ww[c]) * right_pred;
      out[r * stride + c] = (PIXEL)((p + (1u << (log2_scale - 1))) >> log2_scale);
    }
}
/* predict.rs:946-981 */
static void SFX(pred_smooth_h)(PIXEL *out, ptrdiff_t stride, const PIXEL *above, const PIXEL *left, int w, int h) {
  const uint32_t right_pred = above[w - 1];
  const uint8_t *sw = sm_weight_arrays + w;
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) {
      uint32_t p = sw[c] * (uint32_t)left[h - 1 - r] + (256u - sw[c]) * right_pred;
      out[r * stride + c] = (PIXEL)((p + 128u) >> 8);
    }
}
/* predict.rs:983-1018 */
static void SFX(pred_smooth_v)(PIXEL *out, ptrdiff_t stride, const PIXEL *above, const PIXEL *left, int w, int h) {
  const uint32_t below_pred = left[0];
  const uint8_t *sw = sm_weight_arrays + h;
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) {
      uint32_t p = sw[r] * (uint32_t)above[c] + (256u - sw[r]) * below_pred;
      out[r * stride + c] = (PIXEL)((p + 128u) >> 8);
    }
}

/* predict.rs:1020-1063.  bw x bh = the chroma (plane) block size. */
static void SFX(pred_cfl_ac)(int16_t *ac, const PIXEL *luma, ptrdiff_t stride, int bw, int bh,
                             int w_pad, int h_pad, int xdec, int ydec) {
  const int max_luma_w = (bw - w_pad * 4) << xdec, max_luma_h = (bh - h_pad * 4) << ydec;
  const int max_luma_x = (max_luma_w > 8 ? max_luma_w : 8) - (1 << xdec);
  const int max_luma_y = (max_luma_h > 8 ? max_luma_h : 8) - (1 << ydec);
  int32_t sum = 0;
  for (int sub_y = 0; sub_y < bh; sub_y++)
    for (int sub_x = 0; sub_x < bw; sub_x++) {
      int luma_y = sub_y << ydec, luma_x = sub_x << xdec;
      int y = luma_y < max_luma_y ? luma_y : max_luma_y;
      int x = luma_x < max_luma_x ? luma_x : max_luma_x;
      int16_t sample = (int16_t)luma[y * stride + x];
      if (xdec) sample = (int16_t)(sample + (int16_t)luma[y * stride + x + 1]);
      if (ydec) sample = (int16_t)(sample + (int16_t)luma[(y + 1) * stride + x] + (int16_t)luma[(y + 1) * stride + x + 1]);
      sample = (int16_t)(sample << (3 - xdec - ydec));
      ac[sub_y * bw + sub_x] = sample;
      sum += sample;
    }
  int shift = 0;
  for (int v = bw; v > 1; v >>= 1) shift++;
  for (int v = bh; v > 1; v >>= 1) shift++;
  const int16_t average = (int16_t)((sum + (1 << (shift - 1))) >> shift);
  for (int i = 0; i < bw * bh; i++) ac[i] = (int16_t)(ac[i] - average);
}

/* predict.rs:1065-1088 */
static void SFX(pred_cfl_inner)(PIXEL *out, ptrdiff_t stride, const int16_t *ac, int16_t alpha, int w, int h, int bit_depth) {
  if (alpha == 0) return;
  const int32_t sample_max = (1 << bit_depth) - 1;
  const int32_t avg = out[0];
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) {
      int32_t v = avg + get_scaled_luma_q0(alpha, ac[r * w + c]);
      out[r * stride + c] = (PIXEL)(v < 0 ? 0 : v > sample_max ? sample_max : v);
    }
}

/* predict.rs:1203-1232 */
static void SFX(filter_edge)(int size, int strength, PIXEL *edge, int edge_len) {
  static const uint32_t K[3][5] = {{0, 4, 8, 4, 0}, {0, 5, 6, 5, 0}, {2, 4, 4, 4, 2}};
  if (strength == 0) return;
  PIXEL filtered[MAX_TX_SIZE * 4 + 1];
  memcpy(filtered, edge, (size_t)edge_len * sizeof(PIXEL));
  for (int i = 1; i < size; i++) {
    uint32_t s = 0;
    for (int j = 0; j < 5; j++) {
      int k = i + j - 2;
      if (k < 0) k = 0; /* saturating_sub */
      if (k > size - 1) k = size - 1;
      s += K[strength - 1][j] * (uint32_t)edge[k];
    }
    filtered[i] = (PIXEL)((s + 8) >> 4);
  }
  memcpy(edge, filtered, (size_t)edge_len * sizeof(PIXEL));
}

/* predict.rs:1234-1266 */
static void SFX(upsample_edge)(int size, PIXEL *edge, int bit_depth) {
  PIXEL dup[MAX_TX_SIZE];
  dup[0] = edge[0];
  for (int i = 0; i <= size; i++) dup[1 + i] = edge[i];
  dup[size + 2] = edge[size];
  edge[0] = dup[0];
  for (int i = 0; i < size; i++) {
    int32_t s = -(int32_t)dup[i] + 9 * (int32_t)dup[i + 1] + 9 * (int32_t)dup[i + 2] - (int32_t)dup[i + 3];
    s = (s + 8) / 16; /* Rust `/` truncates toward zero, as C */
    if (s < 0) s = 0;
    if (s > (1 << bit_depth) - 1) s = (1 << bit_depth) - 1;
    edge[2 * i + 1] = (PIXEL)s;
    edge[2 * i + 2] = dup[i + 2];
  }
}

static inline int32_t SFX(rs5)(int32_t v) { return (v + 16) >> 5; } /* round_shift(.., 5) */

/* predict.rs:1301-1505.  above: `above_n` pixels; left: `left_n` pixels bottom->top (the
 * left_and_left_below slice); ief: -1 None, else use_smooth_filter (0/1). */
static void SFX(pred_directional)(PIXEL *out, ptrdiff_t stride, const PIXEL *above, int above_n,
                                  const PIXEL *left, int left_n, PIXEL top_left, int p_angle, int width,
                                  int height, int bit_depth, int ief, int plane_w, int plane_h,
                                  int rect_x, int rect_y) {
  const int32_t sample_max = (1 << bit_depth) - 1;
  const int max_x = plane_w - 1, max_y = plane_h - 1;
  int upsample_above = 0, upsample_left = 0;
  const PIXEL *above_edge = above, *left_edge = left;
  int left_edge_len = left_n;
  const int enable_edge_filter = ief >= 0;
  const int flen = (width + height) * 2 + 1;
  PIXEL above_filtered[MAX_TX_SIZE * 4 + 1], left_filtered[MAX_TX_SIZE * 4 + 1];
  memset(above_filtered, 0, sizeof above_filtered);
  memset(left_filtered, 0, sizeof left_filtered);

  if (enable_edge_filter) {
    int above_len = above_n < flen - 1 ? above_n : flen - 1;
    int left_len = left_n < flen - 1 ? left_n : flen - 1;
    for (int i = 0; i < above_len; i++) above_filtered[1 + i] = above[i];
    for (int i = 1; i <= left_len; i++) left_filtered[i] = left[left_n - i];
    const int smooth_filter = ief;
    if (p_angle != 90 && p_angle != 180) {
      above_filtered[0] = top_left;
      left_filtered[0] = top_left;
      int aw = max_x - rect_x + 1, ah = max_y - rect_y + 1;
      int num_above = (width < aw ? width : aw) + (p_angle < 90 ? height : 0) + 1;
      int num_left = (height < ah ? height : ah) + (p_angle > 180 ? width : 0) + 1;
      SFX(filter_edge)(num_above, select_ief_strength(width, height, smooth_filter, p_angle - 90), above_filtered, flen);
      SFX(filter_edge)(num_left, select_ief_strength(width, height, smooth_filter, p_angle - 180), left_filtered, flen);
    }
    int num_above = width + (p_angle < 90 ? height : 0);
    int num_left = height + (p_angle > 180 ? width : 0);
    upsample_above = select_ief_upsample(width, height, smooth_filter, p_angle - 90);
    if (upsample_above) SFX(upsample_edge)(num_above, above_filtered, bit_depth);
    upsample_left = select_ief_upsample(width, height, smooth_filter, p_angle - 180);
    if (upsample_left) SFX(upsample_edge)(num_left, left_filtered, bit_depth);
    for (int i = 0; i < flen / 2; i++) { /* left_filtered.reverse() */
      PIXEL t = left_filtered[i];
      left_filtered[i] = left_filtered[flen - 1 - i];
      left_filtered[flen - 1 - i] = t;
    }
    above_edge = above_filtered;
    left_edge = left_filtered;
    left_edge_len = flen;
  }

  const int dx = p_angle < 90 ? dr_intra_derivative(p_angle)
                 : (p_angle > 90 && p_angle < 180) ? dr_intra_derivative(180 - p_angle) : 0;
  const int dy = (p_angle > 90 && p_angle < 180) ? dr_intra_derivative(p_angle - 90)
                 : p_angle > 180 ? dr_intra_derivative(270 - p_angle) : 0;
  const int offset_above = enable_edge_filter << upsample_above;
  const int offset_left = enable_edge_filter << upsample_left;

  if (p_angle < 90) {
    for (int i = 0; i < height; i++)
      for (int j = 0; j < width; j++) {
        int idx = (i + 1) * dx;
        int base = (idx >> (6 - upsample_above)) + (j << upsample_above);
        int32_t shift = ((idx << upsample_above) >> 1) & 31;
        int max_base_x = (height + width - 1) << upsample_above;
        int32_t v;
        if (base < max_base_x) {
          int32_t a = above_edge[base + offset_above], b = above_edge[base + 1 + offset_above];
          v = SFX(rs5)(a * (32 - shift) + b * shift);
        } else {
          v = above_edge[max_base_x + offset_above];
        }
        out[i * stride + j] = (PIXEL)(v < 0 ? 0 : v > sample_max ? sample_max : v);
      }
  } else if (p_angle > 90 && p_angle < 180) {
    for (int i = 0; i < height; i++)
      for (int j = 0; j < width; j++) {
        int idx = (j << 6) - (i + 1) * dx;
        int base = idx >> (6 - upsample_above); /* arithmetic shift on isize */
        int32_t v;
        if (base >= -(1 << upsample_above)) {
          int32_t shift = ((idx * (1 << upsample_above)) >> 1) & 31; /* isize arithmetic shift */
          int32_t a = (!enable_edge_filter && base < 0) ? (int32_t)top_left : (int32_t)above_edge[base + offset_above];
          int32_t b = above_edge[base + 1 + offset_above];
          v = SFX(rs5)(a * (32 - shift) + b * shift);
        } else {
          int idx2 = (i << 6) - (j + 1) * dy;
          int base2 = idx2 >> (6 - upsample_left);
          int32_t shift = ((idx2 * (1 << upsample_left)) >> 1) & 31;
          int l = left_edge_len - 1;
          int32_t a, b;
          if (!enable_edge_filter && base2 < 0)
            a = top_left;
          else if (base2 + offset_left == -2)
            a = left_edge[0];
          else
            a = left_edge[l - (base2 + offset_left)];
          if (base2 + offset_left == -2)
            b = left_edge[1];
          else
            b = left_edge[l - (base2 + offset_left + 1)];
          v = SFX(rs5)(a * (32 - shift) + b * shift);
        }
        out[i * stride + j] = (PIXEL)(v < 0 ? 0 : v > sample_max ? sample_max : v);
      }
  } else if (p_angle > 180) {
    for (int i = 0; i < height; i++)
      for (int j = 0; j < width; j++) {
        int idx = (j + 1) * dy;
        int base = (idx >> (6 - upsample_left)) + (i << upsample_left);
        int32_t shift = ((idx << upsample_left) >> 1) & 31;
        int l = left_edge_len - 1;
        int ia = l - (base + offset_left), ib = l - (base + offset_left + 1);
        if (ia < 0) ia = 0; /* saturating_sub */
        if (ib < 0) ib = 0;
        int32_t a = left_edge[ia], b = left_edge[ib];
        int32_t v = SFX(rs5)(a * (32 - shift) + b * shift);
        out[i * stride + j] = (PIXEL)(v < 0 ? 0 : v > sample_max ? sample_max : v);
      }
  }
}

/* predict.rs:705-784.  edge: the 4*64+1 element IntraEdge buffer, top-left at [128]. */
static void SFX(dispatch)(int mode, int variant, PIXEL *dst, ptrdiff_t stride, int width, int height,
                          int bit_depth, const int16_t *ac, int angle, int ief, const PIXEL *edge,
                          int left_len, int above_len, int plane_w, int plane_h, int dst_x, int dst_y) {
  const PIXEL *left_all = edge + 2 * MAX_TX_SIZE - left_len; /* bottom -> top */
  const PIXEL top_left = edge[2 * MAX_TX_SIZE];
  const PIXEL *above = edge + 2 * MAX_TX_SIZE + 1;
  const int ls_n = left_len < height ? left_len : height;           /* left_slice */
  const PIXEL *left_slice = left_all + (left_len - ls_n);
  const int llb_n = left_len < width + height ? left_len : width + height; /* left_and_left_below */
  const PIXEL *left_llb = left_all + (left_len - llb_n);
  switch (mode) {
    case DC_PRED:
      if (variant == VAR_NONE) SFX(fill)(dst, stride, width, height, (PIXEL)(128u << (bit_depth - 8)));
      else if (variant == VAR_LEFT) SFX(pred_dc_left)(dst, stride, left_slice, ls_n, width, height);
      else if (variant == VAR_TOP) SFX(pred_dc_top)(dst, stride, above, width, height);
      else SFX(pred_dc)(dst, stride, above, left_slice, width, height);
      break;
    case V_PRED: case H_PRED: case D45_PRED: case D135_PRED: case D113_PRED: case D157_PRED:
    case D203_PRED: case D67_PRED:
      if (mode == V_PRED && angle == 90) SFX(pred_v)(dst, stride, above, width, height);
      else if (mode == H_PRED && angle == 180) SFX(pred_h)(dst, stride, left_slice, width, height);
      else SFX(pred_directional)(dst, stride, above, above_len, left_llb, llb_n, top_left, angle, width,
                                 height, bit_depth, ief, plane_w, plane_h, dst_x, dst_y);
      break;
    case SMOOTH_PRED: SFX(pred_smooth)(dst, stride, above, left_slice, width, height); break;
    case SMOOTH_V_PRED: SFX(pred_smooth_v)(dst, stride, above, left_slice, width, height); break;
    case SMOOTH_H_PRED: SFX(pred_smooth_h)(dst, stride, above, left_slice, width, height); break;
    case PAETH_PRED: SFX(pred_paeth)(dst, stride, above, left_slice, top_left, width, height); break;
    case UV_CFL_PRED:
      if (variant == VAR_NONE) SFX(fill)(dst, stride, width, height, (PIXEL)(128u << (bit_depth - 8)));
      else if (variant == VAR_LEFT) SFX(pred_dc_left)(dst, stride, left_slice, ls_n, width, height);
      else if (variant == VAR_TOP) SFX(pred_dc_top)(dst, stride, above, width, height);
      else SFX(pred_dc)(dst, stride, above, left_slice, width, height);
      SFX(pred_cfl_inner)(dst, stride, ac, (int16_t)angle, width, height, bit_depth);
      break;
    default: break; /* unimplemented!() in the reference */
  }
}
