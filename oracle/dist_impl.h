/* Included twice by dist.c with PIXEL = uint8_t / uint16_t.  Not a standalone header. */

/* dist.rs:31-52 rust::get_sad — sum |org-ref| over w x h. */
uint32_t SFX(orc_get_sad)(const PIXEL *org, ptrdiff_t org_stride, const PIXEL *ref,
                          ptrdiff_t ref_stride, int w, int h) {
  uint32_t sum = 0;
  for (int y = 0; y < h; y++) {
    const PIXEL *po = org + (ptrdiff_t)y * org_stride;
    const PIXEL *pr = ref + (ptrdiff_t)y * ref_stride;
    uint32_t row = 0;
    for (int x = 0; x < w; x++) {
      int32_t d = (int32_t)po[x] - (int32_t)pr[x];
      row += (uint32_t)(d < 0 ? -d : d);
    }
    sum += row;
  }
  return sum;
}

/* dist.rs:156-221 rust::get_satd — tile into size x size chunks (size = min(w,h,8)),
 * Hadamard each full chunk, SAD the partial ones, one rounding shift at the end. */
uint32_t SFX(orc_get_satd)(const PIXEL *org, ptrdiff_t org_stride, const PIXEL *ref,
                           ptrdiff_t ref_stride, int w, int h) {
  int size = w < h ? w : h;
  if (size > 8) size = 8; /* dist.rs:166: w.min(h).min(8) */
  /* dist.rs:167: tx2d = if size == 4 { hadamard4x4 } else { hadamard8x8 } */
  uint64_t sum = 0;
  for (int chunk_y = 0; chunk_y < h; chunk_y += size) {
    int chunk_h = h - chunk_y < size ? h - chunk_y : size;
    for (int chunk_x = 0; chunk_x < w; chunk_x += size) {
      int chunk_w = w - chunk_x < size ? w - chunk_x : size;
      const PIXEL *co = org + (ptrdiff_t)chunk_y * org_stride + chunk_x;
      const PIXEL *cr = ref + (ptrdiff_t)chunk_y * ref_stride + chunk_x;
      if (chunk_w != size || chunk_h != size) { /* dist.rs:185-191 */
        sum += SFX(orc_get_sad)(co, org_stride, cr, ref_stride, chunk_w, chunk_h);
        continue;
      }
      int32_t buf[64];
      for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++)
          buf[y * size + x] = (int32_t)co[(ptrdiff_t)y * org_stride + x] -
                              (int32_t)cr[(ptrdiff_t)y * ref_stride + x];
      /* size is 4 or 8 for every block rav1e produces; like the reference, anything
       * that is not 4 takes the 8x8 transform (and would read 64 entries). */
      if (size == 4)
        hadamard2d(buf, 4, 4);
      else
        hadamard2d(buf, 8, 8);
      for (int i = 0; i < size * size; i++)
        sum += (uint64_t)(buf[i] < 0 ? -(int64_t)buf[i] : (int64_t)buf[i]);
    }
  }
  int ln = orc_msb(size);
  return (uint32_t)((sum + ((1u << ln) >> 1)) >> ln);
}
