/*
 * oracle.h — CPU restatement of rav1e's RDO inner-loop kernels (TEST INFRASTRUCTURE ONLY).
 *
 * This library is the checker for the CUDA product path in rav1e_b200/csrc.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load it.  Nothing under rav1e_b200/ links, imports or calls it.
 *
 * Every function cites the reference file:line (relative to xiph/rav1e @ 564ae3b) that it
 * restates.  Arithmetic that lives in the off-disk crate v_frame 0.3.9 (Cargo.lock:2075)
 * is restated from its published semantics: msb(x)=31-clz(x), round_shift(v,b)=
 * (v+(1<<b>>1))>>b, ILog::ilog(x)=bits-clz(x) (0 for x<=0), clamp, `as` casts wrap.
 *
 * Parity pinning status (see DESIGN.md "Oracle"):
 *   SAD / SATD                : pinned by src/dist.rs:418-441, :477-500 (22 sizes, u8+u16)
 *   intra 4x4 predictors      : pinned by src/predict.rs:1514-1693
 *   cdef first_max_element    : pinned by src/cdef.rs:304-309
 *   fwd + inverse transform   : pinned jointly by the reference's round-trip test
 *                               (transform/mod.rs:479-617, 44 (size, type, tolerance) triples)
 *   mc / cdef dir + filter /  : no stored vectors in the reference (only asm==rust random
 *   intra edge filter,          tests that need rustc) -> "parity unpinned" upstream.  These
 *   upsampling, larger sizes    kernels are AV1-normative, so the AV1 specification is a second
 *                               source: independent restatements of its processes equal this
 *                               oracle bit for bit (tests/test_oracle_cdef_spec.py,
 *                               tests/test_oracle_predict_spec.py, the two-pass form in
 *                               tests/test_oracle_mc.py).
 *   get_intra_edges           : availability tables pinned by sha256 of the reference's 44 tables
 *   compute_rd_cost           : pinned (correctly rounded fma, exact rational check)
 *   quantize chain            : log_tx_scale / divu_pair KATs, sha256 of the 42 scan tables
 *   search stages, RDO dist   : likewise unpinned; checked against an independent Python model
 *                               (tests/test_oracle_search.py) and the reference tests' float
 *                               formulas (tests/test_oracle_rdo_dist.py).
 *
 * All strides are in ELEMENTS (pixels), pointers address pixel (0,0) of a region; planes
 * may be addressed at negative coordinates when the caller padded them.
 */
#ifndef RAV1E_B200_ORACLE_H
#define RAV1E_B200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ dist.rs */
uint32_t orc_get_sad_u8(const uint8_t *org, ptrdiff_t org_stride, const uint8_t *ref,
                        ptrdiff_t ref_stride, int w, int h);
uint32_t orc_get_sad_u16(const uint16_t *org, ptrdiff_t org_stride, const uint16_t *ref,
                         ptrdiff_t ref_stride, int w, int h);
uint32_t orc_get_satd_u8(const uint8_t *org, ptrdiff_t org_stride, const uint8_t *ref,
                         ptrdiff_t ref_stride, int w, int h);
uint32_t orc_get_satd_u16(const uint16_t *org, ptrdiff_t org_stride, const uint16_t *ref,
                          ptrdiff_t ref_stride, int w, int h);

/* -------------------------------------------------------------------- me.rs */
typedef struct {
  int16_t row, col; /* 1/8 pel, src/mc.rs:29-32 */
} orc_mv;

typedef struct {
  uint64_t cost; /* 256*sad + rate*lambda; UINT64_MAX == empty (src/me.rs:128-146) */
  uint32_t sad;
  orc_mv mv;
} orc_me_result;

uint32_t orc_get_mv_rate(orc_mv a, orc_mv b, int allow_high_precision_mv);
void orc_get_mv_range(int w_in_b, int h_in_b, int bo_x, int bo_y, int blk_w, int blk_h,
                      int *mvx_min, int *mvx_max, int *mvy_min, int *mvy_max);
uint64_t orc_mv_cost(uint32_t sad, orc_mv cand, orc_mv pmv0, orc_mv pmv1, uint32_t lambda,
                     int allow_high_precision_mv);

/* me.rs:1464-1509 — exhaustive window scan, first-min argmin.  bpp = 1|2 bytes/pixel.
 * org points at the block's pixel (0,0); ref0 at plane pixel (0,0); po = block position. */
orc_me_result orc_full_search(const void *org, ptrdiff_t org_stride, const void *ref0,
                              ptrdiff_t ref_stride, int bpp, int x_lo, int x_hi, int y_lo,
                              int y_hi, int w, int h, int po_x, int po_y, int step,
                              uint32_t lambda, orc_mv pmv0, orc_mv pmv1, int allow_hp);

/* Batched forms used by the parity tests and the CPU baseline (OpenMP over units).
 * Descriptors mirror include/b200rdo.h (b200_block / b200_cand) field for field. */
typedef struct {
  int16_t x, y; /* luma px of block top-left */
} orc_block;
typedef struct {
  uint32_t block;      /* index into blocks[] */
  int16_t mv_row, mv_col; /* 1/8 pel; fullpel offsets are mv/8 (trunc toward 0, me.rs:1402) */
} orc_cand;

/* me.rs:1386-1409 get_fullpel_mv_rd over a candidate list: out_sad[i] (UINT32_MAX when the
 * mv is out of range), out_cost[i] (UINT64_MAX likewise); either may be NULL. */
void orc_fullpel_candidates(const void *cur0, ptrdiff_t cur_stride, const void *ref0,
                            ptrdiff_t ref_stride, int bpp, int frame_w_in_b, int frame_h_in_b,
                            const orc_block *blocks, const orc_cand *cands, size_t n, int w,
                            int h, int use_satd, uint32_t lambda, const orc_mv *pmv /*2 per block or NULL*/,
                            int allow_hp, uint32_t *out_sad, uint64_t *out_cost, int threads);

/* get_subpel_mv_rd (me.rs:1411-1442) over a candidate list (sub-pel mvs): 8-tap MC + SAD/SATD. */
void orc_subpel_candidates(const void *cur0, ptrdiff_t cur_stride, const void *ref0,
                           ptrdiff_t ref_stride, int bpp, int frame_w_in_b, int frame_h_in_b,
                           const orc_block *blocks, const orc_cand *cands, size_t n, int w, int h,
                           int use_satd, uint32_t lambda, const orc_mv *pmv, int allow_hp,
                           int filter_mode, int bit_depth, uint32_t *out_sad, uint64_t *out_cost,
                           int threads);

/* full_pixel_me's final stage (me.rs:822-846) for every block: window +-range, step. */
void orc_full_search_blocks(const void *cur0, ptrdiff_t cur_stride, const void *ref0,
                            ptrdiff_t ref_stride, int bpp, int frame_w_in_b, int frame_h_in_b,
                            const orc_block *blocks, size_t nblocks, int w, int h, int range_x,
                            int range_y, int step, uint32_t lambda, int allow_hp,
                            orc_me_result *out, int threads);

/* full_pixel_me (me.rs:692-856) minus the final exhaustive grid: get_best_predictor +
 * fullpel_diamond_search per predictor subset, early exits, uneven_multi_hex_search (+
 * hexagon_search).  nsubsets 1 = non-extensive, 3 = extensive (median | subset_b | subset_c). */
void orc_full_pixel_me_blocks(const void *cur0, ptrdiff_t cur_stride, const void *ref0,
                              ptrdiff_t ref_stride, int bpp, int frame_w_in_b, int frame_h_in_b,
                              const orc_block *blocks, size_t nblocks, const orc_cand *preds,
                              const uint32_t *subset_offsets, int nsubsets, const orc_mv *pmv,
                              const uint32_t *thresh, int w, int h, uint32_t lambda, int allow_hp,
                              int umh_range, orc_me_result *out, int threads);

/* ------------------------------------------------- RDO distortion (dist.rs, activity.rs) */
/* get_weighted_sse dist.rs:234-283: one DistortionScale (Q14) per 4x4 chunk. */
uint64_t orc_weighted_sse_u8(const uint8_t *src1, ptrdiff_t stride1, const uint8_t *src2,
                             ptrdiff_t stride2, const uint32_t *scale, size_t scale_stride, int w, int h);
uint64_t orc_weighted_sse_u16(const uint16_t *src1, ptrdiff_t stride1, const uint16_t *src2,
                              ptrdiff_t stride2, const uint32_t *scale, size_t scale_stride, int w, int h);
/* cdef_dist_kernel dist.rs:302-372 (w, h <= 8); raw (may be NULL) receives {svar, dvar, sse},
 * what the asm kernels return (asm/x86/dist/cdef_dist.rs:18-24). */
uint32_t orc_cdef_dist_kernel_u8(const uint8_t *src, ptrdiff_t ss, const uint8_t *dst, ptrdiff_t ds,
                                 int w, int h, int bit_depth, uint32_t raw[3]);
uint32_t orc_cdef_dist_kernel_u16(const uint16_t *src, ptrdiff_t ss, const uint16_t *dst, ptrdiff_t ds,
                                  int w, int h, int bit_depth, uint32_t raw[3]);
uint32_t orc_apply_ssim_boost(uint32_t input, uint32_t svar, uint32_t dvar, int bit_depth);
uint64_t orc_distortion_scale_mul(uint32_t scale, uint64_t dist);
/* ActivityMask::from_plane + fill_scales (activity.rs:21-100) */
uint32_t orc_variance_8x8_u8(const uint8_t *src, ptrdiff_t stride);
uint32_t orc_variance_8x8_u16(const uint16_t *src, ptrdiff_t stride);
void orc_activity_mask(const void *luma, ptrdiff_t stride, int bpp, int width, int height, int bit_depth,
                       uint32_t *variances, uint32_t *scales);

/* -------------------------------------------------------------- transform/ */
/* forward.rs:71-161.  coeff_is_i32: 0 -> int16_t out (8-bit pixels), 1 -> int32_t out. */
int orc_valid_av1_transform(int tx_size, int tx_type);
void orc_forward_transform(const int16_t *input, void *output, size_t stride, int tx_size,
                           int tx_type, int bd, int coeff_is_i32);
void orc_forward_transform_batch(const int16_t *input, void *output, size_t nblocks, int tx_size,
                                 int tx_type, int bd, int coeff_is_i32, int threads);
int orc_tx_width(int tx_size);
int orc_tx_height(int tx_size);

/* ------------------------------------------------- transform/inverse.rs */
/* 1-D inverse transforms (kind = TxType1D: 0 DCT, 1 ADST, 2 FLIPADST, 3 IDTX, 4 WHT); returns 0
 * for the combinations the reference leaves unimplemented. */
int orc_inv_txfm_1d(int kind, int n, const int32_t *in, int32_t *out, int range);
/* rust::inverse_transform_add (inverse.rs:1637-1704): dst += inverse(input), clamped to bd bits. */
void orc_inverse_transform_add(const void *input, int coeff_is_i32, void *dst, ptrdiff_t dst_stride,
                               int bpp, int tx_size, int tx_type, int bd);

/* ------------------------------------------------- quantize/ (encoder.rs:1556-1655) */
int orc_get_log_tx_scale(int tx_size);
void orc_divu_gen(uint32_t d, uint32_t out[3]);
uint32_t orc_divu_pair(uint32_t x, const uint32_t d[3]);
int orc_coded_tx_area(int tx_size);
int orc_scan_kind(int tx_type); /* 0 default, 1 mrow, 2 mcol */
void orc_scan_order(int tx_size, int tx_type, uint16_t *scan, uint16_t *iscan /* may be NULL */);
/* quantize (quantize/mod.rs:269-361) -> dequantize (:368-392) -> raw tx-domain distortion
 * (encoder.rs:1611-1640) for nblocks blocks; dc_quant / ac_quant = dc_q() / ac_q() of the qindex. */
void orc_quantize_chain_batch(const void *coeffs, size_t nblocks, int tx_size, int tx_type,
                              uint32_t dc_quant, uint32_t ac_quant, int is_intra, int coeff_is_i32,
                              void *qcoeffs, void *rcoeffs, uint16_t *eob, uint64_t *tx_dist,
                              int threads);

/* ------------------------------------------------------------------- mc.rs */
/* put_8tap mc.rs:250-353; prep_8tap :360-451; mc_avg :454-479.  mode_x/mode_y = FilterMode
 * (0 REGULAR, 1 SMOOTH, 2 SHARP, 3 BILINEAR); col_frac/row_frac 0..15. */
void orc_put_8tap(void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride, int bpp,
                  int w, int h, int col_frac, int row_frac, int mode_x, int mode_y, int bit_depth);
void orc_prep_8tap(int16_t *tmp, const void *src, ptrdiff_t src_stride, int bpp, int w, int h,
                   int col_frac, int row_frac, int mode_x, int mode_y, int bit_depth);
void orc_mc_avg(void *dst, ptrdiff_t dst_stride, int bpp, const int16_t *tmp1, const int16_t *tmp2,
                int w, int h, int bit_depth);
void orc_get_filter(int mode, int frac, int length, int32_t out[8]);
void orc_get_mv_params(int mv_row, int mv_col, int xdec, int ydec, int *row_off, int *col_off,
                       int *row_frac, int *col_frac);
void orc_mc_blocks(const void *ref0, ptrdiff_t ref_stride, int bpp, const orc_block *blocks,
                   const orc_mv *mvs, size_t n, int w, int h, int mode_x, int mode_y, int bit_depth,
                   int xdec, int ydec, int kind, void *out, int threads);

/* -------------------------------------------------------------- predict.rs */
/* dispatch_predict_intra predict.rs:705-784.  `edge` is the reference's IntraEdge buffer
 * (partition.rs:600-637): 4*64+1 pixels, top-left at index 128, left[k] (bottom -> top) in
 * [128-left_len, 128), above in [129, 129+above_len).  mode: PredictionMode discriminant
 * (predict.rs:73-87, 0..13); variant: 0 NONE, 1 LEFT, 2 TOP, 3 BOTH; angle: the prediction angle
 * in degrees for directional modes, alpha for UV_CFL_PRED; ief: -1 = None, 0/1 =
 * Some(params) with use_smooth_filter() false/true; plane_w/h and dst_x/y: plane_cfg and
 * rect() of the destination (only read by the edge filter's num_px clipping). */
void orc_predict_intra(int mode, int variant, void *dst, ptrdiff_t dst_stride, int bpp, int w,
                       int h, int bit_depth, const int16_t *ac, int angle, int ief,
                       const void *edge, int left_len, int above_len, int plane_w, int plane_h,
                       int dst_x, int dst_y);
void orc_pred_cfl_ac(int16_t *ac, const void *luma, ptrdiff_t luma_stride, int bpp, int bw, int bh,
                     int w_pad, int h_pad, int xdec, int ydec);

/* ----------------------------------------------------------------- cdef.rs */
int orc_first_max_element(const int32_t *elems, int n, int32_t *max_out);
int orc_cdef_find_dir(const void *img, ptrdiff_t stride, int bpp, uint32_t *var, int coeff_shift);
/* asm-facing form: `in` is the padded u16 tile (asm/x86/cdef.rs:16-37) */
void orc_cdef_filter_block(void *dst, ptrdiff_t dst_stride, int dst_bpp, const uint16_t *in,
                           ptrdiff_t in_stride, int pri_strength, int sec_strength, int dir,
                           int damping, int bit_depth, int xdec, int ydec, int edges);
/* pixel-facing form (cdef.rs:198-231): pads into tmp16 according to `edges` first */
void orc_cdef_filter_block_px(void *dst, ptrdiff_t dst_stride, const void *in, ptrdiff_t istride,
                              int bpp, int pri_strength, int sec_strength, int dir, int damping,
                              int bit_depth, int xdec, int ydec, int edges);
int orc_cdef_adjust_strength(int strength, int var);
void orc_cdef_analyze_frame(const void *luma, ptrdiff_t stride, int bpp, int width, int height,
                            int bit_depth, const uint8_t *skip8, uint8_t *dir, int32_t *var);
void orc_cdef_filter_plane(const void *in, ptrdiff_t in_stride, void *out, ptrdiff_t out_stride,
                           int bpp, int plane, int xdec, int ydec, int width, int height,
                           int bit_depth, int damping, const uint8_t *skip8, const uint8_t *dir,
                           const int32_t *var, const uint8_t *strength_sb);

int orc_num_threads(void);

void orc_subpel_diamond_search_blocks(const void *cur0, ptrdiff_t cur_stride, const void *ref0, ptrdiff_t ref_stride,
                                      int bpp, int frame_w_in_b, int frame_h_in_b, const orc_block *blocks, size_t n,
                                      int w, int h, int use_satd, uint32_t lambda, const orc_mv *pmv, int allow_hp,
                                      int filter_mode, int bit_depth, orc_me_result *results, int threads);

/* ------------------------------------------- partition.rs / recon_intra.rs: intra edges */
int orc_block_size_index(int w, int h);
int orc_intra_avail_table(int kind, int bsize, uint8_t *out);
int orc_has_top_right(int bsize, int mi_col, int mi_row, int top_available, int right_available, int tx_w,
                      int row_off, int col_off, int ss_x, int ss_y);
int orc_has_bottom_left(int bsize, int mi_col, int mi_row, int bottom_available, int left_available, int tx_h,
                        int row_off, int col_off, int ss_x, int ss_y);
void orc_get_intra_edges(void *edge, const void *region, ptrdiff_t stride, int bpp, int plane_w, int plane_h,
                         int rect_x, int rect_y, int rect_w, int rect_h, int xdec, int ydec, int part_bo_x,
                         int part_bo_y, int bx, int by, int partition_bsize, int po_x, int po_y, int tx_w,
                         int tx_h, int bit_depth, int mode, int enable_intra_edge_filter, int angle_delta,
                         int *out_init_left, int *out_init_above);

/* ------------------------------------------------ api/lookahead.rs + v_frame Plane::downsampled */
void orc_plane_downsample(const void *src, ptrdiff_t src_stride, int src_w, int src_h, void *dst,
                          ptrdiff_t dst_stride, int dst_pad, int bpp, int pad_w, int pad_h);
void orc_estimate_intra_costs(const void *luma, ptrdiff_t stride, int width, int height, int bpp, int bit_depth,
                              uint32_t *costs);
double orc_importance_block_difference(const void *org, ptrdiff_t org_stride, const void *ref, ptrdiff_t ref_stride,
                                       int width, int height, int bpp);
double orc_estimate_inter_costs(const void *org, ptrdiff_t org_stride, const void *ref, ptrdiff_t ref_stride,
                                int width, int height, int bpp, const int16_t *mvs, uint32_t *costs);

/* -------------------------------------------------------------------- rdo.rs
 * compute_rd_cost (rdo.rs:718-723): lambda.mul_add(rate / 8.0, distortion as f64). */
double orc_compute_rd_cost(double lambda, uint32_t rate, uint64_t distortion);
void orc_compute_rd_cost_batch(double lambda, const uint32_t *rate, const uint64_t *distortion,
                               size_t n, double *out);

#ifdef __cplusplus
}
#endif
#endif
