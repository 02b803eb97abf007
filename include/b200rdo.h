/*
 * b200rdo.h — C ABI of the B200-native backend for rav1e's RDO inner loop.
 *
 * Drop-in boundary: rav1e selects a kernel backend per file with a `cfg_if!` module swap
 * (src/dist.rs:10-18, src/mc.rs:10-18, src/predict.rs:16-24, src/cdef.rs:20-28,
 * src/transform/forward.rs:15-23) and, inside the wrapper module, a per-CpuFeatureLevel
 * table of `unsafe extern fn` pointers (src/cpu_features/x86.rs:97-158,
 * src/asm/x86/dist/mod.rs:483-729).  This header declares what such a wrapper module
 * (`asm::cuda::*`) binds:
 *
 *   (1) per-call entry points with EXACTLY the signatures of the x86 asm symbols
 *       (`rav1e_sad16x16_avx2(src, src_stride_bytes, dst, dst_stride_bytes) -> u32` ...),
 *       suffix `_cuda`, host pointers in / scalar out — so they slot into the existing
 *       SAD_FNS/SATD_FNS/... tables unchanged; and
 *   (2) batched entry points (`*_batch` = host buffers, `*_dev` = device-resident) that
 *       evaluate thousands of candidate blocks per launch — the form the north star asks
 *       for, which the reference's one-block-per-call ABI cannot express.
 *
 * Conventions (same as the asm ABI, src/asm/x86/dist/mod.rs:130-133): per-call entry
 * points take BYTE strides (ptrdiff_t); batched entry points take ELEMENT strides inside
 * b200_plane.  No torch types.  Every function returning int returns a b200_status;
 * b200_last_error() gives the message.  No entry point ever falls back to CPU code: if no
 * CUDA device is usable the call fails with B200_ERR_NODEV (per-call scalar forms abort,
 * mirroring the reference where a bad table entry is UB, not a silent wrong answer).
 */
#ifndef B200RDO_H
#define B200RDO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RDO_ABI_VERSION 1

typedef enum {
  B200_OK = 0,
  B200_ERR_CUDA = 1,  /* a CUDA runtime call failed */
  B200_ERR_ARG = 2,   /* precondition violated (the reference would assert!/panic) */
  B200_ERR_NODEV = 3, /* no usable sm_100 device */
  B200_ERR_OOM = 4
} b200_status;

typedef struct b200_ctx b200_ctx; /* one per host thread / rayon worker (encoder.rs:3253) */

int b200_abi_version(void);
int b200_device_count(void);
int b200_ctx_create(int device, b200_ctx **out);
void b200_ctx_destroy(b200_ctx *ctx);
const char *b200_last_error(const b200_ctx *ctx); /* ctx may be NULL: global last error */
/* Enqueue on an externally owned cudaStream_t (e.g. the host framework's current stream;
 * NULL = the CUDA default stream).  b200_ctx_reset_stream returns to the ctx's own stream.
 * ORDERING CONTRACT: every `*_dev` entry point only enqueues work on the ctx's stream.  The ctx's
 * own stream is created cudaStreamNonBlocking, i.e. it does NOT synchronise with the legacy default
 * stream: device buffers a caller fills on another stream (a framework's zero-fill, an upload) must
 * be complete - or that stream must be the one given to b200_ctx_set_stream - before they are
 * handed to a `*_dev` call, and results must be read after b200_ctx_synchronize() or on the same
 * stream. */
int b200_ctx_set_stream(b200_ctx *ctx, void *cuda_stream);
int b200_ctx_reset_stream(b200_ctx *ctx);
void *b200_ctx_get_stream(b200_ctx *ctx);
int b200_ctx_synchronize(b200_ctx *ctx);
/* Asynchronous host-buffer mode: when enabled, the `*_batch` / `*_resident` entry points enqueue
 * their copies and kernels and return without waiting; output buffers are valid after
 * b200_ctx_synchronize().  Host input/output buffers must stay alive (and should be pinned) until
 * then.  Lets a caller pipeline frames over several contexts so PCIe copies overlap kernels. */
int b200_ctx_set_async(b200_ctx *ctx, int enable);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
uint64_t b200_ctx_launch_count(const b200_ctx *ctx);

/* Raw device memory helpers so non-torch hosts (the Rust shim) can keep data resident. */
int b200_malloc(b200_ctx *ctx, size_t bytes, void **dptr);
int b200_free(b200_ctx *ctx, void *dptr);
int b200_memcpy_h2d(b200_ctx *ctx, void *dptr, const void *host, size_t bytes);
int b200_memcpy_d2h(b200_ctx *ctx, void *host, const void *dptr, size_t bytes);

/* ------------------------------------------------------------------ planes
 * Device image of v_frame::Plane<T> as seen through PlaneRegion (tiling/plane_region.rs:
 * 116-135): `data` addresses pixel (0,0); [-pad, width+pad) x [-pad, height+pad) is
 * readable (frame/mod.rs:22-23 pads 88 luma px; regions may address the padding,
 * plane_region.rs:164-172). */
typedef struct {
  void *data;     /* DEVICE pointer to pixel (0,0) */
  int32_t stride; /* elements */
  int32_t width, height;
  int32_t pad; /* readable border on all four sides, pixels */
  int32_t bpp; /* bytes per pixel: 1 (u8) or 2 (u16) */
  void *alloc; /* allocation base (owned by b200_plane_alloc) */
} b200_plane;

int b200_plane_alloc(b200_ctx *ctx, int width, int height, int pad, int bpp, b200_plane *out);
int b200_plane_free(b200_ctx *ctx, b200_plane *p);
/* Upload the visible width x height area from host (row pitch host_stride_bytes) and
 * replicate edges into the padding on-device (v_frame Plane::pad semantics). */
int b200_plane_upload(b200_ctx *ctx, const b200_plane *p, const void *host,
                      ptrdiff_t host_stride_bytes);
int b200_plane_download(b200_ctx *ctx, const b200_plane *p, void *host,
                        ptrdiff_t host_stride_bytes);

/* --------------------------------------------------- SAD / SATD / ME (dist.rs, me.rs)
 * Per-call, reference-signature entry points.  Replaces rav1e_sad{W}x{H}_{sse2,avx2},
 * rav1e_satd_{W}x{H}_{ssse3,sse4,avx2} and the _hbd variants (asm/x86/dist/mod.rs:21-43,
 * tables :483-729).  22 block sizes each: see B200_FOR_EACH_BLOCK_SIZE below. */
#define B200_FOR_EACH_BLOCK_SIZE(X)                                                      \
  X(4, 4) X(4, 8) X(4, 16) X(8, 4) X(8, 8) X(8, 16) X(8, 32) X(16, 4) X(16, 8) X(16, 16) \
  X(16, 32) X(16, 64) X(32, 8) X(32, 16) X(32, 32) X(32, 64) X(64, 16) X(64, 32)         \
  X(64, 64) X(64, 128) X(128, 64) X(128, 128)

#define B200_DECL_DIST(W, H)                                                              \
  uint32_t rav1e_sad##W##x##H##_cuda(const uint8_t *src, ptrdiff_t src_stride,            \
                                     const uint8_t *dst, ptrdiff_t dst_stride);           \
  uint32_t rav1e_sad_##W##x##H##_hbd_cuda(const uint16_t *src, ptrdiff_t src_stride,      \
                                          const uint16_t *dst, ptrdiff_t dst_stride);     \
  uint32_t rav1e_satd_##W##x##H##_cuda(const uint8_t *src, ptrdiff_t src_stride,          \
                                       const uint8_t *dst, ptrdiff_t dst_stride);         \
  uint32_t rav1e_satd_##W##x##H##_hbd_cuda(const uint16_t *src, ptrdiff_t src_stride,     \
                                           const uint16_t *dst, ptrdiff_t dst_stride,     \
                                           uint32_t bdmax);
B200_FOR_EACH_BLOCK_SIZE(B200_DECL_DIST)
#undef B200_DECL_DIST

/* Generic w x h (the `rust::get_sad` fallback for non-canonical crops,
 * asm/x86/dist/mod.rs:299) — host pointers, byte strides, bpp 1|2. */
uint32_t b200_get_sad(const void *org, ptrdiff_t org_stride, const void *ref,
                      ptrdiff_t ref_stride, int w, int h, int bpp);
uint32_t b200_get_satd(const void *org, ptrdiff_t org_stride, const void *ref,
                       ptrdiff_t ref_stride, int w, int h, int bpp);

/* Batched motion-estimation distortion.  One launch = all candidates of all blocks. */
typedef struct {
  int16_t x, y; /* luma px of the block's top-left (PlaneBlockOffset << 2) */
} b200_block;

typedef struct {
  uint32_t block;         /* index into blocks[] */
  int16_t mv_row, mv_col; /* MotionVector, 1/8 pel (mc.rs:29-32); fullpel offset = mv/8 */
} b200_cand;

typedef struct {
  uint64_t cost; /* 256*sad + rate*lambda (me.rs:1460); UINT64_MAX = empty (me.rs:139-146) */
  uint32_t sad;
  int16_t mv_row, mv_col;
} b200_me_result;

typedef struct {
  int32_t w, h;                       /* block size in px (<= 128) */
  int32_t frame_w_in_b, frame_h_in_b; /* fi.w_in_b / fi.h_in_b (4x4 units), for get_mv_range */
  uint32_t lambda;                    /* me.rs:549-552 */
  int32_t allow_high_precision_mv;
  int32_t use_satd; /* 0: get_sad, 1: get_satd (me.rs:1450-1454) */
  int32_t bit_depth;
  int32_t window_hint_px; /* 0 = unknown; else bound on |mv|/8 of the candidates, used only to
                             size the shared-memory window (wrong hints cost speed, not results) */
} b200_me_params;

/* get_fullpel_mv_rd (me.rs:1386-1409) over a candidate list, device-resident.
 * d_pmv: NULL (both predictors zero) or 2 MotionVectors (row,col int16) per block.
 * d_cand_offsets: NULL, or nblocks+1 CSR offsets when cands are grouped by block in
 *   ascending block order — required for d_best (per-block first-minimum winner in list
 *   order, the serial scan's tie-break, me.rs:898,974).  A block's list holds fewer than 2^24
 *   candidates (the winner's index inside its block is packed into 24 bits).
 * Outputs (each may be NULL): d_sad[ncands], d_cost[ncands], d_best[nblocks].
 * Padding contract (the reference's, frame/mod.rs:22-23 + me.rs:339-362): candidates are accepted
 * up to get_mv_range's border, 16 + block size pixels outside the frame, so both planes must be
 * readable that far (b200_plane.pad); nothing checks it per candidate. */
int b200_me_candidates_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                           const b200_block *d_blocks, size_t nblocks, const b200_cand *d_cands,
                           size_t ncands, const uint32_t *d_cand_offsets, const int16_t *d_pmv,
                           const b200_me_params *params, uint32_t *d_sad, uint64_t *d_cost,
                           b200_me_result *d_best);

/* The same over `npairs` (cur, ref) plane pairs in ONE launch — the reference evaluates every
 * allowed reference frame of every superblock in turn (estimate_tile_motion, me.rs:178-212) and
 * runs tiles and frames concurrently; a 30 us launch per pair leaves a B200 mostly ramping up and
 * draining, so the pairs share a grid.  Blocks, candidates, offsets, pmv and the outputs are the
 * concatenation pair after pair: pair k owns blocks [pair_block_end[k-1], pair_block_end[k]) and
 * candidates [pair_cand_end[k-1], pair_cand_end[k]) (HOST arrays, npairs entries; the last entries
 * equal nblocks / ncands); b200_cand.block and d_cand_offsets index the concatenated arrays.
 * curs/refs: HOST arrays of npairs plane descriptors sharing bpp and the frame size in `params`.
 * Results are identical to npairs calls of b200_me_candidates_dev. */
int b200_me_candidates_multi_dev(b200_ctx *ctx, size_t npairs, const b200_plane *curs,
                                 const b200_plane *refs, const uint32_t *pair_block_end,
                                 const uint32_t *pair_cand_end, const b200_block *d_blocks,
                                 size_t nblocks, const b200_cand *d_cands, size_t ncands,
                                 const uint32_t *d_cand_offsets, const int16_t *d_pmv,
                                 const b200_me_params *params, uint32_t *d_sad, uint64_t *d_cost,
                                 b200_me_result *d_best);

/* The move-to-best search stages of full_pixel_me (me.rs:692-856) for every block, on the device:
 * per predictor subset get_best_predictor (:884-909) + fullpel_diamond_search (:955-998), the
 * extensive ladder's early exits (`best.rd.sad < thresh`, :777-790) and uneven_multi_hex_search
 * (:1170-1303, ends in hexagon_search :1055-1135).  The exhaustive grid that may follow
 * (:822-846) is b200_me_full_search_dev.
 *   nsubsets 1: non-extensive search, d_preds holds `subsets.all_mvs()` per block (:757-759);
 *   nsubsets 3: extensive search, subsets = {median (skipped when empty), subset_b, subset_c},
 *               d_thresh[nblocks] = the caller's `thresh` (:771-772), umh_range = 24 (:806) or 0
 *               to stop before the UMH stage.
 * d_subset_offsets: nblocks * nsubsets + 1 offsets into d_preds (b200_cand.block is ignored).
 * Predictor gathering (get_subset_predictors, :381-533) stays with the caller: it reads the
 * neighbouring blocks' results of the same pass.  A block whose candidates are all out of range
 * (where the reference would assert) yields the empty result. */
int b200_me_search_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                       const b200_block *d_blocks, size_t nblocks, const b200_cand *d_preds,
                       const uint32_t *d_subset_offsets, int nsubsets, const int16_t *d_pmv,
                       const uint32_t *d_thresh, const b200_me_params *params, int umh_range,
                       b200_me_result *d_best);
/* The same over `npairs` plane pairs in one launch (block ranges as in
 * b200_me_candidates_multi_dev; offsets, pmv, thresh and results follow the concatenated blocks). */
int b200_me_search_multi_dev(b200_ctx *ctx, size_t npairs, const b200_plane *curs,
                             const b200_plane *refs, const uint32_t *pair_block_end,
                             const b200_block *d_blocks, size_t nblocks, const b200_cand *d_preds,
                             const uint32_t *d_subset_offsets, int nsubsets, const int16_t *d_pmv,
                             const uint32_t *d_thresh, const b200_me_params *params, int umh_range,
                             b200_me_result *d_best);

/* get_subpel_mv_rd (me.rs:1411-1442) over a candidate list of SUB-PEL vectors: each candidate is
 * predicted with the 8-tap filter `filter_mode` (fi.default_filter; predict_inter_single,
 * predict.rs:304-336) and measured with SAD or SATD against the source block — the unit of work
 * of subpel_diamond_search (me.rs:1311-1383).  Same outputs and tie-break as the full-pel form. */
int b200_me_subpel_candidates_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                                  const b200_block *d_blocks, size_t nblocks,
                                  const b200_cand *d_cands, size_t ncands,
                                  const uint32_t *d_cand_offsets, const int16_t *d_pmv,
                                  const b200_me_params *params, int filter_mode, uint32_t *d_sad,
                                  uint64_t *d_cost, b200_me_result *d_best);

/* Sub-pel refinement FUSED with the winner's residual and forward transform (BASELINE configs[3]): per
 * block, every candidate of its list is predicted (8-tap `filter_mode`), measured (SAD / SATD + mv cost)
 * and the first minimum kept, exactly like b200_me_subpel_candidates_dev; then the winner's prediction -
 * still in shared memory - is subtracted from the source and run through the forward transform
 * (tx_size / tx_type; block size = transform size; tx_size < 0: no transform), like
 * b200_mc_blocks_dev + b200_fwd_txfm_pred_dev would, without any prediction touching HBM.
 * 8x8, 16x16 and 32x32 blocks; d_cand_offsets required.  d_coeffs: nblocks x w*h (i16 for 8-bit planes,
 * i32 for HBD); blocks whose candidates are all out of range get zeros. */
int b200_subpel_rdo_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref, const b200_block *d_blocks,
                        size_t nblocks, const b200_cand *d_cands, size_t ncands, const uint32_t *d_cand_offsets,
                        const int16_t *d_pmv, const b200_me_params *params, int filter_mode, int tx_size,
                        int tx_type, uint32_t *d_sad, uint64_t *d_cost, b200_me_result *d_best, void *d_coeffs);

/* subpel_diamond_search (me.rs:1311-1383) for every block, the move-to-best loop on the device: starting
 * from d_start[i] (the full-pel stage's MotionSearchResult: vector, cost, sad - blocks whose start is the
 * empty result are undefined upstream, `assert!(!current.is_empty())`), the four diamond points at 1/2 pel
 * are evaluated with get_subpel_mv_rd (8-tap prediction + SAD / SATD + mv cost), the centre moves while a
 * point is strictly better, then the radius halves down to 1/4 pel (1/8 with allow_high_precision_mv).
 * d_best[i] = the final result (d_start may equal d_best); with tx_size >= 0 the final vector's prediction
 * is subtracted from the source and transformed like b200_subpel_rdo_dev does.  Same block sizes. */
int b200_subpel_search_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref, const b200_block *d_blocks,
                           size_t nblocks, const b200_me_result *d_start, const int16_t *d_pmv,
                           const b200_me_params *params, int filter_mode, int tx_size, int tx_type,
                           b200_me_result *d_best, void *d_coeffs);

/* full_search (me.rs:1464-1509) as called from full_pixel_me (me.rs:822-846) for every
 * block: window po +- (range_x, range_y) px clamped to get_mv_range, positions every
 * `step` px, pmv = 0, first-minimum winner in row-major scan order. */
int b200_me_full_search_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                            const b200_block *d_blocks, size_t nblocks,
                            const b200_me_params *params, int range_x, int range_y, int step,
                            b200_me_result *d_best);

/* Residual of every block against the reference displaced by the full-pel part of the
 * winners' motion vectors (the `diff` of encode_tx_block, encoder.rs:1533); d_mv_src may be
 * NULL (zero motion).  d_out: nblocks packed w x h int16 blocks, the input layout of
 * b200_fwd_txfm_dev with in_block_stride = w*h, in_row_stride = w. */
int b200_block_residual_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                            const b200_block *d_blocks, size_t nblocks,
                            const b200_me_result *d_mv_src, int w, int h, int16_t *d_out);

/* Host-buffer forms (what a Rust caller holding Plane<T> memory calls): planes are given
 * as host pointers to pixel (0,0) + byte strides and must be readable over the padding the
 * candidates can reach; everything is copied H2D, computed, and copied back before return. */
typedef struct {
  const void *data;    /* HOST pointer to pixel (0,0) */
  ptrdiff_t stride;    /* bytes */
  int32_t width, height;
  int32_t pad;         /* readable border present in host memory, pixels */
  int32_t bpp;
} b200_host_plane;

int b200_me_candidates_batch(b200_ctx *ctx, const b200_host_plane *cur,
                             const b200_host_plane *ref, const b200_block *blocks,
                             size_t nblocks, const b200_cand *cands, size_t ncands,
                             const uint32_t *cand_offsets, const int16_t *pmv,
                             const b200_me_params *params, uint32_t *sad, uint64_t *cost,
                             b200_me_result *best);
/* Same as b200_me_candidates_batch but with the planes already resident on the device
 * (b200_plane_alloc + b200_plane_upload once per frame, reused by every call of the frame): only
 * descriptors travel host->device and results device->host. */
int b200_me_candidates_resident(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                                const b200_block *blocks, size_t nblocks, const b200_cand *cands,
                                size_t ncands, const uint32_t *cand_offsets, const int16_t *pmv,
                                const b200_me_params *params, uint32_t *sad, uint64_t *cost,
                                b200_me_result *best);
/* The same with the candidates as plain MotionVector lists: mvs[2*i] = row, mvs[2*i+1] = col of
 * candidate i (1/8 pel), the lists of consecutive blocks concatenated and delimited by
 * cand_offsets (required) — the `&[MotionVector]` a search stage holds per block, 4 bytes per
 * candidate over PCIe instead of 8. */
int b200_me_mvs_resident(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                         const b200_block *blocks, size_t nblocks, const int16_t *mvs, size_t ncands,
                         const uint32_t *cand_offsets, const int16_t *pmv,
                         const b200_me_params *params, uint32_t *sad, uint64_t *cost,
                         b200_me_result *best);
int b200_me_full_search_batch(b200_ctx *ctx, const b200_host_plane *cur,
                              const b200_host_plane *ref, const b200_block *blocks,
                              size_t nblocks, const b200_me_params *params, int range_x,
                              int range_y, int step, b200_me_result *best);

/* ---------------------------------------------------------------- RDO distortion kernels
 * get_weighted_sse (dist.rs:234-283; asm `rav1e_weighted_sse_{W}x{H}`, asm/x86/dist/sse.rs:18-35)
 * and cdef_dist_kernel (dist.rs:302-372; asm `rav1e_cdef_dist_kernel_{W}x{H}`,
 * asm/x86/dist/cdef_dist.rs:18-52) — the distortion terms of compute_distortion / sse_wxh /
 * cdef_dist_wxh (rdo.rs:142-330).
 *
 * Per-call forms: HOST pointers, BYTE strides, like the asm symbols.  `scale`: one Q14
 * DistortionScale per 4x4 chunk, rows `scale_stride_bytes` apart.  b200_cdef_dist_kernel returns
 * apply_ssim_boost(sse, svar, dvar) (activity.rs:159-186) and, when ret != NULL, the asm kernels'
 * raw triple {svar, dvar, sse}. */
uint64_t b200_weighted_sse(const void *src, ptrdiff_t src_stride, const void *dst, ptrdiff_t dst_stride,
                           const uint32_t *scale, ptrdiff_t scale_stride_bytes, int w, int h, int bpp);
uint32_t b200_cdef_dist_kernel(const void *src, ptrdiff_t src_stride, const void *dst, ptrdiff_t dst_stride,
                               int w, int h, int bit_depth, uint32_t ret[3]);
/* Batched, device-resident: block i = the w x h area at d_blocks[i] of both planes.
 * d_scale: one DistortionScale per 4x4 chunk of the PLANE (chunk (cx, cy) at
 * d_scale[cy * scale_stride + cx], entries); blocks sit on multiples of 4 pixels.
 * d_out[i] = get_weighted_sse of block i. */
int b200_weighted_sse_dev(b200_ctx *ctx, const b200_plane *src1, const b200_plane *src2,
                          const b200_block *d_blocks, size_t nblocks, int w, int h,
                          const uint32_t *d_scale, size_t scale_stride, uint64_t *d_out);
/* cdef_dist_kernel for many blocks (w, h <= 8; cdef_dist_wxh walks a block in 8x8 steps,
 * rdo.rs:152-170).  d_out[i] (may be NULL) = the boosted distortion, d_raw[3*i..] (may be NULL) =
 * {svar, dvar, sse}. */
int b200_cdef_dist_dev(b200_ctx *ctx, const b200_plane *src, const b200_plane *dst,
                       const b200_block *d_blocks, size_t nblocks, int w, int h, int bit_depth,
                       uint32_t *d_out, uint32_t *d_raw);

/* ---------------------------------------------------------------- RDO cost (rdo.rs)
 * compute_rd_cost (rdo.rs:718-723): fi.lambda.mul_add(rate as f64 / 8.0, distortion.0 as f64) - the
 * only floating-point value on the path.  Device DFMA + round-to-nearest conversions: bit-identical
 * to f64::mul_add (0 ULP; the north star allows 1).
 * Batched: d_cost[i] (may be NULL when d_best is asked for) for n (rate, distortion) pairs; with
 * d_group_offsets (ngroups + 1 CSR offsets) d_best[g] = index inside group g of its first minimum
 * (the RDO loops' strict `rd < best.rd_cost`, e.g. rdo.rs:1003-1009), 0xffffffff for an empty group. */
int b200_compute_rd_cost_dev(b200_ctx *ctx, double lambda, const uint32_t *d_rate,
                             const uint64_t *d_distortion, size_t n, const uint32_t *d_group_offsets,
                             size_t ngroups, double *d_cost, uint32_t *d_best);
/* Per-call form (host scalars in, cost out; one launch on the calling thread's context). */
double b200_compute_rd_cost(double lambda, uint32_t rate, uint64_t distortion);

/* ------------------------------------------------ forward transform (transform/forward.rs)
 * The reference has no extern-C symbol here: the boundary is the generic Rust fn
 *   forward_transform<T: Coefficient>(input: &[i16], output: &mut [MaybeUninit<T>],
 *                                     stride, tx_size: TxSize, tx_type: TxType, bd, cpu)
 * (asm/x86/transform/forward.rs:444-447).  tx_size / tx_type are the enum discriminants
 * (transform/mod.rs:56-74 and :101-123); coeff_is_i32 selects T::Coeff (0: i16 for 8-bit
 * pixels, 1: i32 for HBD).  Output order is the reference's: column-major within 32x32
 * chunks (forward.rs:135-159).  Invalid (size, type) pairs: the reference panics
 * (forward.rs:75); the status-returning forms return B200_ERR_ARG, the per-call form aborts. */
int b200_valid_av1_transform(int tx_size, int tx_type);
int b200_tx_width(int tx_size);
int b200_tx_height(int tx_size);
void b200_forward_transform(const int16_t *input, void *output, size_t stride, int tx_size,
                            int tx_type, int bd, int coeff_is_i32);
/* Batched: block i reads input[i*in_block_stride + r*in_row_stride + c] and writes
 * output[i*w*h ...].  `_dev`: device pointers, asynchronous on the ctx stream. */
int b200_fwd_txfm_dev(b200_ctx *ctx, const int16_t *d_input, size_t in_block_stride,
                      size_t in_row_stride, void *d_output, size_t nblocks, int tx_size,
                      int tx_type, int bd, int coeff_is_i32);
/* Fused residual (`diff`, encoder.rs:1533) + forward transform: block i = cur(block) minus ref
 * displaced by the full-pel part of d_mv_src[i] (NULL: zero motion).  u8 planes produce i16
 * coefficients, u16 planes i32; bd must match the planes' depth. */
int b200_fwd_txfm_residual_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                               const b200_block *d_blocks, size_t nblocks,
                               const b200_me_result *d_mv_src, void *d_output, int tx_size,
                               int tx_type, int bd);
/* The same over `npairs` plane pairs in one launch; pair k owns blocks
 * [pair_block_end[k-1], pair_block_end[k]) (HOST array) of the concatenated d_blocks / d_mv_src /
 * d_output, as in b200_me_candidates_multi_dev. */
int b200_fwd_txfm_residual_multi_dev(b200_ctx *ctx, size_t npairs, const b200_plane *curs,
                                     const b200_plane *refs, const uint32_t *pair_block_end,
                                     const b200_block *d_blocks, size_t nblocks,
                                     const b200_me_result *d_mv_src, void *d_output, int tx_size,
                                     int tx_type, int bd);
/* ActivityMask::from_plane + fill_scales (activity.rs:21-100): d_variances[by * wb + bx] =
 * variance_8x8 of luma block (bx, by), wb = ceil(width / 8), hb = ceil(height / 8) (the plane's
 * padding must cover the rounding up); d_scales (may be NULL) = ssim_boost(var, var, bit_depth),
 * the Q14 DistortionScale per importance block. */
int b200_activity_mask_dev(b200_ctx *ctx, const b200_plane *luma, int bit_depth, uint32_t *d_variances,
                           uint32_t *d_scales);

/* ---------------------------------------------------------------- frame pipe (host buffers, per frame)
 * One call per frame runs the ME and transform legs of the RDO inner loop for every block of the frame
 * with the state an encoder thread keeps per tile: pushing a frame uploads it ONCE (the previously
 * pushed frame is the reference, like the reference slots rav1e searches, me.rs:178-212), the block
 * grid and every intermediate stay on the device, candidate lists travel as 2 bytes per candidate
 * ((row, col) full-pel offsets around a per-block centre MotionVector: a search stage's pattern around
 * its predictor, me.rs:884-1303; candidate = centre + 8 * offset with i16 wrapping), the SAD winners
 * feed the fused residual + forward transform (+ quantize chain, when ac_quant != 0) on the device and
 * only results cross PCIe.  Asynchronous like the other host-buffer forms under b200_ctx_set_async:
 * host buffers (pinned for speed) are valid after b200_ctx_synchronize. */
typedef struct b200_frame_pipe b200_frame_pipe;
typedef struct {
  int32_t width, height, pad, bpp, bit_depth;
  int32_t block_w, block_h; /* the blocks tile the frame (full blocks only), row-major */
  uint32_t lambda;
  int32_t sad_per_block, satd_per_block; /* list lengths (uniform over the blocks; 0 = leg off) */
  int32_t window_hint_px;                /* bound on the offsets' reach, as in b200_me_params */
  int32_t tx_size, tx_type; /* transform of the SAD winner's residual (block size = transform size); tx_size < 0: off */
  uint32_t dc_quant, ac_quant; /* ac_quant 0: raw coefficients out; else qcoeffs + eob + tx-domain distortion */
} b200_frame_pipe_cfg;
int b200_frame_pipe_create(b200_ctx *ctx, const b200_frame_pipe_cfg *cfg, b200_frame_pipe **out);
void b200_frame_pipe_destroy(b200_frame_pipe *pipe);
size_t b200_frame_pipe_nblocks(const b200_frame_pipe *pipe);
/* frame: HOST pointer to the visible width x height area.  sad_offsets / satd_offsets: nblocks x
 * per_block (row, col) int8 pairs; centers: nblocks (row, col) int16 MotionVectors in 1/8 pel or NULL
 * (zero).  Outputs (each may be NULL): best_sad / best_satd [nblocks]; coeffs: nblocks x w*h raw
 * coefficients, or nblocks x b200_coded_tx_area() quantized ones with eob [nblocks] and tx_dist
 * [nblocks] (i16 for 8-bit frames, i32 for HBD).  The first push only uploads (there is no reference
 * yet) and leaves the outputs untouched.
 * Copies: inputs that lie back to back in host memory in the order [frame (dense rows) | sad_offsets |
 * satd_offsets | centers] travel as ONE host->device copy, outputs laid out [best_sad | best_satd |
 * coeffs | eob | (8-byte aligned) tx_dist] as ONE device->host copy; any other layout works with a copy
 * per buffer. */
int b200_frame_pipe_push(b200_frame_pipe *pipe, const void *frame, ptrdiff_t frame_stride_bytes,
                         const int8_t *sad_offsets, const int8_t *satd_offsets, const int16_t *centers,
                         b200_me_result *best_sad, b200_me_result *best_satd, void *coeffs, uint16_t *eob,
                         uint64_t *tx_dist);
/* Resident candidate lists.  A search stage's pattern is an encoder constant, not per-frame data: upload the
 * lists (same layout as for b200_frame_pipe_push; centers may be NULL) ONCE; the device keeps them expanded, and
 * later pushes may pass NULL for sad_offsets / satd_offsets / centers - then only the frame crosses PCIe.
 * Synchronous. */
int b200_frame_pipe_set_lists(b200_frame_pipe *pipe, const int8_t *sad_offsets, const int8_t *satd_offsets,
                              const int16_t *centers);
/* A push whose results come back PACKED: what the entropy coder reads from a transform block is
 * coeffs[scan[0 .. eob)] (src/context/block_unit.rs write_coeffs_lv_map walks the scan order up to eob), so
 * a quantizing pipe (tx_size >= 0, ac_quant != 0) with resident lists returns, per block, its eob and tx-domain
 * distortion and - block after block in `packed` - only those eob quantized coefficients in scan order
 * (i16 for 8-bit frames, i32 for HBD).  *packed_count = total number of coefficients; when it exceeds
 * packed_capacity (in coefficients) only the first packed_capacity were copied (the call still succeeds: the
 * caller retries with a larger buffer).  best_sad / best_satd / eob / tx_dist may be NULL.  The call waits
 * once for the device (the size of the packed copy is a result); in asynchronous mode the packed copy itself is
 * still in flight when it returns.  The first push of a pipe only uploads: *packed_count = 0. */
int b200_frame_pipe_push_packed(b200_frame_pipe *pipe, const void *frame, ptrdiff_t frame_stride_bytes,
                                b200_me_result *best_sad, b200_me_result *best_satd, uint16_t *eob,
                                uint64_t *tx_dist, void *packed, size_t packed_capacity, size_t *packed_count);

/* ---------------------------------------------------------------- lookahead (api/lookahead.rs)
 * Plane::downsampled (v_frame 0.3.9; encoder.rs:476-477 builds the half / quarter resolution planes of
 * the ME pyramid): dst(c, r) = (src(2c, 2r) + src(2c+1, 2r) + src(2c, 2r+1) + src(2c+1, 2r+1) + 2) >> 2,
 * dst is ((src.width + 1) / 2) x ((src.height + 1) / 2), followed by Plane::pad(frame_w, frame_h):
 * everything outside [0, pad_w) x [0, pad_h) is a replica of the nearest pixel inside, with
 * pad_w = (frame_w + xdec) >> xdec, pad_h = (frame_h + ydec) >> ydec of the NEW plane (xdec = 1 for
 * the half, 2 for the quarter resolution plane).  The whole padded area of dst is written. */
int b200_plane_downsample_dev(b200_ctx *ctx, const b200_plane *src, const b200_plane *dst, int pad_w, int pad_h);
/* estimate_intra_costs (lookahead.rs:30-128): for every 8x8 importance block of the source luma,
 * get_intra_edges(DC_PRED, TX_8X8) -> DC prediction -> get_satd against the source, fused;
 * d_costs[(y / 8) * (width / 8) + x / 8]. */
int b200_estimate_intra_costs_dev(b200_ctx *ctx, const b200_plane *luma, int bit_depth, uint32_t *d_costs);
/* The cost part of estimate_inter_costs (lookahead.rs:238-270): SATD between every importance block and
 * the reference block displaced by its motion vector (d_mvs: (row, col) int16 per importance block =
 * stats[y * 2][x * 2].mv of the reference, e.g. the winners of b200_me_search_dev) -> d_costs (may be
 * NULL) and *d_mean = sum / count as f64 (one correctly rounded division, bit-identical to the
 * reference's).  d_scratch: 8 bytes of device memory.  The reference plane must be readable wherever
 * the vectors point (padding). */
int b200_estimate_inter_costs_dev(b200_ctx *ctx, const b200_plane *org, const b200_plane *ref, const int16_t *d_mvs,
                                  uint32_t *d_costs, uint64_t *d_scratch, double *d_mean);
/* estimate_importance_block_difference (lookahead.rs:131-180): mean over the importance blocks of
 * |round(mean(org block)) - round(mean(ref block))| as f64. */
int b200_importance_block_difference_dev(b200_ctx *ctx, const b200_plane *org, const b200_plane *ref,
                                         uint64_t *d_scratch, double *d_mean);

/* ---------------------------------------------------------------- quantize chain
 * The steps of encode_tx_block after the forward transform (encoder.rs:1556-1655) for nblocks
 * transform blocks of one (tx_size, tx_type), device-resident:
 *   QuantizationContext::quantize (quantize/mod.rs:269-361)  -> d_qcoeffs, d_eob
 *   dequantize (quantize/mod.rs:368-392)                     -> d_rcoeffs   (may be NULL)
 *   raw transform-domain distortion (encoder.rs:1611-1640)   -> d_tx_dist   (may be NULL)
 * d_coeffs: nblocks x (w*h) coefficients exactly as b200_fwd_txfm_* writes them (i16, or i32 when
 * coeff_is_i32); d_qcoeffs / d_rcoeffs: nblocks x b200_coded_tx_area(tx_size) (64-point dimensions
 * code 32, av1_get_coded_tx_size).  dc_quant / ac_quant are the caller's dc_q() / ac_q() of the
 * block's qindex (quantize/mod.rs:36-48; the lookup is control plane).  tx_type 0..15 (WHT_WHT is
 * lossless-only and has no scan order).  The rate estimate and the bias multiplications that
 * follow (encoder.rs:1642-1652) consume d_tx_dist on the host. */
int b200_coded_tx_area(int tx_size);
int b200_quantize_dev(b200_ctx *ctx, const void *d_coeffs, size_t nblocks, int tx_size, int tx_type,
                      uint32_t dc_quant, uint32_t ac_quant, int is_intra, int coeff_is_i32,
                      void *d_qcoeffs, void *d_rcoeffs, uint16_t *d_eob, uint64_t *d_tx_dist);

/* inverse_transform_add (transform/inverse.rs:1637-1704) for nblocks blocks of one (tx_size,
 * tx_type): dst(block i) += inverse(coefficients i), clamped to bd bits — the reconstruction step
 * of encode_tx_block (encoder.rs:1600-1609).  d_coeffs: b200_coded_tx_area(tx_size) coefficients
 * per block in the forward transform's layout (i16 for 8-bit planes, i32 for HBD), e.g. the
 * d_rcoeffs of b200_quantize_dev.  Blocks of one call must not overlap.  tx_type 0..16; pairs the
 * reference leaves `unimplemented!()` (INV_TXFM_FNS :1593-1623) are rejected. */
int b200_inverse_transform_add_dev(b200_ctx *ctx, const void *d_coeffs, const b200_plane *dst,
                                   const b200_block *d_blocks, size_t nblocks, int tx_size, int tx_type,
                                   int bd);

/* encode_tx_block's numeric core (encoder.rs:1492-1655) for nblocks transform blocks in ONE kernel - the
 * coefficients of a block stay in shared memory from the forward transform to the reconstruction:
 * residual against `ref` displaced by d_mv_src (NULL: zero motion) -> forward transform -> quantize ->
 * dequantize -> raw tx-domain distortion, and, when need_recon_pixel, the inverse transform added into
 * `rec` (which must already hold the prediction in the blocks' areas).
 * d_qcoeffs: nblocks x b200_coded_tx_area(tx_size) (required; i16, or i32 for HBD).  Optional outputs (each
 * may be NULL): d_coeffs nblocks x w*h raw coefficients, d_rcoeffs dequantized ones (coded area), d_eob,
 * d_tx_dist.  Entropy coding of d_qcoeffs and the rate / bias terms stay with the caller.  Results are
 * identical to b200_fwd_txfm_residual_dev + b200_quantize_dev + b200_inverse_transform_add_dev. */
int b200_encode_tx_blocks_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                              const b200_plane *rec, const b200_block *d_blocks, size_t nblocks,
                              const b200_me_result *d_mv_src, int tx_size, int tx_type, int bd,
                              uint32_t dc_quant, uint32_t ac_quant, int is_intra, int need_recon_pixel,
                              void *d_coeffs, void *d_qcoeffs, void *d_rcoeffs, uint16_t *d_eob,
                              uint64_t *d_tx_dist);

/* Fused residual + transform with resident planes and HOST descriptors / outputs. */
int b200_fwd_txfm_residual_resident(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                                    const b200_block *blocks, size_t nblocks,
                                    const b200_me_result *mv_src, void *output, int tx_size,
                                    int tx_type, int bd);
/* Residual against PACKED predictions (nblocks x h x w pixels, the output of b200_mc_blocks_dev or
 * b200_predict_intra_dev) + forward transform: predict -> diff -> forward_transform of
 * encode_tx_block (encoder.rs:1492-1544) without the residual touching HBM. */
int b200_fwd_txfm_pred_dev(b200_ctx *ctx, const b200_plane *cur, const void *d_pred,
                           const b200_block *d_blocks, size_t nblocks, void *d_output, int tx_size,
                           int tx_type, int bd);
int b200_fwd_txfm_batch(b200_ctx *ctx, const int16_t *input, size_t in_block_stride,
                        size_t in_row_stride, void *output, size_t nblocks, int tx_size,
                        int tx_type, int bd, int coeff_is_i32);

/* ------------------------------------------------ motion compensation (mc.rs, predict.rs)
 * Per-call forms: the argument order of the dav1d-style symbols rav1e binds
 * (`rav1e_put_8tap_regular_smooth_8bpc_avx2(dst, dst_stride, src, src_stride, w, h, mx, my)`,
 * asm/x86/mc.rs:17-76, table index (fx + 4*fy)&15 :80-82) with the filter pair passed as
 * arguments instead of being baked into the symbol name.  Host pointers, BYTE strides; src must
 * be readable over [-3, w+4) x [-3, h+4) (asm/x86/mc.rs:122-123).  FilterMode: 0 REGULAR,
 * 1 SMOOTH, 2 SHARP, 3 BILINEAR (mc.rs:98-106).  Preconditions as the reference asserts
 * (mc.rs:256-257): even height, power-of-two width in 2..128; violations abort. */
void b200_put_8tap(void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride, int w,
                   int h, int col_frac, int row_frac, int mode_x, int mode_y, int bit_depth);
void b200_prep_8tap(int16_t *tmp, const void *src, ptrdiff_t src_stride, int w, int h,
                    int col_frac, int row_frac, int mode_x, int mode_y, int bit_depth);
void b200_mc_avg(void *dst, ptrdiff_t dst_stride, const int16_t *tmp1, const int16_t *tmp2, int w,
                 int h, int bit_depth);
/* Batched predict_inter_single (predict.rs:304-336): block i is predicted from `ref` at
 * d_blocks[i] displaced by d_mvs[2i] (row), d_mvs[2i+1] (col) in 1/8 pel through get_mv_params
 * (predict.rs:284-297; xdec/ydec = plane decimation).  kind 0 = put (packed w*h pixels per block),
 * kind 1 = prep (packed w*h int16 per block, for b200_mc_avg_dev). */
int b200_mc_blocks_dev(b200_ctx *ctx, const b200_plane *ref, const b200_block *d_blocks,
                       const int16_t *d_mvs, size_t nblocks, int w, int h, int mode_x, int mode_y,
                       int bit_depth, int xdec, int ydec, int kind, void *d_out);
int b200_mc_avg_dev(b200_ctx *ctx, const int16_t *d_tmp1, const int16_t *d_tmp2, void *d_dst,
                    size_t nblocks, int w, int h, int bit_depth);

/* ------------------------------------------------------------------ CDEF (cdef.rs)
 * Per-call forms with the asm signatures rav1e binds (asm/x86/cdef.rs:16-37, :184-191):
 *   rav1e_cdef_dir_{8,16}bpc(img, stride_bytes, &var[, bitdepth_max]) -> dir
 *   rav1e_cdef_filter_{4x4,4x8,8x8}(dst, dst_stride, tmp_u16, tmp_stride, pri, sec, dir, damping)
 * `tmp` addresses the block's top-left inside the caller-built padded u16 tile (2 px border,
 * CDEF_VERY_LARGE = 0x8000 where pixels are unavailable, cdef.rs:161-194); strides in BYTES. */
int32_t b200_cdef_dir(const void *img, ptrdiff_t stride, uint32_t *var, int bit_depth);
void b200_cdef_filter_block(void *dst, ptrdiff_t dst_stride, const uint16_t *tmp,
                            ptrdiff_t tmp_stride, int pri_strength, int sec_strength, int dir,
                            int damping, int bit_depth, int xdec, int ydec);
/* Frame-level batch = cdef_filter_tile (cdef.rs:597-625) with the tile rect equal to the frame.
 * b200_cdef_find_dir_dev: cdef_analyze_superblock for every 8x8 luma block (d_skip8: one byte
 * per 8x8 block = AND of its four 4x4 `skip` flags, row-major, may be NULL); writes dir/var
 * (0 where skipped).  b200_cdef_filter_plane_dev: cdef_filter_superblock for one plane; `in`
 * and `out` are distinct planes of the same geometry; d_strength_sb holds the 6-bit strength
 * (pri*4+sec; fi.cdef_y_strengths / cdef_uv_strengths[cdef_index]) per 64x64 superblock,
 * row-major with ceil(luma_width/64) entries per row; damping = fi.cdef_damping. */
int b200_cdef_find_dir_dev(b200_ctx *ctx, const b200_plane *luma, int bit_depth,
                           const uint8_t *d_skip8, uint8_t *d_dir, int32_t *d_var);
int b200_cdef_filter_plane_dev(b200_ctx *ctx, const b200_plane *in, const b200_plane *out,
                               int plane, int xdec, int ydec, int luma_width, int luma_height,
                               int bit_depth, int damping, const uint8_t *d_skip8,
                               const uint8_t *d_dir, const int32_t *d_var,
                               const uint8_t *d_strength_sb);

/* The same restricted to the luma 8x8 blocks [rx8, rx8 + rw8) x [ry8, ry8 + rh8) of the frame (a tile of
 * cdef_filter_tile, cdef.rs:597-625: tiles run concurrently, taps still read across tile borders and
 * see the sentinel only outside the FRAME); dir / var / skip / strength arrays stay frame-indexed. */
int b200_cdef_find_dir_rect_dev(b200_ctx *ctx, const b200_plane *luma, int bit_depth, const uint8_t *d_skip8,
                                uint8_t *d_dir, int32_t *d_var, int rx8, int ry8, int rw8, int rh8);
int b200_cdef_filter_rect_dev(b200_ctx *ctx, const b200_plane *in, const b200_plane *out, int plane, int xdec,
                              int ydec, int luma_width, int luma_height, int bit_depth, int damping,
                              const uint8_t *d_skip8, const uint8_t *d_dir, const int32_t *d_var,
                              const uint8_t *d_strength_sb, int rx8, int ry8, int rw8, int rh8);

/* Many (plane, tile) items per call - the tiles a rank owns across the frames of a batch - served by
 * as few launches as possible (32 items per launch): with find_dir the directions / variances of every
 * item's blocks are (re)computed from `in` (luma only), with filter every item's rect of `in` is
 * filtered into `out`.  `items` is a HOST array; the arrays an item points to are frame-indexed. */
typedef struct {
  const b200_plane *in, *out; /* out may be NULL when only the directions are wanted */
  const uint8_t *d_skip8;     /* may be NULL */
  uint8_t *d_dir;
  int32_t *d_var;
  int32_t rx8, ry8, rw8, rh8; /* the tile in 8x8 luma blocks */
} b200_cdef_item;
int b200_cdef_tiles_dev(b200_ctx *ctx, const b200_cdef_item *items, size_t nitems, int plane, int xdec, int ydec,
                        int luma_width, int luma_height, int bit_depth, int damping, const uint8_t *d_strength_sb,
                        int find_dir, int filter);

/* ------------------------------------------------------- intra prediction (predict.rs)
 * Boundary = rust::dispatch_predict_intra(mode, variant, dst, tx_size, bit_depth, ac, angle,
 * ief_params, edge_buf, cpu) (predict.rs:705-784; the asm wrappers, asm/x86/predict.rs:239-873,
 * have the same Rust signature).  `edge` is the reference's IntraEdge buffer (partition.rs:
 * 600-637): 4*64+1 pixels, top-left at index 128, left pixels bottom->top in [128-left_len, 128),
 * above pixels in [129, 129+above_len).  mode = PredictionMode discriminant 0..13
 * (predict.rs:73-87); variant 0 NONE, 1 LEFT, 2 TOP, 3 BOTH (:112-118); angle = prediction angle
 * in degrees (intra_mode_to_angle + 3*angle_delta) or the CfL alpha; ief = -1 for None, else
 * IntraEdgeFilterParameters::use_smooth_filter() as 0/1 (:574-596); plane_w/plane_h/dst_x/dst_y =
 * dst.plane_cfg.{width,height} and dst.rect().{x,y} (edge-filter clipping, :1357-1364). */
void b200_predict_intra(int mode, int variant, void *dst, ptrdiff_t dst_stride, int w, int h,
                        int bit_depth, const int16_t *ac, int angle, int ief, const void *edge,
                        int left_len, int above_len, int plane_w, int plane_h, int dst_x, int dst_y);

typedef struct {
  uint32_t edge;    /* index of the IntraEdge buffer inside d_edges (257 pixels each) */
  uint32_t ac;      /* index of the w*h int16 CfL ac block inside d_ac (UV_CFL_PRED only) */
  int16_t x, y;     /* dst.rect().x / .y */
  int16_t angle;    /* degrees, or CfL alpha */
  uint8_t mode;     /* PredictionMode 0..13 */
  uint8_t variant;  /* PredictionVariant */
  int8_t ief;       /* -1 none, 0/1 = Some(use_smooth_filter) */
  uint8_t left_len, above_len;
  uint8_t pad_;
} b200_intra_item;

/* get_intra_edges (partition.rs:639-898) for a batch of transform blocks of one PlaneRegion: builds
 * the IntraEdge buffers b200_predict_intra_dev consumes, with the reference's availability rules
 * (has_top_right / has_bottom_left, recon_intra.rs:174-452), border constants and replication.
 * `plane` with (rect_x, rect_y, rect_w, rect_h) is the `dst` region (tile) the offsets are relative
 * to; plane->width / height are plane_cfg.width / height; xdec / ydec the plane's decimation.
 * d_edges: nitems x 257 pixels (top-left at [128]); entries outside
 * [128 - init_left, 129 + init_above) are zero (uninitialised upstream); d_lens (may be NULL):
 * {init_left, init_above} per item = the slice lengths IntraEdge::new records.  What may share a
 * batch is the caller's business: blocks whose neighbours are final (all of a frame in the lookahead,
 * one wavefront of the encode loop). */
typedef struct {
  int16_t po_x, po_y;     /* PlaneOffset of the transform block inside the region, pixels */
  int16_t part_x, part_y; /* partition_bo (TileBlockOffset), 4x4 luma units */
  uint8_t bx, by;         /* transform-block index inside the partition */
  uint8_t bsize;          /* partition BlockSize discriminant 0..21 (partition.rs:130-153) */
  uint8_t tx_size;        /* TxSize discriminant 0..18 */
  uint8_t mode;           /* PredictionMode 0..13, 255 = None (every edge wanted) */
  int8_t angle_delta;     /* IntraParam::AngleDelta, else 0 */
  uint8_t enable_ief;     /* enable_intra_edge_filter */
  uint8_t pad_;
} b200_edge_item;
int b200_get_intra_edges_dev(b200_ctx *ctx, const b200_plane *plane, int rect_x, int rect_y, int rect_w,
                             int rect_h, int xdec, int ydec, int bit_depth, const b200_edge_item *d_items,
                             size_t nitems, void *d_edges, uint8_t *d_lens);

/* Batched: one prediction per item (typically blocks x candidate modes sharing edge buffers);
 * d_out receives nitems packed w x h blocks (pixels). */
int b200_predict_intra_dev(b200_ctx *ctx, const void *d_edges, const b200_intra_item *d_items,
                           size_t nitems, const int16_t *d_ac, int w, int h, int bit_depth,
                           int plane_w, int plane_h, void *d_out);
/* pred_cfl_ac (predict.rs:1020-1063; asm cfl_ac_{420,422,444}, asm/x86/predict.rs:142-186):
 * bw x bh = chroma block size; d_blocks[i] = LUMA position of block i. */
int b200_pred_cfl_ac_dev(b200_ctx *ctx, const b200_plane *luma, const b200_block *d_blocks,
                         size_t nblocks, int bw, int bh, int w_pad, int h_pad, int xdec, int ydec,
                         int16_t *d_ac);

/* ------------------------------------------------ reference-signature symbols (MC / intra / CDEF)
 * The reference's own extern "C" prototypes, symbol for symbol, with the ISA suffix replaced by
 * `_cuda`: they slot into PUT_FNS / PUT_HBD_FNS / PREP_FNS / PREP_HBD_FNS / AVG_FNS / AVG_HBD_FNS
 * (asm/x86/mc.rs:322-620), the per-mode calls of asm::x86::predict::dispatch_predict_intra and
 * pred_cfl_ac (asm/x86/predict.rs:239-960), CDEF_FILTER_FNS and CDEF_DIR_{LBD,HBD}_FNS
 * (asm/x86/cdef.rs:161-176, :236-260) unchanged.  Host pointers, BYTE strides, exactly the
 * reference's argument lists (defined in rav1e_b200/csrc/ref_abi.cu over the per-call forms above).
 * Re-entrant: each host thread runs on its own context (one rayon worker per tile, encoder.rs:3253). */
/* X(name, FilterMode x, FilterMode y): table slot (x + 4 y) & 15, asm/x86/mc.rs:80-82, :394-405 */
#define B200_FOR_EACH_MC_FILTER(X)                                                              \
  X(8tap_regular, 0, 0) X(8tap_regular_smooth, 0, 1) X(8tap_regular_sharp, 0, 2)                \
  X(8tap_smooth_regular, 1, 0) X(8tap_smooth, 1, 1) X(8tap_smooth_sharp, 1, 2)                  \
  X(8tap_sharp_regular, 2, 0) X(8tap_sharp_smooth, 2, 1) X(8tap_sharp, 2, 2) X(bilin, 3, 3)
#define B200_DECL_MC(NAME, MX, MY)                                                                            \
  void rav1e_put_##NAME##_8bpc_cuda(uint8_t *dst, ptrdiff_t dst_stride, const uint8_t *src,                   \
                                    ptrdiff_t src_stride, int w, int h, int mx, int my);                      \
  void rav1e_put_##NAME##_16bpc_cuda(uint16_t *dst, ptrdiff_t dst_stride, const uint16_t *src,                \
                                     ptrdiff_t src_stride, int w, int h, int mx, int my, int bitdepth_max);   \
  void rav1e_prep_##NAME##_8bpc_cuda(int16_t *tmp, const uint8_t *src, ptrdiff_t src_stride, int w, int h,    \
                                     int mx, int my);                                                         \
  void rav1e_prep_##NAME##_16bpc_cuda(int16_t *tmp, const uint16_t *src, ptrdiff_t src_stride, int w, int h,  \
                                      int mx, int my, int bitdepth_max);
B200_FOR_EACH_MC_FILTER(B200_DECL_MC)
#undef B200_DECL_MC
void rav1e_avg_8bpc_cuda(uint8_t *dst, ptrdiff_t dst_stride, const int16_t *tmp1, const int16_t *tmp2, int w,
                         int h);
void rav1e_avg_16bpc_cuda(uint16_t *dst, ptrdiff_t dst_stride, const int16_t *tmp1, const int16_t *tmp2, int w,
                          int h, int bitdepth_max);

/* X(name, PredictionMode, PredictionVariant): asm/x86/predict.rs:21-120.  `topleft` = IntraEdge::
 * top_left_ptr() (element 128 of the 257-pixel buffer, partition.rs:600-627). */
#define B200_FOR_EACH_IPRED(X)                                                              \
  X(dc, 0, 3) X(dc_128, 0, 0) X(dc_left, 0, 1) X(dc_top, 0, 2) X(v, 1, 3) X(h, 2, 3)        \
  X(smooth, 9, 3) X(smooth_v, 10, 3) X(smooth_h, 11, 3) X(paeth, 12, 3)
#define B200_DECL_IPRED(NAME, MODE, VAR)                                                                      \
  void rav1e_ipred_##NAME##_8bpc_cuda(uint8_t *dst, ptrdiff_t stride, const uint8_t *topleft, int width,      \
                                      int height, int angle);                                                 \
  void rav1e_ipred_##NAME##_16bpc_cuda(uint16_t *dst, ptrdiff_t stride, const uint16_t *topleft, int width,   \
                                       int height, int angle, int max_width, int max_height,                  \
                                       int bitdepth_max);
B200_FOR_EACH_IPRED(B200_DECL_IPRED)
#undef B200_DECL_IPRED
/* directional zones; `angle` = degrees | enable_ief << 10 | smooth << 9 (asm/x86/predict.rs:287-322);
 * z2: dx, dy = distance from the block to the frame edge rounded up to 8 (:304-313) */
void rav1e_ipred_z1_8bpc_cuda(uint8_t *dst, ptrdiff_t stride, const uint8_t *topleft, int width, int height,
                              int angle);
void rav1e_ipred_z2_8bpc_cuda(uint8_t *dst, ptrdiff_t stride, const uint8_t *topleft, int width, int height,
                              int angle, int dx, int dy);
void rav1e_ipred_z3_8bpc_cuda(uint8_t *dst, ptrdiff_t stride, const uint8_t *topleft, int width, int height,
                              int angle);
void rav1e_ipred_z1_16bpc_cuda(uint16_t *dst, ptrdiff_t stride, const uint16_t *topleft, int width, int height,
                               int angle, int max_width, int max_height, int bitdepth_max);
void rav1e_ipred_z2_16bpc_cuda(uint16_t *dst, ptrdiff_t stride, const uint16_t *topleft, int width, int height,
                               int angle, int dx, int dy, int bitdepth_max);
void rav1e_ipred_z3_16bpc_cuda(uint16_t *dst, ptrdiff_t stride, const uint16_t *topleft, int width, int height,
                               int angle, int max_width, int max_height, int bitdepth_max);
/* X(name, PredictionVariant): asm/x86/predict.rs:188-234 */
#define B200_FOR_EACH_CFL(X) X(cfl, 3) X(cfl_128, 0) X(cfl_left, 1) X(cfl_top, 2)
#define B200_DECL_CFL(NAME, VAR)                                                                              \
  void rav1e_ipred_##NAME##_8bpc_cuda(uint8_t *dst, ptrdiff_t stride, const uint8_t *topleft, int width,      \
                                      int height, const int16_t *ac, int alpha);                              \
  void rav1e_ipred_##NAME##_16bpc_cuda(uint16_t *dst, ptrdiff_t stride, const uint16_t *topleft, int width,   \
                                       int height, const int16_t *ac, int alpha, int bitdepth_max);
B200_FOR_EACH_CFL(B200_DECL_CFL)
#undef B200_DECL_CFL
/* X(layout, xdec, ydec): asm/x86/predict.rs:142-186 */
#define B200_FOR_EACH_CFL_AC(X) X(420, 1, 1) X(422, 1, 0) X(444, 0, 0)
#define B200_DECL_CFL_AC(NAME, XDEC, YDEC)                                                                    \
  void rav1e_ipred_cfl_ac_##NAME##_8bpc_cuda(int16_t *ac, const uint8_t *src, ptrdiff_t stride, int w_pad,    \
                                             int h_pad, int width, int height);                               \
  void rav1e_ipred_cfl_ac_##NAME##_16bpc_cuda(int16_t *ac, const uint16_t *src, ptrdiff_t stride, int w_pad,  \
                                              int h_pad, int width, int height);
B200_FOR_EACH_CFL_AC(B200_DECL_CFL_AC)
#undef B200_DECL_CFL_AC

/* X(size, xdec, ydec): CdefFilterFn, slot decimate_index(xdec, ydec) (asm/x86/cdef.rs:16-42, :146-167).
 * The reference has no HBD filter asm (its CDEF_FILTER_HBD_FNS table is empty, :174-178): the
 * `_16bpc` forms take CdefFilterHBDFn's argument list (:27-37) over the same padded u16 tile. */
#define B200_FOR_EACH_CDEF_SIZE(X) X(4x4, 1, 1) X(4x8, 1, 0) X(8x8, 0, 0)
#define B200_DECL_CDEF(NAME, XDEC, YDEC)                                                                      \
  void rav1e_cdef_filter_##NAME##_cuda(uint8_t *dst, ptrdiff_t dst_stride, const uint16_t *tmp,               \
                                       ptrdiff_t tmp_stride, int pri_strength, int sec_strength, int dir,     \
                                       int damping);                                                          \
  void rav1e_cdef_filter_##NAME##_16bpc_cuda(uint16_t *dst, ptrdiff_t dst_stride, const uint16_t *tmp,        \
                                             ptrdiff_t tmp_stride, int pri_strength, int sec_strength,        \
                                             int dir, int damping, int bitdepth_max);
B200_FOR_EACH_CDEF_SIZE(B200_DECL_CDEF)
#undef B200_DECL_CDEF
int32_t rav1e_cdef_dir_8bpc_cuda(const uint8_t *img, ptrdiff_t stride, uint32_t *var);
int32_t rav1e_cdef_dir_16bpc_cuda(const uint16_t *img, ptrdiff_t stride, uint32_t *var, int bitdepth_max);

#ifdef __cplusplus
}
#endif
#endif /* B200RDO_H */
