// encode_tx.cu — encode_tx_block's numeric core (src/encoder.rs:1492-1655) for a batch of transform
// blocks, as ONE call that keeps every intermediate on the device:
//   diff (:1533) + forward_transform (:1544)     b200_fwd_txfm_residual_dev
//   ts.qc.quantize (:1556) + dequantize (:1590) + raw tx-domain distortion (:1611-1640)
//                                                 b200_quantize_dev
//   inverse_transform_add (:1600-1609)           b200_inverse_transform_add_dev
// The prediction is the reference plane displaced by the full-pel part of d_mv_src (NULL: zero
// motion), i.e. what b200_fwd_txfm_residual_dev subtracts; `rec` must already hold that prediction
// in the blocks' areas (predict-then-add, like the reference writes the prediction into `rec` before
// reconstructing).  Entropy coding of the quantized coefficients (write_coeffs_lv_map) and the rate /
// bias terms stay with the caller.
#include <cuda_runtime.h>

#include "common.cuh"

extern "C" int b200_encode_tx_blocks_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                                         const b200_plane *rec, const b200_block *d_blocks, size_t nblocks,
                                         const b200_me_result *d_mv_src, int tx_size, int tx_type, int bd,
                                         uint32_t dc_quant, uint32_t ac_quant, int is_intra,
                                         int need_recon_pixel, void *d_coeffs, void *d_qcoeffs,
                                         void *d_rcoeffs, uint16_t *d_eob, uint64_t *d_tx_dist) {
  B200_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
  B200_REQUIRE(ctx, cur && ref && d_coeffs && d_qcoeffs && d_rcoeffs, "NULL plane / coefficient buffers");
  B200_REQUIRE(ctx, !need_recon_pixel || rec, "need_recon_pixel without a reconstruction plane");
  if (int st = b200_fwd_txfm_residual_dev(ctx, cur, ref, d_blocks, nblocks, d_mv_src, d_coeffs, tx_size, tx_type, bd))
    return st;
  if (int st = b200_quantize_dev(ctx, d_coeffs, nblocks, tx_size, tx_type, dc_quant, ac_quant, is_intra,
                                 cur->bpp == 2, d_qcoeffs, d_rcoeffs, d_eob, d_tx_dist))
    return st;
  // encoder.rs:1598-1609: all-zero blocks are a no-op for the inverse (it adds zeros)
  if (need_recon_pixel)
    return b200_inverse_transform_add_dev(ctx, d_rcoeffs, rec, d_blocks, nblocks, tx_size, tx_type, bd);
  return B200_OK;
}
