// rav1e_b200.hpp — host-side mirror (C++17, header only) of the reference's kernel interface,
// above the C ABI of include/b200rdo.h.
//
// rav1e's L1 kernels are generic Rust functions over `PlaneRegion<T>` views with a trailing
// `cpu: CpuFeatureLevel` argument (dist.rs:31,156; transform/forward.rs:71; mc.rs:250,360,454;
// predict.rs:705; cdef.rs:84,198); `src/asm/<arch>/` re-implements each with the same signature
// and a table lookup.  Rust is not available in this image, so this header restates that
// interface in C++ — same names, argument order and meaning, same precondition behaviour
// (violations throw std::invalid_argument where the reference `assert!`s) — and routes every
// call to the CUDA backend.  There is deliberately no CPU implementation here: any level other
// than CUDA_SM100 throws.  The Rust module a maintainer would add is shown in INTEGRATION.md.
#pragma once

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/b200rdo.h"

namespace rav1e_b200 {

// cpu_features/x86.rs:14-25 with the new top level appended (ordered: `cpu >= AVX2` style
// comparisons in the wrappers keep working, asm/x86/transform/forward.rs:449).
enum class CpuFeatureLevel { RUST, SSE2, SSSE3, SSE4_1, AVX2, AVX512, AVX512ICL, CUDA_SM100 };

inline void require_cuda(CpuFeatureLevel cpu) {
  if (cpu != CpuFeatureLevel::CUDA_SM100)
    throw std::invalid_argument("rav1e_b200: only CpuFeatureLevel::CUDA_SM100 is implemented (no CPU fallback)");
}

// tiling/plane_region.rs:116-135: a bounded view; data_ptr() = element pointer at (rect.x, rect.y),
// plane_cfg.stride in ELEMENTS (the asm ABI takes bytes: T::to_asm_stride).
template <typename T>
struct PlaneRegion {
  const T *data;
  std::ptrdiff_t stride;
  std::size_t width, height;
  const T *data_ptr() const { return data; }
  std::ptrdiff_t asm_stride() const { return stride * (std::ptrdiff_t)sizeof(T); }
  PlaneRegion subregion(std::size_t x, std::size_t y, std::size_t w, std::size_t h) const {
    if (x + w > width || y + h > height) throw std::out_of_range("subregion outside region");
    return {data + (std::ptrdiff_t)y * stride + (std::ptrdiff_t)x, stride, w, h};
  }
};
template <typename T>
struct PlaneRegionMut {
  T *data;
  std::ptrdiff_t stride;
  std::size_t width, height;
};

// dist.rs:31-52 / asm/x86/dist/mod.rs:287-334
template <typename T>
std::uint32_t get_sad(const PlaneRegion<T> &plane_org, const PlaneRegion<T> &plane_ref, std::size_t w,
                      std::size_t h, std::size_t /*bit_depth*/, CpuFeatureLevel cpu) {
  require_cuda(cpu);
  if (w > 128 || h > 128) throw std::invalid_argument("get_sad: w and h can be at most 128");
  return b200_get_sad(plane_org.data_ptr(), plane_org.asm_stride(), plane_ref.data_ptr(),
                      plane_ref.asm_stride(), (int)w, (int)h, (int)sizeof(T));
}

// dist.rs:156-221 / asm/x86/dist/mod.rs:341-389
template <typename T>
std::uint32_t get_satd(const PlaneRegion<T> &plane_org, const PlaneRegion<T> &plane_ref, std::size_t w,
                       std::size_t h, std::size_t /*bit_depth*/, CpuFeatureLevel cpu) {
  require_cuda(cpu);
  if (w > 128 || h > 128) throw std::invalid_argument("get_satd: w and h can be at most 128");  // dist.rs:160
  if (plane_org.width < w || plane_org.height < h || plane_ref.width < w || plane_ref.height < h)
    throw std::invalid_argument("get_satd: region smaller than the block");                       // dist.rs:161-162
  return b200_get_satd(plane_org.data_ptr(), plane_org.asm_stride(), plane_ref.data_ptr(),
                       plane_ref.asm_stride(), (int)w, (int)h, (int)sizeof(T));
}

// dist.rs:234-283 / asm/x86/dist/sse.rs:118-148.  scale: one DistortionScale (Q14) per 4x4 chunk,
// scale_stride in ENTRIES like the Rust signature.
template <typename T>
std::uint64_t get_weighted_sse(const PlaneRegion<T> &src1, const PlaneRegion<T> &src2, const std::uint32_t *scale,
                               std::size_t scale_stride, std::size_t w, std::size_t h, std::size_t /*bit_depth*/,
                               CpuFeatureLevel cpu) {
  require_cuda(cpu);
  if (w > 128 || h > 128 || (w & 3) || (h & 3)) throw std::invalid_argument("get_weighted_sse: bad block size");
  if (src1.width < w || src1.height < h || src2.width < w || src2.height < h)
    throw std::invalid_argument("get_weighted_sse: region smaller than the block");
  return b200_weighted_sse(src1.data_ptr(), src1.asm_stride(), src2.data_ptr(), src2.asm_stride(), scale,
                           (std::ptrdiff_t)(scale_stride * sizeof(std::uint32_t)), (int)w, (int)h, (int)sizeof(T));
}

// dist.rs:302-372 / asm/x86/dist/cdef_dist.rs:56-118 (w, h <= 8)
template <typename T>
std::uint32_t cdef_dist_kernel(const PlaneRegion<T> &src, const PlaneRegion<T> &dst, std::size_t w, std::size_t h,
                               std::size_t bit_depth, CpuFeatureLevel cpu) {
  require_cuda(cpu);
  if (w > 8 || h > 8 || w == 0 || h == 0) throw std::invalid_argument("cdef_dist_kernel: w and h must be <= 8");
  if ((sizeof(T) == 1) != (bit_depth == 8)) throw std::invalid_argument("cdef_dist_kernel: pixel type / bit depth");
  return b200_cdef_dist_kernel(src.data_ptr(), src.asm_stride(), dst.data_ptr(), dst.asm_stride(), (int)w, (int)h,
                               (int)bit_depth, nullptr);
}

// transform/mod.rs:56-123
enum class TxType : int { DCT_DCT = 0, ADST_DCT, DCT_ADST, ADST_ADST, FLIPADST_DCT, DCT_FLIPADST,
                          FLIPADST_FLIPADST, ADST_FLIPADST, FLIPADST_ADST, IDTX, V_DCT, H_DCT, V_ADST,
                          H_ADST, V_FLIPADST, H_FLIPADST, WHT_WHT };
enum class TxSize : int { TX_4X4 = 0, TX_8X8, TX_16X16, TX_32X32, TX_64X64, TX_4X8, TX_8X4, TX_8X16,
                          TX_16X8, TX_16X32, TX_32X16, TX_32X64, TX_64X32, TX_4X16, TX_16X4, TX_8X32,
                          TX_32X8, TX_16X64, TX_64X16 };
inline bool valid_av1_transform(TxSize s, TxType t) { return b200_valid_av1_transform((int)s, (int)t) != 0; }

// transform/forward.rs:71-161.  Coeff = int16_t (8-bit pixels) or int32_t (HBD), `T::Coeff`.
template <typename Coeff>
void forward_transform(const std::int16_t *input, Coeff *output, std::size_t stride, TxSize tx_size,
                       TxType tx_type, std::size_t bd, CpuFeatureLevel cpu) {
  static_assert(std::is_same<Coeff, std::int16_t>::value || std::is_same<Coeff, std::int32_t>::value,
                "Coefficient is i16 or i32");
  require_cuda(cpu);
  if (!valid_av1_transform(tx_size, tx_type)) throw std::invalid_argument("invalid (tx_size, tx_type)");  // :75
  b200_forward_transform(input, output, stride, (int)tx_size, (int)tx_type, (int)bd, sizeof(Coeff) == 4);
}

// mc.rs:98-106
enum class FilterMode : int { REGULAR = 0, SMOOTH = 1, SHARP = 2, BILINEAR = 3 };

inline void check_mc_dims(std::size_t width, std::size_t height) {
  if (height & 1) throw std::invalid_argument("put/prep_8tap: height must be even");                        // mc.rs:256
  if (width < 2 || width > 128 || (width & (width - 1))) throw std::invalid_argument("width must be 2^k in 2..128");  // :257
}

// mc.rs:250-353.  `src` addresses the block's top-left in a plane readable over -3..+4.
template <typename T>
void put_8tap(PlaneRegionMut<T> &dst, const T *src, std::ptrdiff_t src_stride, std::size_t width,
              std::size_t height, int col_frac, int row_frac, FilterMode mode_x, FilterMode mode_y,
              std::size_t bit_depth, CpuFeatureLevel cpu) {
  require_cuda(cpu);
  check_mc_dims(width, height);
  b200_put_8tap(dst.data, dst.stride * (std::ptrdiff_t)sizeof(T), src, src_stride * (std::ptrdiff_t)sizeof(T),
                (int)width, (int)height, col_frac, row_frac, (int)mode_x, (int)mode_y, (int)bit_depth);
}

// mc.rs:360-451
template <typename T>
void prep_8tap(std::int16_t *tmp, const T *src, std::ptrdiff_t src_stride, std::size_t width,
               std::size_t height, int col_frac, int row_frac, FilterMode mode_x, FilterMode mode_y,
               std::size_t bit_depth, CpuFeatureLevel cpu) {
  require_cuda(cpu);
  check_mc_dims(width, height);
  b200_prep_8tap(tmp, src, src_stride * (std::ptrdiff_t)sizeof(T), (int)width, (int)height, col_frac,
                 row_frac, (int)mode_x, (int)mode_y, (int)bit_depth);
}

// mc.rs:454-479
template <typename T>
void mc_avg(PlaneRegionMut<T> &dst, const std::int16_t *tmp1, const std::int16_t *tmp2, std::size_t width,
            std::size_t height, std::size_t bit_depth, CpuFeatureLevel cpu) {
  require_cuda(cpu);
  check_mc_dims(width, height);
  b200_mc_avg(dst.data, dst.stride * (std::ptrdiff_t)sizeof(T), tmp1, tmp2, (int)width, (int)height,
              (int)bit_depth);
}

// cdef.rs:84-143: returns the direction, writes the variance.
template <typename T>
int cdef_find_dir(const T *img, std::ptrdiff_t stride, std::uint32_t *var, std::size_t coeff_shift,
                  CpuFeatureLevel cpu) {
  require_cuda(cpu);
  return b200_cdef_dir(img, stride * (std::ptrdiff_t)sizeof(T), var, (int)coeff_shift + 8);
}

// predict.rs:58-118
enum class PredictionMode : int { DC_PRED = 0, V_PRED, H_PRED, D45_PRED, D135_PRED, D113_PRED, D157_PRED,
                                  D203_PRED, D67_PRED, SMOOTH_PRED, SMOOTH_V_PRED, SMOOTH_H_PRED,
                                  PAETH_PRED, UV_CFL_PRED };
enum class PredictionVariant : int { NONE = 0, LEFT, TOP, BOTH };

// partition.rs:600-637: 4*64+1 pixels, top-left at index 128
template <typename T>
struct IntraEdge {
  const T *buf;
  int left_len, above_len;
};

// predict.rs:705-784.  ief: -1 = None, else IntraEdgeFilterParameters::use_smooth_filter().
template <typename T>
void dispatch_predict_intra(PredictionMode mode, PredictionVariant variant, PlaneRegionMut<T> &dst, int plane_w,
                            int plane_h, int dst_x, int dst_y, std::size_t tx_w, std::size_t tx_h,
                            std::size_t bit_depth, const std::int16_t *ac, int angle, int ief,
                            const IntraEdge<T> &edge, CpuFeatureLevel cpu) {
  require_cuda(cpu);
  b200_predict_intra((int)mode, (int)variant, dst.data, dst.stride * (std::ptrdiff_t)sizeof(T), (int)tx_w,
                     (int)tx_h, (int)bit_depth, ac, angle, ief, edge.buf, edge.left_len, edge.above_len,
                     plane_w, plane_h, dst_x, dst_y);
}

// rdo.rs:718-723: `compute_rd_cost(fi, rate, distortion)` with lambda = fi.lambda: one correctly rounded
// f64 fma of rate / 8 (rate is in 1/8 bit units, OD_BITRES = 3).
inline double compute_rd_cost(double lambda, std::uint32_t rate, std::uint64_t distortion, CpuFeatureLevel cpu) {
  require_cuda(cpu);
  return b200_compute_rd_cost(lambda, rate, distortion);
}

// ------------------------------------------------------------------------------------------
// The lookahead-shaped call (b200_frame_pipe_*): one push per frame, the previous frame is the reference.
// Host buffers in and out; winners of the SAD and SATD lists and the coefficients of the SAD winner's residual.
class FramePipe {
 public:
  FramePipe(b200_ctx *ctx, const b200_frame_pipe_cfg &cfg) {
    if (b200_frame_pipe_create(ctx, &cfg, &pipe_) != B200_OK) throw std::runtime_error(b200_last_error(ctx));
  }
  ~FramePipe() { b200_frame_pipe_destroy(pipe_); }
  FramePipe(const FramePipe &) = delete;
  FramePipe &operator=(const FramePipe &) = delete;
  std::size_t nblocks() const { return b200_frame_pipe_nblocks(pipe_); }
  // offsets: (row, col) full-pel int8 pairs per candidate; NULL outputs are skipped
  int push(const void *frame, std::ptrdiff_t stride_bytes, const std::int8_t *sad_offsets,
           const std::int8_t *satd_offsets, const std::int16_t *centers, b200_me_result *best_sad,
           b200_me_result *best_satd, void *coeffs, std::uint16_t *eob = nullptr, std::uint64_t *tx_dist = nullptr) {
    return b200_frame_pipe_push(pipe_, frame, stride_bytes, sad_offsets, satd_offsets, centers, best_sad, best_satd,
                                coeffs, eob, tx_dist);
  }

 private:
  b200_frame_pipe *pipe_ = nullptr;
};

// ------------------------------------------------------------------------------------------
// Batched motion search helper: what me.rs's serial candidate loops become.  One call
// evaluates every candidate of every block and returns per-block winners with the reference's
// first-minimum tie-break.
class Context {
 public:
  explicit Context(int device = 0) {
    if (b200_ctx_create(device, &ctx_) != B200_OK) throw std::runtime_error(b200_last_error(nullptr));
  }
  ~Context() { b200_ctx_destroy(ctx_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  b200_ctx *get() const { return ctx_; }

  std::vector<b200_me_result> me_candidates(const b200_host_plane &cur, const b200_host_plane &ref,
                                            const std::vector<b200_block> &blocks,
                                            const std::vector<b200_cand> &cands,
                                            const std::vector<std::uint32_t> &offsets,
                                            const b200_me_params &params,
                                            std::vector<std::uint32_t> *sad_out = nullptr) {
    std::vector<b200_me_result> best(blocks.size());
    if (sad_out) sad_out->resize(cands.size());
    const int st = b200_me_candidates_batch(ctx_, &cur, &ref, blocks.data(), blocks.size(), cands.data(),
                                            cands.size(), offsets.data(), nullptr, &params,
                                            sad_out ? sad_out->data() : nullptr, nullptr, best.data());
    if (st != B200_OK) throw std::runtime_error(b200_last_error(ctx_));
    return best;
  }

 private:
  b200_ctx *ctx_ = nullptr;
};

}  // namespace rav1e_b200
