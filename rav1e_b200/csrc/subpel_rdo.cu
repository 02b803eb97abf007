// subpel_rdo.cu — sub-pel refinement fused with the winner's residual and forward transform (sm_100a):
// BASELINE configs[3] "speed-2 inter RDO (mc subpel + SATD + fwd-txfm fused)".
//
//   get_subpel_mv_rd           src/me.rs:1411-1442   (per candidate: predict_inter_single -> compute_mv_rd)
//   predict_inter_single       src/predict.rs:304-336 -> put_8tap, src/mc.rs:250-353
//   get_sad / get_satd         src/dist.rs:31-52, :156-221
//   compute_mv_rd, get_mv_rate src/me.rs:1445-1461, :1512-1523
//   diff + forward_transform   src/encoder.rs:1533-1544, src/transform/forward.rs:71-161
//
// The two-launch form (b200_me_subpel_candidates_dev: mc kernel -> packed predictions in HBM ->
// distortion kernel) writes and re-reads w*h pixels per candidate.  Here ONE warp owns a block and walks
// its candidates: the (w+7) x (h+7) source footprint is staged in shared memory, the 8-tap prediction is
// formed in shared memory (the reference's four cases incl. the double-rounded H-only path and the i16
// intermediate), measured against the source block where it sits - SAD by lanes striding the pixels,
// SATD with 8 lanes per 8x8 chunk (lane = row: horizontal butterflies in registers, vertical ones by
// shfl.xor) - and the best prediction so far is kept by swapping two shared buffers.  When the list is
// exhausted the winner's prediction is still on chip: its residual goes straight through the column /
// row passes of the forward transform (thread per column, padded transposition tile, thread per row).
// Nothing but the source footprints is read from HBM and nothing but results is written.
#include <cuda_runtime.h>

#include <algorithm>
#include <mutex>

#include "fwd_txfm_dev.cuh"
#include "mc_dev.cuh"

namespace {

constexpr unsigned long long kEmptyCost = ~0ull;
constexpr uint32_t kEmptySad = ~0u;

struct SrdoArgs {
  const void *cur, *ref;
  int cur_stride, ref_stride;
  const b200_block *blocks;
  const b200_cand *cands;
  const uint32_t *cand_offsets;
  const short *pmv;
  const b200_me_result *start;  // non-NULL: subpel_diamond_search from this result instead of a list
  size_t nblocks;
  int w_in_b, h_in_b;
  uint32_t lambda;
  int allow_hp, use_satd, mode, bit_depth;
  uint32_t *out_sad;
  unsigned long long *out_cost;
  b200_me_result *out_best;
  int tx_on;
  TxSetup tx;
  void *out_coef;
  int smem_per_warp, pred_off;  // bytes: a warp's slice, and where its two prediction buffers start
};

__device__ __forceinline__ void bfly(int &a, int &b) {
  const int s = a + b, t = a - b;
  a = s;
  b = t;
}

// SATD of the W x H block: 8 lanes per 8x8 chunk (lane = row), four chunks per step
template <typename T, int W, int H>
__device__ __forceinline__ uint32_t warp_satd(const T *org, int os, const T *pred, int lane) {
  constexpr int NCH = (W / 8) * (H / 8);
  const int row = lane & 7, g = lane >> 3;
  uint32_t acc = 0;
#pragma unroll
  for (int c0 = 0; c0 < NCH; c0 += 4) {
    const int ch = c0 + g;
    int v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = 0;
    if (NCH % 4 == 0 || ch < NCH) {
      const int cy = (ch / (W / 8)) * 8, cx = (ch % (W / 8)) * 8;
      const T *o = org + (long long)(cy + row) * os + cx;
      const T *p = pred + (cy + row) * W + cx;
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = (int)o[k] - (int)p[k];
    }
    // horizontal pass in registers (dist.rs:88-121 hadamard8_1d)
    bfly(v[0], v[1]);
    bfly(v[2], v[3]);
    bfly(v[4], v[5]);
    bfly(v[6], v[7]);
    bfly(v[0], v[2]);
    bfly(v[1], v[3]);
    bfly(v[4], v[6]);
    bfly(v[5], v[7]);
    bfly(v[0], v[4]);
    bfly(v[1], v[5]);
    bfly(v[2], v[6]);
    bfly(v[3], v[7]);
    // vertical pass across the 8 lanes of the chunk
#pragma unroll
    for (int s = 1; s < 8; s <<= 1) {
      const bool hi = (row & s) != 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int other = __shfl_xor_sync(0xffffffffu, v[k], s);
        v[k] = hi ? other - v[k] : v[k] + other;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) acc += (uint32_t)abs(v[k]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return (acc + 4u) >> 3;  // dist.rs:219-220 (8x8 chunks: ln = 3), one rounding for the block
}

template <typename T, int W, int H>
__device__ __forceinline__ uint32_t warp_sad(const T *org, int os, const T *pred, int lane) {
  uint32_t acc = 0;
#pragma unroll 4
  for (int i = lane; i < W * H; i += 32) {
    const int r = i / W, c = i - r * W;
    acc += (uint32_t)abs((int)org[(long long)r * os + c] - (int)pred[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

template <typename T, int W, int H>
__global__ void __launch_bounds__(256) subpel_rdo_kernel(const __grid_constant__ SrdoArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using CoefT = typename std::conditional<sizeof(T) == 1, int16_t, int32_t>::type;
  using ML = McLayout<T, W, H>;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned char *base = smem_raw + (size_t)wid * a.smem_per_warp;
  T *tile = (T *)base;                                  // [H+7][W+8] source footprint
  short *inter = (short *)(base + ML::TILE_BYTES);      // [W][H+8] transposed i16 intermediate
  T *pred0 = (T *)(base + a.pred_off);                  // two W x H predictions: current / best so far
  T *pred1 = pred0 + W * H;
  int *txtile = (int *)base;                            // aliases tile + inter once the list is done
  const int xb = filter_bank(a.mode, W), yb = filter_bank(a.mode, H);

  for (size_t blk = (size_t)blockIdx.x * nw + wid; blk < a.nblocks; blk += (size_t)gridDim.x * nw) {
    const b200_block b = a.blocks[blk];
    const uint32_t lo = a.start ? 0u : a.cand_offsets[blk], hi = a.start ? 0u : a.cand_offsets[blk + 1];
    const MvRange rng = b200_mv_range(a.w_in_b, a.h_in_b, b.x / MI_SIZE, b.y / MI_SIZE, W, H);
    int p0r = 0, p0c = 0, p1r = 0, p1c = 0;
    if (a.pmv) {
      const short *p = a.pmv + 4 * blk;
      p0r = p[0], p0c = p[1], p1r = p[2], p1c = p[3];
    }
    const T *org = (const T *)a.cur + (long long)b.y * a.cur_stride + b.x;
    T *pcur = pred0, *pbest = pred1;
    unsigned long long best_cost = kEmptyCost;
    uint32_t best_sad = kEmptySad;
    int best_row = 0, best_col = 0;
    // get_subpel_mv_rd (me.rs:1411-1442) of one vector: prediction into `pcur`, distortion, cost
    auto eval = [&](int mv_row, int mv_col, uint32_t &sad) -> unsigned long long {
      sad = kEmptySad;
      if (mv_col < rng.x_min || mv_col > rng.x_max || mv_row < rng.y_min || mv_row > rng.y_max) return kEmptyCost;
      // predict.rs:284-297 get_mv_params (luma)
      const int y0 = b.y + (mv_row >> 3), x0 = b.x + (mv_col >> 3);
      const int row_frac = (int)(((unsigned)mv_row << 1) & 0xf), col_frac = (int)(((unsigned)mv_col << 1) & 0xf);
      const T *src = (const T *)a.ref + (long long)(y0 - 3) * a.ref_stride + (x0 - 3);
      __syncwarp();
      mc_stage_tile<T, W, H>(tile, src, a.ref_stride, lane);
      __syncwarp();
      mc_put_warp<T, W, H>(tile, inter, pcur, xb, col_frac, yb, row_frac, a.bit_depth, lane);
      __syncwarp();
      sad = a.use_satd ? warp_satd<T, W, H>(org, a.cur_stride, pcur, lane) : warp_sad<T, W, H>(org, a.cur_stride, pcur, lane);
      return b200_mv_cost(sad, mv_row, mv_col, p0r, p0c, p1r, p1c, a.lambda, a.allow_hp);
    };
    if (a.start) {
      // ---- subpel_diamond_search, me.rs:1311-1383: move to the best of the four diamond points while
      // one is strictly better, then halve the radius (1/2 pel down to 1/4, or 1/8 with high precision)
      const b200_me_result st = a.start[blk];
      best_cost = st.cost, best_sad = st.sad, best_row = st.mv_row, best_col = st.mv_col;
      int radius_log2 = 2;
      const int end_log2 = a.allow_hp ? 0 : 1;
      for (;;) {
        unsigned long long cc = kEmptyCost;
        uint32_t cs = kEmptySad;
        int cr = 0, ccol = 0;
#pragma unroll 1
        for (int k = 0; k < 4; k++) {  // DIAMOND_R1_PATTERN_SUBPEL, me.rs:931-934: row+1, col+1, row-1, col-1
          const int dr = (k == 0) - (k == 2), dc = (k == 1) - (k == 3);
          const int mr = (short)(best_row + (short)(dr << radius_log2)), mc = (short)(best_col + (short)(dc << radius_log2));
          uint32_t sad;
          const unsigned long long cost = eval(mr, mc, sad);
          if (cost < cc) cc = cost, cs = sad, cr = mr, ccol = mc;
        }
        if (best_cost <= cc) {
          if (radius_log2 == end_log2) break;
          radius_log2--;
        } else {
          best_cost = cc, best_sad = cs, best_row = cr, best_col = ccol;
        }
      }
      if (a.tx_on && best_cost != kEmptyCost) {  // the winner's prediction for the transform below
        uint32_t sad;
        (void)eval(best_row, best_col, sad);
        T *t = pcur;
        pcur = pbest;
        pbest = t;
      }
    } else {
      for (uint32_t i = lo; i < hi; i++) {
        const b200_cand c = a.cands[i];
        uint32_t sad;
        const unsigned long long cost = eval(c.mv_row, c.mv_col, sad);
        if (cost < best_cost) {  // strict: the first minimum in list order wins (me.rs:1360-1365)
          best_cost = cost;
          best_sad = sad;
          best_row = c.mv_row;
          best_col = c.mv_col;
          T *t = pcur;
          pcur = pbest;
          pbest = t;
        }
        if (lane == 0) {
          if (a.out_sad) a.out_sad[i] = sad;
          if (a.out_cost) a.out_cost[i] = cost;
        }
      }
    }
    if (lane == 0 && a.out_best) {
      b200_me_result res;
      res.cost = best_cost;
      res.sad = best_sad;
      res.mv_row = (int16_t)best_row;
      res.mv_col = (int16_t)best_col;
      a.out_best[blk] = res;
    }
    if (a.tx_on) {
      // ---- residual of the winner's prediction -> forward transform (forward.rs:94-160)
      constexpr int PITCH = W + 1;
      CoefT *dst = (CoefT *)a.out_coef + blk * (size_t)(W * H);
      __syncwarp();
      if (best_cost == kEmptyCost) {
        for (int k = lane; k < W * H; k += 32) dst[k] = 0;
      } else {
        for (int t = lane; t < W; t += 32) {
          TXV c[H];
#pragma unroll
          for (int r = 0; r < H; r++) {
            const int rr = a.tx.ud_flip ? H - 1 - r : r;
            c[r] = round_shift_bit((int)org[(long long)rr * a.cur_stride + t] - (int)pbest[rr * W + t], a.tx.bit0);
          }
          run_1d<H>(a.tx.col_type, c);
          const int cc = a.tx.lr_flip ? W - 1 - t : t;
#pragma unroll
          for (int r = 0; r < H; r++) txtile[r * PITCH + cc] = round_shift_bit(c[r], a.tx.bit1);
        }
        __syncwarp();
        for (int t = lane; t < H; t += 32) {
          TXV c[W];
#pragma unroll
          for (int k = 0; k < W; k++) c[k] = txtile[t * PITCH + k];
          run_1d<W>(a.tx.row_type, c);
#pragma unroll
          for (int k = 0; k < W; k++) dst[k * H + t] = (CoefT)round_shift_bit(c[k], a.tx.bit2);  // W, H <= 32
        }
      }
      __syncwarp();
    }
  }
}

template <typename T, int W, int H>
int launch_srdo(b200_ctx *ctx, SrdoArgs a) {
  using ML = McLayout<T, W, H>;
  const size_t mc_bytes = b200_align_up(ML::TILE_BYTES, 4) + ML::INTER_BYTES;
  const size_t tx_bytes = (size_t)H * (W + 1) * 4;  // aliases the MC buffers
  // the predictions sit behind whichever of the two is larger
  const size_t pred_off = b200_align_up(std::max(mc_bytes, tx_bytes), 16);
  const size_t per_warp = b200_align_up(pred_off + 2 * (size_t)W * H * sizeof(T), 16);
  const int wpc = (int)std::min<size_t>(8, std::max<size_t>(1, (160 * 1024) / per_warp));
  const size_t smem = per_warp * wpc;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(subpel_rdo_kernel<T, W, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  B200_CUDA(ctx, err);
  a.smem_per_warp = (int)per_warp;
  a.pred_off = (int)pred_off;
  const int grid = (int)std::min<size_t>((a.nblocks + wpc - 1) / wpc, (size_t)ctx->num_sms * 16);
  subpel_rdo_kernel<T, W, H><<<grid, wpc * 32, smem, ctx->stream>>>(a);
  B200_LAUNCH_CHECK(ctx);
  return B200_OK;
}

}  // namespace

static int subpel_rdo_impl(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref, const b200_block *d_blocks,
                           size_t nblocks, const b200_cand *d_cands, size_t ncands, const uint32_t *d_cand_offsets,
                           const b200_me_result *d_start, const int16_t *d_pmv, const b200_me_params *p,
                           int filter_mode, int tx_size, int tx_type, uint32_t *d_sad, uint64_t *d_cost,
                           b200_me_result *d_best, void *d_coeffs) {
  B200_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
  B200_REQUIRE(ctx, cur && ref && p && cur->data && ref->data, "NULL plane / params");
  B200_REQUIRE(ctx, cur->bpp == ref->bpp && (cur->bpp == 1 || cur->bpp == 2), "planes must share bpp (1 or 2)");
  B200_REQUIRE(ctx, (cur->bpp == 1) == (p->bit_depth == 8) && (p->bit_depth == 8 || p->bit_depth == 10 || p->bit_depth == 12),
               "bpp %d vs bit depth %d", cur->bpp, p->bit_depth);
  B200_REQUIRE(ctx, p->w == p->h && (p->w == 8 || p->w == 16 || p->w == 32),
               "fused sub-pel RDO serves 8x8, 16x16 and 32x32 blocks (got %dx%d): use "
               "b200_me_subpel_candidates_dev + b200_fwd_txfm_pred_dev for other sizes", p->w, p->h);
  B200_REQUIRE(ctx, filter_mode >= 0 && filter_mode <= 3, "bad FilterMode %d", filter_mode);
  B200_REQUIRE(ctx, d_start != nullptr || d_cand_offsets != nullptr, "candidates must be grouped by block (CSR d_cand_offsets)");
  B200_REQUIRE(ctx, tx_size < 0 || (valid_transform(tx_size, tx_type) && kTxW[tx_size] == p->w && kTxH[tx_size] == p->h),
               "transform %d/%d does not match the %dx%d blocks", tx_size, tx_type, p->w, p->h);
  B200_REQUIRE(ctx, tx_size < 0 || d_coeffs, "transform requested without an output buffer");
  if (nblocks == 0) return B200_OK;
  B200_REQUIRE(ctx, d_blocks && (d_start || d_cands || ncands == 0), "NULL blocks / candidates");
  B200_CUDA(ctx, cudaSetDevice(ctx->device));
  SrdoArgs a{};
  a.cur = cur->data, a.ref = ref->data;
  a.cur_stride = cur->stride, a.ref_stride = ref->stride;
  a.blocks = d_blocks;
  a.cands = d_cands;
  a.cand_offsets = d_cand_offsets;
  a.start = d_start;
  a.pmv = d_pmv;
  a.nblocks = nblocks;
  a.w_in_b = p->frame_w_in_b, a.h_in_b = p->frame_h_in_b;
  a.lambda = p->lambda;
  a.allow_hp = p->allow_high_precision_mv;
  a.use_satd = p->use_satd;
  a.mode = filter_mode;
  a.bit_depth = p->bit_depth;
  a.out_sad = d_sad;
  a.out_cost = (unsigned long long *)d_cost;
  a.out_best = d_best;
  a.tx_on = tx_size >= 0;
  if (a.tx_on) a.tx = tx_setup(tx_size, tx_type, p->bit_depth);
  a.out_coef = d_coeffs;
#define B200_SRDO(N)                                                                     \
  if (p->w == N)                                                                         \
    return cur->bpp == 1 ? launch_srdo<uint8_t, N, N>(ctx, a) : launch_srdo<uint16_t, N, N>(ctx, a);
  B200_SRDO(8)
  B200_SRDO(16)
  B200_SRDO(32)
#undef B200_SRDO
  return b200_fail(ctx, B200_ERR_ARG, "unreachable block size");
}

extern "C" int b200_subpel_rdo_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                                   const b200_block *d_blocks, size_t nblocks, const b200_cand *d_cands,
                                   size_t ncands, const uint32_t *d_cand_offsets, const int16_t *d_pmv,
                                   const b200_me_params *p, int filter_mode, int tx_size, int tx_type,
                                   uint32_t *d_sad, uint64_t *d_cost, b200_me_result *d_best, void *d_coeffs) {
  B200_REQUIRE(ctx, ctx != nullptr && d_cand_offsets != nullptr, "NULL ctx / candidates must be grouped by block (CSR)");
  return subpel_rdo_impl(ctx, cur, ref, d_blocks, nblocks, d_cands, ncands, d_cand_offsets, nullptr, d_pmv, p,
                         filter_mode, tx_size, tx_type, d_sad, d_cost, d_best, d_coeffs);
}

extern "C" int b200_subpel_search_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                                      const b200_block *d_blocks, size_t nblocks, const b200_me_result *d_start,
                                      const int16_t *d_pmv, const b200_me_params *p, int filter_mode, int tx_size,
                                      int tx_type, b200_me_result *d_best, void *d_coeffs) {
  B200_REQUIRE(ctx, ctx != nullptr && d_start != nullptr && d_best != nullptr, "NULL ctx / start results / output");
  return subpel_rdo_impl(ctx, cur, ref, d_blocks, nblocks, nullptr, 0, nullptr, d_start, d_pmv, p, filter_mode, tx_size,
                         tx_type, nullptr, nullptr, d_best, d_coeffs);
}
