// inv_txfm_dev.cuh — per-thread row / column passes of the inverse transform (host + device, so that
// tests/cpp/inv_txfm_emul.cu can replay them on the CPU), shared by inv_txfm.cu and encode_tx.cu.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>

#include "common.cuh"

namespace {

const int kItxW[19] = {4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64};
const int kItxH[19] = {4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16};
// inverse.rs:1710-1711
const int kInterShift[19] = {0, 1, 2, 2, 2, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2};
// get_1d_tx_types (transform/mod.rs:342-362): {column, row}; 0 DCT 1 ADST 2 FLIPADST 3 IDTX 4 WHT
const int kTx1D[17][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}, {2, 0}, {0, 2}, {2, 2}, {1, 2}, {2, 1},
                          {3, 3}, {0, 3}, {3, 0}, {1, 3}, {3, 1}, {2, 3}, {3, 2}, {4, 4}};

// i32 arithmetic wraps in Rust release builds
#define WADD(a, b) ((int)((unsigned)(a) + (unsigned)(b)))
#define WSUB(a, b) ((int)((unsigned)(a) - (unsigned)(b)))
#define WMUL(a, b) ((int)((unsigned)(a) * (unsigned)(b)))
#define WNEG(a) ((int)(0u - (unsigned)(a)))

// Everything below the launch is host+device so that tests/cpp/test_inv_txfm_emul.cu can run the
// very same row / column passes on the CPU against the oracle (the product never does).
#define B200_HDI __host__ __device__ __forceinline__

// transform/mod.rs:296-307 with INV_COS_BIT = 12
B200_HDI int half_btf12(int w0, int in0, int w1, int in1) {
  return WADD(WADD(WMUL(w0, in0), WMUL(w1, in1)), 1 << 11) >> 12;
}
#define HALF_BTF(w0, in0, w1, in1) half_btf12((w0), (in0), (w1), (in1))

// transform/mod.rs:310-315
B200_HDI int clamp_value(int v, int bit) {
  const int hi = (int)((1ll << (bit - 1)) - 1), lo = (int)(-(1ll << (bit - 1)));
  return v < lo ? lo : v > hi ? hi : v;
}
#define CLAMP_VALUE(v, range) clamp_value((v), (range))

B200_HDI int round_shift_i(int v, int b) { return WADD(v, (1 << b) >> 1) >> b; }
#define ROUND_SHIFT(v, b) round_shift_i((v), (b))

#include "inv_txfm_networks.cuh"

constexpr int kSqrt2 = 5793, kInvSqrt2 = 2896;  // transform/mod.rs:47-49 (12 fractional bits)

// INV_TXFM_FNS[kind][log2(N) - 2] (inverse.rs:1593-1623) on a register array
template <int N>
B200_HDI void run_inv_1d(int kind, const int (&in)[N], int (&out)[N], int range) {
  if (kind == 3) {  // identity: sqrt2, 2, 2 sqrt2, 4
#pragma unroll
    for (int i = 0; i < N; i++) {
      if (N == 4) out[i] = round_shift_i(WMUL(kSqrt2, in[i]), 12);
      else if (N == 8) out[i] = WMUL(2, in[i]);
      else if (N == 16) out[i] = round_shift_i(WMUL(2 * kSqrt2, in[i]), 12);
      else out[i] = WMUL(4, in[i]);
    }
    return;
  }
  if constexpr (N == 4) {
    if (kind == 0) {
      d_av1_idct4(in, out, range);
    } else if (kind == 4) {  // av1_iwht4, inverse.rs:35-53
      const int s0 = WADD(in[0], in[1]);
      const int s2 = WSUB(in[2], in[3]);
      const int s4 = WSUB(s0, s2) >> 1;
      const int s3 = WSUB(s4, in[3]);
      const int s1 = WSUB(s4, in[1]);
      out[0] = WSUB(s0, s3);
      out[1] = s3;
      out[2] = s1;
      out[3] = WADD(s2, s1);
    } else {
      d_av1_iadst4(in, out, range);
    }
  } else if constexpr (N == 8) {
    if (kind == 0) d_av1_idct8(in, out, range);
    else d_av1_iadst8(in, out, range);
  } else if constexpr (N == 16) {
    if (kind == 0) d_av1_idct16(in, out, range);
    else d_av1_iadst16(in, out, range);
  } else if constexpr (N == 32) {
    d_av1_idct32(in, out, range);
  } else {
    d_av1_idct64(in, out, range);
  }
  if (kind == 2) {  // flipped ADST = reversed ADST output
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
      const int t = out[i];
      out[i] = out[N - 1 - i];
      out[N - 1 - i] = t;
    }
  }
}

struct InvArgs {
  const void *coeffs;  // n x (min(W,32) * min(H,32)), index = col * min(H,32) + row
  void *dst;           // plane pixel (0,0)
  int dst_stride;      // elements
  const b200_block *blocks;
  size_t n;
  int row_kind, col_kind, inter_shift, bd, rect, lossless;
};

constexpr int kInvThreads = 128;

// Row pass of one block-transform for "thread" t < H (inverse.rs:1659-1682): row t of the coded
// coefficients -> scaling + clamp -> 1-D row transform -> row t of the tile.  Rows >= 32 of a
// 64-point block stay zero.
template <int W, int H, typename CoefT>
B200_HDI void inv_row_pass_ptr(const InvArgs &a, const CoefT *block_coeffs, int t, int *tile) {
  constexpr int PITCH = W + 1;
  constexpr int W32 = W < 32 ? W : 32, H32 = H < 32 ? H : 32;
  int out[W];
  if (t < H32) {
    int in[W];
    const CoefT *src = block_coeffs + t;
    const int range = a.bd + 8;
#pragma unroll
    for (int c = 0; c < W; c++) {
      int val = 0;
      if (c < W32) {
        const int raw = (int)src[(size_t)c * H32];
        val = a.rect ? round_shift_i(WMUL(raw, kInvSqrt2), 12) : a.lossless ? raw >> 2 : raw;
        val = clamp_value(val, range);
      }
      in[c] = val;
    }
    run_inv_1d<W>(a.row_kind, in, out, range);
  } else {
#pragma unroll
    for (int c = 0; c < W; c++) out[c] = 0;
  }
#pragma unroll
  for (int c = 0; c < W; c++) tile[t * PITCH + c] = out[c];
}

template <int W, int H, typename CoefT>
B200_HDI void inv_row_pass(const InvArgs &a, size_t blk, int t, int *tile) {
  constexpr int W32 = W < 32 ? W : 32, H32 = H < 32 ? H : 32;
  inv_row_pass_ptr<W, H, CoefT>(a, (const CoefT *)a.coeffs + blk * (size_t)(W32 * H32), t, tile);
}

// Column pass for "thread" t < W (inverse.rs:1684-1703): column t of the tile -> intermediate
// shift + clamp -> 1-D column transform -> final shift -> add into the plane with the pixel clamp.
template <int W, int H, typename Px>
B200_HDI void inv_col_pass_ptr(const InvArgs &a, Px *block_origin, int t, const int *tile) {
  constexpr int PITCH = W + 1;
  const int range = a.bd + 6 > 16 ? a.bd + 6 : 16;
  int in[H], out[H];
#pragma unroll
  for (int r = 0; r < H; r++) in[r] = clamp_value(round_shift_i(tile[r * PITCH + t], a.inter_shift), range);
  run_inv_1d<H>(a.col_kind, in, out, range);
  Px *p = block_origin + t;
  const int maxv = (1 << a.bd) - 1;
#pragma unroll
  for (int r = 0; r < H; r++) {
    const int rr = a.lossless ? out[r] : round_shift_i(out[r], 4);
    const int v = WADD((int)p[(long long)r * a.dst_stride], rr);
    p[(long long)r * a.dst_stride] = (Px)(v < 0 ? 0 : v > maxv ? maxv : v);
  }
}

template <int W, int H, typename Px>
B200_HDI void inv_col_pass(const InvArgs &a, size_t blk, int t, const int *tile) {
  const b200_block b = a.blocks[blk];
  inv_col_pass_ptr<W, H, Px>(a, (Px *)a.dst + (long long)b.y * a.dst_stride + b.x, t, tile);
}

// INV_TXFM_FNS holes (`unimplemented!()`), inverse.rs:1593-1623
inline bool inv_1d_exists(int kind, int n) {
  if (kind == 0) return true;
  if (kind == 1 || kind == 2) return n <= 16;
  if (kind == 3) return n <= 32;
  return n == 4;
}

// everything but the buffers of one (tx_size, tx_type, bit depth)
inline void inv_setup(int tx_size, int tx_type, int bd, InvArgs *a) {
  a->row_kind = kTx1D[tx_type][1];
  a->col_kind = kTx1D[tx_type][0];
  a->inter_shift = kInterShift[tx_size];
  a->bd = bd;
  int lw = 0, lh = 0;
  while ((1 << lw) < kItxW[tx_size]) lw++;
  while ((1 << lh) < kItxH[tx_size]) lh++;
  a->rect = (lw - lh == 1 || lw - lh == -1) ? 1 : 0;  // rect_type.abs() == 1, inverse.rs:1672
  a->lossless = tx_type == 16;
}

}  // namespace
