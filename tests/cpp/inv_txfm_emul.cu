// Host emulation of inv_txfm_add_kernel for the CPU test suite (tests/test_inv_txfm_emul.py).
//
// The inverse-transform kernel keeps its per-thread work in two host+device functions
// (inv_row_pass / inv_col_pass in rav1e_b200/csrc/inv_txfm.cu).  This harness includes that very
// source file and replays the kernel's schedule on the CPU — for every block, every "thread" t runs
// the row pass into a tile of the kernel's geometry, then every t runs the column pass — so the
// generated butterfly networks, the coefficient indexing, the scalings, clamps and the pixel add of
// the CUDA source are checked against the oracle without a GPU.  Not part of the product: it is
// compiled by the test into a scratch directory, never into libb200rdo.so.
#include "../../rav1e_b200/csrc/common.cuh"

char g_b200_last_error[512];  // lives in capi.cu in the product library

#include "../../rav1e_b200/csrc/inv_txfm.cu"

#include <vector>

namespace {

template <int W, int H, typename CoefT, typename Px>
void emulate(const InvArgs &a) {
  constexpr int PITCH = W + 1;
  constexpr int REGION = H * PITCH + ((H * PITCH) % 2 == 0 ? 1 : 0);
  std::vector<int> tile(REGION);
  for (size_t blk = 0; blk < a.n; blk++) {
    std::fill(tile.begin(), tile.end(), 0x5a5a5a5a);  // shared memory is not zero-initialised
    for (int t = 0; t < H; t++) inv_row_pass<W, H, CoefT>(a, blk, t, tile.data());
    for (int t = 0; t < W; t++) inv_col_pass<W, H, Px>(a, blk, t, tile.data());
  }
}

}  // namespace

// host pointers everywhere; returns 0 on success, -1 for pairs the reference leaves unimplemented
extern "C" int emul_inverse_transform_add(const void *coeffs, void *dst, int dst_stride, int bpp,
                                          const b200_block *blocks, size_t n, int tx_size, int tx_type,
                                          int bd) {
  const int w = kItxW[tx_size], h = kItxH[tx_size];
  if (!inv_1d_exists(kTx1D[tx_type][1], w) || !inv_1d_exists(kTx1D[tx_type][0], h)) return -1;
  InvArgs a{};
  a.coeffs = coeffs;
  a.dst = dst;
  a.dst_stride = dst_stride;
  a.blocks = blocks;
  a.n = n;
  a.row_kind = kTx1D[tx_type][1];
  a.col_kind = kTx1D[tx_type][0];
  a.inter_shift = kInterShift[tx_size];
  a.bd = bd;
  int lw = 0, lh = 0;
  while ((1 << lw) < w) lw++;
  while ((1 << lh) < h) lh++;
  a.rect = (lw - lh == 1 || lw - lh == -1) ? 1 : 0;
  a.lossless = tx_type == 16;
  const bool hbd = bpp == 2;
  switch (tx_size) {
#define EMUL(ID, W_, H_)                                          \
  case ID:                                                        \
    if (hbd)                                                      \
      emulate<W_, H_, int32_t, uint16_t>(a);                      \
    else                                                          \
      emulate<W_, H_, int16_t, uint8_t>(a);                       \
    return 0;
    EMUL(0, 4, 4)
    EMUL(1, 8, 8)
    EMUL(2, 16, 16)
    EMUL(3, 32, 32)
    EMUL(4, 64, 64)
    EMUL(5, 4, 8)
    EMUL(6, 8, 4)
    EMUL(7, 8, 16)
    EMUL(8, 16, 8)
    EMUL(9, 16, 32)
    EMUL(10, 32, 16)
    EMUL(11, 32, 64)
    EMUL(12, 64, 32)
    EMUL(13, 4, 16)
    EMUL(14, 16, 4)
    EMUL(15, 8, 32)
    EMUL(16, 32, 8)
    EMUL(17, 16, 64)
    EMUL(18, 64, 16)
#undef EMUL
  }
  return -2;
}
