// stream_frames.cpp - what an encoder thread does with the frame pipe, through the C ABI only (no torch, no
// Python): create a context and a pipe, make the search patterns resident, then push one frame per call and
// read back the winners, the per-block eob / tx-domain distortion and the packed quantized coefficients.
//
//   g++ -std=c++17 -O2 examples/stream_frames.cpp -o stream_frames -Lrav1e_b200 -lb200rdo -Wl,-rpath,$PWD/rav1e_b200
//   ./stream_frames [frames]            (synthetic 1080p frames; prints frames/s and candidate blocks/s)
//
// The same calls from Rust are declared in INTEGRATION.md section 3.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/b200rdo.h"

#define CHECK(call)                                                                       \
  do {                                                                                    \
    const int st_ = (call);                                                               \
    if (st_ != B200_OK) {                                                                 \
      std::fprintf(stderr, "%s failed (%d): %s\n", #call, st_, b200_last_error(ctx));     \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

int main(int argc, char **argv) {
  const int nframes = argc > 1 ? std::atoi(argv[1]) : 64;
  const int W = 1920, H = 1080, SAD_PER_BLOCK = 64, SATD_PER_BLOCK = 8;
  b200_ctx *ctx = nullptr;
  if (b200_ctx_create(0, &ctx) != B200_OK) {
    std::fprintf(stderr, "no CUDA device: %s\n", b200_last_error(nullptr));
    return 1;
  }
  b200_frame_pipe_cfg cfg{};
  cfg.width = W, cfg.height = H, cfg.pad = 96, cfg.bpp = 1, cfg.bit_depth = 8;
  cfg.block_w = cfg.block_h = 16;
  cfg.lambda = 6400;
  cfg.sad_per_block = SAD_PER_BLOCK, cfg.satd_per_block = SATD_PER_BLOCK;
  cfg.window_hint_px = 64;                 // every offset below stays within +-64 px
  cfg.tx_size = 2, cfg.tx_type = 0;        // TX_16X16, DCT_DCT of the SAD winner's residual
  cfg.dc_quant = 88, cfg.ac_quant = 100;   // dc_q / ac_q of the frame's qindex
  b200_frame_pipe *pipe = nullptr;
  CHECK(b200_frame_pipe_create(ctx, &cfg, &pipe));
  const size_t nb = b200_frame_pipe_nblocks(pipe);

  // the search patterns: (row, col) full-pel offsets per candidate, uploaded once
  std::vector<int8_t> sad_offs(nb * SAD_PER_BLOCK * 2), satd_offs(nb * SATD_PER_BLOCK * 2);
  uint32_t seed = 12345;
  auto rnd = [&seed](int range) {  // xorshift: a stand-in for predictor + diamond / hexagon patterns
    seed ^= seed << 13, seed ^= seed >> 17, seed ^= seed << 5;
    return (int)(seed % (uint32_t)(2 * range + 1)) - range;
  };
  for (auto &o : sad_offs) o = (int8_t)rnd(64);
  for (auto &o : satd_offs) o = (int8_t)rnd(2);
  CHECK(b200_frame_pipe_set_lists(pipe, sad_offs.data(), satd_offs.data(), nullptr));

  // per-frame buffers (an encoder would pin them: cudaHostRegister / cudaHostAlloc)
  std::vector<uint8_t> frame((size_t)W * H);
  std::vector<b200_me_result> best_sad(nb), best_satd(nb);
  std::vector<uint16_t> eob(nb);
  std::vector<uint64_t> tx_dist(nb);
  std::vector<int16_t> packed(nb * 64);
  size_t total_coeffs = 0, count = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (int f = 0; f < nframes; f++) {
    for (int y = 0; y < H; y++)       // a smooth pattern that drifts one pixel per frame
      for (int x = 0; x < W; x++) frame[(size_t)y * W + x] = (uint8_t)(128 + 60 * (((x + f) >> 5) & 1) + ((y >> 4) & 15));
    CHECK(b200_frame_pipe_push_packed(pipe, frame.data(), W, best_sad.data(), best_satd.data(), eob.data(),
                                      tx_dist.data(), packed.data(), packed.size(), &count));
    total_coeffs += count;   // the first push only uploads (count == 0): there is no reference frame yet
  }
  CHECK(b200_ctx_synchronize(ctx));
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("%d frames in %.3f s: %.0f frames/s, %.2f G candidate blocks/s, %.1f coefficients per block; "
              "block 0 of the last frame: mv (%d, %d) sad %u eob %u\n",
              nframes, s, nframes / s, (nframes - 1) * (double)nb * (SAD_PER_BLOCK + SATD_PER_BLOCK + 1) / s / 1e9,
              total_coeffs / ((nframes - 1.0) * nb), best_sad[0].mv_row, best_sad[0].mv_col, best_sad[0].sad, eob[0]);
  b200_frame_pipe_destroy(pipe);
  b200_ctx_destroy(ctx);
  return 0;
}
