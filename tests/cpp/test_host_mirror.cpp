// Exercises rav1e_b200/host/rav1e_b200.hpp the way the reference's own unit tests call its
// kernels: get_sad / get_satd on the closed-form planes of src/dist.rs:384-413 with the golden
// values of :418-441 / :477-500, plus forward_transform / put_8tap smoke calls.  Built and run by
// tests/test_host_mirror_gpu.py.  Exit code 0 = all good.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../rav1e_b200/host/rav1e_b200.hpp"

using namespace rav1e_b200;

int main() {
  const int W = 640 + 2 * 136, H = 480 + 2 * 136, W2 = 640 + 2 * 264, H2 = 480 + 2 * 264;
  std::vector<uint8_t> in((size_t)W * H), rec((size_t)W2 * H2);
  const int xpad_off = (136 - 136) - 8;
  for (int i = 0; i < H; i++)
    for (int j = 0; j < W; j++) in[(size_t)i * W + j] = (uint8_t)(((j + i) - xpad_off) & 255);
  for (int i = 0; i < H2; i++)
    for (int j = 0; j < W2; j++) rec[(size_t)i * W2 + j] = (uint8_t)(((j - i) - xpad_off) & 255);
  PlaneRegion<uint8_t> org{in.data() + (size_t)(136 + 40) * W + 136 + 32, W, 128, 128};
  PlaneRegion<uint8_t> ref{rec.data() + (size_t)(264 + 40) * W2 + 264 + 32, W2, 128, 128};
  struct { int w, h; uint32_t sad, satd; } kat[] = {
      {4, 4, 1912, 1408}, {8, 8, 7824, 3984}, {16, 16, 31136, 9984}, {32, 16, 59552, 13760},
      {64, 64, 438912, 84176}, {128, 128, 1689792, 321456}, {16, 4, 6664, 2632}};
  int bad = 0;
  for (auto &k : kat) {
    const uint32_t s = get_sad(org, ref, k.w, k.h, 8, CpuFeatureLevel::CUDA_SM100);
    const uint32_t t = get_satd(org, ref, k.w, k.h, 8, CpuFeatureLevel::CUDA_SM100);
    if (s != k.sad || t != k.satd) {
      std::printf("MISMATCH %dx%d sad %u (want %u) satd %u (want %u)\n", k.w, k.h, s, k.sad, t, k.satd);
      bad++;
    }
  }
  // the mirror refuses non-CUDA levels instead of silently computing on the CPU
  try {
    get_sad(org, ref, 16, 16, 8, CpuFeatureLevel::AVX2);
    std::printf("expected a throw for a non-CUDA level\n");
    bad++;
  } catch (const std::invalid_argument &) {
  }
  // forward_transform: constant 4x4 block through WHT (see tests/test_oracle_txfm.py) -> DC 80
  std::vector<int16_t> x(16, 5), c(16, 0);
  forward_transform<int16_t>(x.data(), c.data(), 4, TxSize::TX_4X4, TxType::WHT_WHT, 8, CpuFeatureLevel::CUDA_SM100);
  if (c[0] != 80) { std::printf("WHT DC %d\n", c[0]); bad++; }
  try {
    forward_transform<int16_t>(x.data(), c.data(), 4, TxSize::TX_64X64, TxType::ADST_ADST, 8, CpuFeatureLevel::CUDA_SM100);
    bad++;
  } catch (const std::invalid_argument &) {
  }
  // put_8tap at an integer position copies
  std::vector<uint8_t> dst(8 * 8, 0);
  PlaneRegionMut<uint8_t> d{dst.data(), 8, 8, 8};
  put_8tap(d, org.data, org.stride, 8, 8, 0, 0, FilterMode::REGULAR, FilterMode::REGULAR, 8, CpuFeatureLevel::CUDA_SM100);
  for (int r = 0; r < 8; r++)
    for (int col = 0; col < 8; col++)
      if (dst[r * 8 + col] != org.data[(size_t)r * org.stride + col]) bad++;
  // compute_rd_cost is the correctly rounded fma of rdo.rs:718-723
  const double lam = 123.456789, want = std::fma(lam, 98765 / 8.0, (double)1234567890123ull);
  if (compute_rd_cost(lam, 98765, 1234567890123ull, CpuFeatureLevel::CUDA_SM100) != want) {
    std::printf("compute_rd_cost mismatch\n");
    bad++;
  }
  std::printf(bad ? "FAILED (%d)\n" : "host mirror ok\n", bad);
  return bad ? 1 : 0;
}
