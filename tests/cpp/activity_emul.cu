// Host replay of the activity-mask kernel's per-thread variance function
// (rav1e_b200/csrc/rdo_dist.cu: variance_8x8_px, a __host__ __device__ function) for the CPU test
// suite (tests/test_activity.py).  Compiled by the test into a scratch directory; not product code.
#include "../../rav1e_b200/csrc/common.cuh"

char g_b200_last_error[512];  // lives in capi.cu in the product library
b200_ctx *b200_default_ctx() { return nullptr; }  // per-call forms are not exercised here
extern "C" const char *b200_last_error(const b200_ctx *) { return g_b200_last_error; }

#include "../../rav1e_b200/csrc/rdo_dist.cu"

extern "C" void emul_variances(const void *luma, long long stride, int bpp, int wb, int hb, uint32_t *out) {
  for (int y = 0; y < hb; y++)
    for (int x = 0; x < wb; x++)
      out[y * wb + x] = bpp == 1 ? variance_8x8_px<uint8_t>((const uint8_t *)luma + (long long)(8 * y) * stride + 8 * x, stride)
                                 : variance_8x8_px<uint16_t>((const uint16_t *)luma + (long long)(8 * y) * stride + 8 * x, stride);
}
