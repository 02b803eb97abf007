// tma_probe: one 2-D u8 cp.async.bulk.tensor load with the given geometry; prints OK/BAD/ERR.
// usage: tma_probe W H STRIDE BW BH CX CY DST_OFF
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap map, int cx, int cy, int bytes, int dst_off, uint8_t *out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) unsigned long long bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(sm + dst_off);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(d), "l"(&map), "r"(cx), "r"(cy), "r"(b) : "memory");
  }
  for (uint32_t spin = 0;; spin++) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(b), "r"(0) : "memory");
    if (done) break;
    if (spin > (1u << 22)) { if (threadIdx.x == 0) printf("TIMEOUT\n"); return; }
  }
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[dst_off + i];
}
int main(int argc, char **argv) {
  if (argc < 9) return 2;
  long W = atol(argv[1]), H = atol(argv[2]), S = atol(argv[3]);
  int bw = atoi(argv[4]), bh = atoi(argv[5]), cx = atoi(argv[6]), cy = atoi(argv[7]), off = atoi(argv[8]);
  void *f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) { printf("ERR entry\n"); return 1; }
  std::vector<uint8_t> h(S * H);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
  uint8_t *d, *o; cudaMalloc(&d, S * H); cudaMalloc(&o, bw * bh); cudaMemcpy(d, h.data(), S * H, cudaMemcpyHostToDevice);
  CUtensorMap m; cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, st[1] = {(cuuint64_t)S}; cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)f)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("ENCODE_FAIL %d\n", (int)r); return 0; }
  size_t smem = off + bw * bh;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<1, 256, smem>>>(m, cx, cy, bw * bh, off, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ERR %s\n", cudaGetErrorString(e)); return 0; }
  std::vector<uint8_t> g(bw * bh); cudaMemcpy(g.data(), o, bw * bh, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int y = 0; y < bh; y++) for (int x = 0; x < bw; x++) {
    long gx = cx + x, gy = cy + y; uint8_t want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[gy * S + gx] : 0;
    if (g[y * bw + x] != want) bad++;
  }
  printf(bad ? "BAD %ld\n" : "OK\n", bad);
  return 0;
}
