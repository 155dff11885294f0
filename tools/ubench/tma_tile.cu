// Stand-alone probe of the TMA tile load used by polar_tma_kernel (cp.async.bulk.tensor.3d + mbarrier, OOB zero fill).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/ubench/tma_tile tools/ubench/tma_tile.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int x0, int y0, int b, int pitch, int rows) {
  extern __shared__ __align__(128) float box[];
  __shared__ __align__(8) unsigned long long mbar;
  const unsigned bar = smem_u32(&mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(pitch * rows * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(box)), "l"(&map), "r"(x0), "r"(y0), "r"(b), "r"(bar) : "memory");
  }
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
  for (int i = threadIdx.x; i < pitch * rows; i += blockDim.x) out[i] = box[i];
}
int main() {
  const int W = 640, H = 480, B = 2, pitch = 52, rows = 47;
  std::vector<float> h((size_t)W * H * B);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, pitch * rows * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = (PFN_cuTensorMapEncodeTiled)fn;
  CUtensorMap map;
  cuuint64_t dims[3] = {W, H, B}; cuuint64_t str[2] = {W * 4ull, (cuuint64_t)W * H * 4ull};
  cuuint32_t box[3] = {pitch, rows, 1}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d entry %d\n", (int)r, (int)q);
  for (int t = 0; t < 3; ++t) {
    const int x0 = t == 0 ? 100 : (t == 1 ? -4 : 620), y0 = t == 0 ? 50 : (t == 1 ? -2 : 470), b = t % 2;
    k<<<1, 128, pitch * rows * 4>>>(map, o, x0, y0, b, pitch, rows);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> g(pitch * rows);
    cudaMemcpy(g.data(), o, g.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < rows; ++y) for (int x = 0; x < pitch; ++x) {
      const int gx = x0 + x, gy = y0 + y;
      const float want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0.f : h[(size_t)b * W * H + (size_t)gy * W + gx];
      bad += g[y * pitch + x] != want;
    }
    printf("tile %d: %s, mismatches %d\n", t, cudaGetErrorString(e), bad);
  }
  return 0;
}
