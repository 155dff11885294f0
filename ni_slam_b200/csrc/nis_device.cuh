// nis_device.cuh -- device-only helpers: block reductions used by the epilogues' finish() step.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "nis_ops.cuh"

#ifndef NIS_SMEM_CARVEOUT_DEFAULT
#define NIS_SMEM_CARVEOUT_DEFAULT -1
#endif

namespace nis {

struct DeviceSync {
  // max|k| of the CTA -> atomicMax on the per-image scalar (non-negative floats order like their bit patterns)
  __device__ __forceinline__ void block_max_to(unsigned int* p, float mx, int tid) {
    __shared__ float red[32];
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int warp = tid >> 5, lane = tid & 31, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (warp == 0) {
      float v = lane < nw ? red[lane] : 0.f;
#pragma unroll
      for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
      if (lane == 0) atomicMax(p, __float_as_uint(v));
    }
  }
  // arg-max key + sum + sum of squares of the CTA -> one 64-bit atomicMax and two double atomicAdds
  __device__ __forceinline__ void block_peak_to(PeakStats* st, unsigned long long key, double s, double q, int tid) {
    __shared__ unsigned long long rk[32];
    __shared__ double rs[32], rq[32];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
      key = k2 > key ? k2 : key;
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const int warp = tid >> 5, lane = tid & 31, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) { rk[warp] = key; rs[warp] = s; rq[warp] = q; }
    __syncthreads();
    if (warp == 0) {
      key = lane < nw ? rk[lane] : 0ull;
      s = lane < nw ? rs[lane] : 0.0;
      q = lane < nw ? rq[lane] : 0.0;
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
        key = k2 > key ? k2 : key;
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (lane == 0) {
        atomicMax(&st->key, key);
        atomicAdd(&st->sum, s);
        atomicAdd(&st->sumsq, q);
      }
    }
  }
};

// Shared-memory carveout override for A/B runs: NIS_SMEM_CARVEOUT = percent of the maximum for EVERY kernel of the library (-1, the
// default, leaves the driver's per-kernel choice).  Measured (profiles/ab_r02.md): 100 % costs every FFT kernel 10-20 % (the twiddle
// and table loads want the L1), 75 % equals the default, 50 % loses occupancy -- the driver's choice stays.
inline int smem_carveout() {
  static const int pct = [] {
    const char* e = getenv("NIS_SMEM_CARVEOUT");
    return e ? atoi(e) : NIS_SMEM_CARVEOUT_DEFAULT;
  }();
  return pct;
}
// kernels without a dynamic shared-memory request: the carveout alone (call once per kernel: `static int a = set_carveout(k);`)
template <class K> inline int set_carveout(K kernel) {
  const int pct = smem_carveout();
  return pct >= 0 ? (int)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct) : 0;
}
template <class K> inline int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  const int pct = smem_carveout();
  if (pct >= 0) return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  return 0;
}

}  // namespace nis
