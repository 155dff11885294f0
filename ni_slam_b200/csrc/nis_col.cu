// nis_col.cu -- column-pass kernels (transform along image rows; r2c forward / c2r inverse), sm_100a.
// One CTA = 16 adjacent real columns of one image (8 complex lines), grid = (W/16, batch).
#include "nis_device.cuh"
#include "nis_internal.h"
#include "nis_sizes.h"
#include "nis_warp.cuh"

namespace nis {

// resident CTAs per SM the register budget of the column kernels is sized for: 768 threads (85 registers each): the paired
// radix-12 stages hold 24 complex values per thread; a 64-register budget spills there (measured -6 %, profiles/ab_r02.md)
template <class Pro, class = void> struct ProTraits { static constexpr bool kSmemLut = false; };
template <class Pro> struct ProTraits<Pro, decltype((void)Pro::kSmemLut)> { static constexpr bool kSmemLut = Pro::kSmemLut; };
template <class Pro> __device__ __forceinline__ auto bind_pro(const Pro& pro, int b, int c0, const float* lut_s) {
  if constexpr (ProTraits<Pro>::kSmemLut) return pro.bind(b, c0, lut_s);
  else return pro.bind(b, c0);
}

// Two-stage plans (B == 1: one paired radix-A stage and one radix-C stage, e.g. 480 = 15 x 32) hold up to 32 complex values per
// thread and are sized for 480 resident threads (136 registers).
#ifndef NIS_COL_MINB
#define NIS_COL_MINB (B == 1 ? (480 / T > 0 ? 480 / T : 1) : 768 / T)
#endif

template <int N, int A, int B, int C, int T, class Pro, int LN>
__global__ void __launch_bounds__(T, NIS_COL_MINB) col_fwd_kernel(Pro pro, Twiddles twd, Dst<cpx> out, int W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* smem = reinterpret_cast<cpx*>(smem_raw);
  const int b = blockIdx.y, c0 = blockIdx.x * (2 * LN), tid = threadIdx.x;
  const float* lut_s = nullptr;
  if constexpr (ProTraits<Pro>::kSmemLut) {          // gather prologues on u8 images: the u8 -> f32/255 table comes from shared memory
    __shared__ float lut_sh[256];
    for (int i = tid; i < 256; i += T) lut_sh[i] = __ldg(pro.lut + i);
    __syncthreads();
    lut_s = lut_sh;
  }
  const auto bp = bind_pro(pro, b, c0, lut_s);
  col_fwd_stage_a<N, A, B, C, T, decltype(bp), LN>(tid, smem, twd, bp);
  __syncthreads();
  if constexpr (B > 1) {
    col_stage_b<N, A, B, C, T, false, false, LN>(tid, smem, twd);
    __syncthreads();
  }
  col_fwd_stage_c<N, A, B, C, T, LN>(tid, smem, out.at(b), W, c0);
}

template <int N, int A, int B, int C, int T, class Epi>
__global__ void __launch_bounds__(T, NIS_COL_MINB) col_inv_kernel(Src<cpx> in, Twiddles twd, Epi epi, int W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* smem = reinterpret_cast<cpx*>(smem_raw);
  const int b = blockIdx.y, c0 = blockIdx.x * kColTile, tid = threadIdx.x;
  col_inv_stage_a<N, A, B, C, T>(tid, smem, twd, in.at(b), W, c0);
  __syncthreads();
  if constexpr (B > 1) {
    col_stage_b<N, A, B, C, T, true, false>(tid, smem, twd);
    __syncthreads();
  }
  auto be = epi.bind(b, c0);
  col_inv_stage_c<N, A, B, C, T>(tid, smem, be);
  DeviceSync sync;
  be.finish(tid, sync);
}

// fused inverse column pass -> kernel function -> forward column pass (correlation_flow.cc:212-215 / :222-225 between
// the IFFT and the FFT): the real kernel image exists only in registers, between the two radix-C butterflies of one thread.
// Four barriers, four shared-memory round trips.
template <int N, int A, int B, int C, int T>
__global__ void __launch_bounds__(T, NIS_COL_MINB) colcol_kernel(Src<cpx> in, Dst<cpx> out, Twiddles twd, KernelFn kfn, int W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* smem = reinterpret_cast<cpx*>(smem_raw);
  const int b = blockIdx.y, c0 = blockIdx.x * kColTile, tid = threadIdx.x;
  col_inv_stage_a<N, A, B, C, T>(tid, smem, twd, in.at(b), W, c0);
  __syncthreads();
  if constexpr (B > 1) {
    col_stage_b<N, A, B, C, T, true, false>(tid, smem, twd);
    __syncthreads();
  }
  auto fn = kfn.bind(b);
  col_inv_fn_fwd_stage_c<N, A, B, C, T>(tid, smem, fn);
  __syncthreads();
  if constexpr (B > 1) {
    col_stage_b<N, A, B, C, T, false, true>(tid, smem, twd);
    __syncthreads();
  }
  col_fwd_dit_stage_a<N, A, B, C, T>(tid, smem, twd, out.at(b), W, c0);
  DeviceSync sync;
  fn.finish(tid, sync);
}

template <int N, int A, int B, int C, int T>
static int run_colcol(Twiddles tw, Src<cpx> in, Dst<cpx> out, KernelFn fn, int W, int Bn, cudaStream_t s) {
  auto k = colcol_kernel<N, A, B, C, T>;
  const size_t smem = ColGeom<N, A, B, C, T>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  k<<<dim3(W / kColTile, Bn), T, smem, s>>>(in, out, tw, fn, W);
  return (int)cudaGetLastError();
}

template <int N, int A, int B, int C, int T, int LN = kColLanes, class Pro>
static int run_col_fwd(Twiddles tw, Pro pro, Dst<cpx> out, int W, int Bn, cudaStream_t s) {
  auto k = col_fwd_kernel<N, A, B, C, T, Pro, LN>;
  const size_t smem = ColGeom<N, A, B, C, T, LN>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  if (W % (2 * LN)) return -1;
  k<<<dim3(W / (2 * LN), Bn), T, smem, s>>>(pro, tw, out, W);
  return (int)cudaGetLastError();
}
template <int N, int A, int B, int C, int T, class Epi>
static int run_col_inv(Twiddles tw, Src<cpx> in, Epi epi, int W, int Bn, cudaStream_t s) {
  auto k = col_inv_kernel<N, A, B, C, T, Epi>;
  const size_t smem = ColGeom<N, A, B, C, T>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  k<<<dim3(W / kColTile, Bn), T, smem, s>>>(in, tw, epi, W);
  return (int)cudaGetLastError();
}

bool col_size_supported(int N) {
#define X(n, a, b, c, t) if (N == n) return true;
  NIS_COL_PLANS(X)
#undef X
  return false;
}
void plan_radices_col(int N, int r[3]) {
#define X(n, a, b, c, t) if (N == n) { r[0] = a; r[1] = b; r[2] = c; return; }
  NIS_COL_PLANS(X)
#undef X
  r[0] = r[1] = r[2] = 0;
}

#define FWD_DISPATCH(PRO)                                                                          \
  if (B <= 0) return 0;                                                                            \
  switch (N) {                                                                                     \
    NIS_COL_PLANS(FWD_CASE)                                                                        \
    default: return -1;                                                                            \
  }
#define FWD_CASE(n, a, b, c, t) case n: return run_col_fwd<n, a, b, c, t>(tw, pro, out, W, B, s);
int launch_col_fwd_f32(int N, Twiddles tw, ProRealF32 pro, Dst<cpx> out, int W, int B, cudaStream_t s) { FWD_DISPATCH(ProRealF32) }
int launch_col_fwd_u8(int N, Twiddles tw, ProRealU8 pro, Dst<cpx> out, int W, int B, cudaStream_t s) { FWD_DISPATCH(ProRealU8) }
#undef FWD_CASE
// Lane count of the forward pass with the rotation prologue.  16 lanes (32 adjacent columns per CTA) make the gather itself faster
// (col_fwd_rotate 1.71 -> 1.38 ms per 1000 frames) but the whole step SLOWER (73.8 k -> 66.2 k solves/s, profiles/ab_r02.md): three
// 69 KB tiles per SM push the SM's shared-memory carveout to its maximum, and every kernel of the other lanes that shares the SM then
// runs with the smallest L1 (the library-wide carveout experiment shows what that costs: +10-20 % per kernel).  Kept at 8.
#ifndef NIS_ROT_LANES
#define NIS_ROT_LANES 8
#endif
constexpr int rot_lanes(int n) { return n <= 512 ? NIS_ROT_LANES : kColLanes; }
#define FWD_CASE(n, a, b, c, t) case n: return run_col_fwd<n, a, b, c, t, rot_lanes(n)>(tw, pro, out, W, B, s);
int launch_col_fwd_rotate(int N, Twiddles tw, RotateArgs ra, Dst<cpx> out, int W, int B, cudaStream_t s) {
  if (ra.is_u8) {
    ProRotate<true> pro{ra.f32, ra.u8, ra.lut, ra.H, ra.W, ra.mats, ra.sel, ra.rowtab, ra.polar, ra.D, ra.loop};
    FWD_DISPATCH(ProRotate<true>)
  } else {
    ProRotate<false> pro{ra.f32, ra.u8, ra.lut, ra.H, ra.W, ra.mats, ra.sel, ra.rowtab, ra.polar, ra.D, ra.loop};
    FWD_DISPATCH(ProRotate<false>)
  }
}
#undef FWD_CASE

#define INV_DISPATCH                                                                               \
  if (B <= 0) return 0;                                                                            \
  switch (N) {                                                                                     \
    NIS_COL_PLANS(INV_CASE)                                                                        \
    default: return -1;                                                                            \
  }
#define INV_CASE(n, a, b, c, t) case n: return run_col_inv<n, a, b, c, t>(tw, in, epi, W, B, s);
int launch_col_inv_store(int N, Twiddles tw, Src<cpx> in, EpiStore epi, int W, int B, cudaStream_t s) { INV_DISPATCH }
int launch_col_inv_store_shift(int N, Twiddles tw, Src<cpx> in, EpiStoreShift epi, int W, int B, cudaStream_t s) { INV_DISPATCH }
int launch_col_inv_peak(int N, Twiddles tw, Src<cpx> in, EpiPeak epi, int W, int B, cudaStream_t s) { INV_DISPATCH }
#undef INV_CASE

int launch_colcol(int N, Twiddles tw, Src<cpx> in, Dst<cpx> out, KernelFn fn, int W, int B, cudaStream_t s) {
  if (B <= 0) return 0;
  switch (N) {
#define X(n, a, b, c, t) case n: return run_colcol<n, a, b, c, t>(tw, in, out, fn, W, B, s);
    NIS_COL_PLANS(X)
#undef X
    default: return -1;
  }
}

}  // namespace nis
