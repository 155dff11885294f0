// nis_col.cu -- column-pass kernels (transform along image rows; r2c forward / c2r inverse), sm_100a.
// One CTA = 32 adjacent real columns of one image (16 complex lines), grid = (W/32, batch).
#include "nis_device.cuh"
#include "nis_internal.h"
#include "nis_sizes.h"

namespace nis {

template <int N, int R0, int R1, int R2, int T, class Pro>
__global__ void __launch_bounds__(T) col_fwd_kernel(Pro pro, Twiddles twd, Dst<cpx> out, int W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* smem = reinterpret_cast<cpx*>(smem_raw);
  const int b = blockIdx.y, c0 = blockIdx.x * 32, tid = threadIdx.x;
  const auto bp = pro.bind(b, c0);
  col_fwd_phase0<N, R0, R1, R2, T>(tid, smem, bp);
  __syncthreads();
  CarryRegs<R1, ColGeom<N, R0, R1, R2, T>::ROUNDS1> st;
  col_stage1_read<N, R0, R1, R2, T, false>(tid, smem, twd, st);
  __syncthreads();
  col_stage1_write<N, R0, R1, R2, T, false>(tid, smem, st);
  __syncthreads();
  col_fwd_phase2<N, R0, R1, R2, T>(tid, smem, twd, out.at(b), W, c0);
}

template <int N, int R0, int R1, int R2, int T, class Epi>
__global__ void __launch_bounds__(T) col_inv_kernel(Src<cpx> in, Twiddles twd, Epi epi, int W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* smem = reinterpret_cast<cpx*>(smem_raw);
  const int b = blockIdx.y, c0 = blockIdx.x * 32, tid = threadIdx.x;
  col_inv_phase0<N, R0, R1, R2, T>(tid, smem, in.at(b), W, c0);
  __syncthreads();
  CarryRegs<R1, ColGeom<N, R0, R1, R2, T>::ROUNDS1> st;
  col_stage1_read<N, R0, R1, R2, T, true>(tid, smem, twd, st);
  __syncthreads();
  col_stage1_write<N, R0, R1, R2, T, true>(tid, smem, st);
  __syncthreads();
  auto be = epi.bind(b, c0);
  col_inv_phase2<N, R0, R1, R2, T>(tid, smem, twd, be);
  DeviceSync sync;
  be.finish(tid, sync);
}

template <int N, int R0, int R1, int R2, int T, class Pro>
static int run_col_fwd(Twiddles tw, Pro pro, Dst<cpx> out, int W, int B, cudaStream_t s) {
  auto k = col_fwd_kernel<N, R0, R1, R2, T, Pro>;
  const size_t smem = ColGeom<N, R0, R1, R2, T>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  k<<<dim3(W / 32, B), T, smem, s>>>(pro, tw, out, W);
  return (int)cudaGetLastError();
}
template <int N, int R0, int R1, int R2, int T, class Epi>
static int run_col_inv(Twiddles tw, Src<cpx> in, Epi epi, int W, int B, cudaStream_t s) {
  auto k = col_inv_kernel<N, R0, R1, R2, T, Epi>;
  const size_t smem = ColGeom<N, R0, R1, R2, T>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  k<<<dim3(W / 32, B), T, smem, s>>>(in, tw, epi, W);
  return (int)cudaGetLastError();
}

bool col_size_supported(int N) {
#define X(n, f0, f1, f2, i0, i1, i2, t) if (N == n) return true;
  NIS_COL_PLANS(X)
#undef X
  return false;
}
void plan_radices_col(int N, bool inverse, int r[3]) {
#define X(n, f0, f1, f2, i0, i1, i2, t) \
  if (N == n) { if (inverse) { r[0] = i0; r[1] = i1; r[2] = i2; } else { r[0] = f0; r[1] = f1; r[2] = f2; } return; }
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
#define FWD_CASE(n, f0, f1, f2, i0, i1, i2, t) case n: return run_col_fwd<n, f0, f1, f2, t>(tw, pro, out, W, B, s);
int launch_col_fwd_f32(int N, Twiddles tw, ProRealF32 pro, Dst<cpx> out, int W, int B, cudaStream_t s) { FWD_DISPATCH(ProRealF32) }
int launch_col_fwd_u8(int N, Twiddles tw, ProRealU8 pro, Dst<cpx> out, int W, int B, cudaStream_t s) { FWD_DISPATCH(ProRealU8) }
#undef FWD_CASE

#define INV_DISPATCH                                                                               \
  if (B <= 0) return 0;                                                                            \
  switch (N) {                                                                                     \
    NIS_COL_PLANS(INV_CASE)                                                                        \
    default: return -1;                                                                            \
  }
#define INV_CASE(n, f0, f1, f2, i0, i1, i2, t) case n: return run_col_inv<n, i0, i1, i2, t>(tw, in, epi, W, B, s);
int launch_col_inv_store(int N, Twiddles tw, Src<cpx> in, EpiStore epi, int W, int B, cudaStream_t s) { INV_DISPATCH }
int launch_col_inv_kernel(int N, Twiddles tw, Src<cpx> in, EpiKernel epi, int W, int B, cudaStream_t s) { INV_DISPATCH }
int launch_col_inv_peak(int N, Twiddles tw, Src<cpx> in, EpiPeak epi, int W, int B, cudaStream_t s) { INV_DISPATCH }
#undef INV_CASE

}  // namespace nis
