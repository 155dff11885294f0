// nis_col.cu -- column-pass kernels (transform along image rows; r2c forward / c2r inverse), sm_100a.
// One CTA = 32 adjacent real columns of one image (16 complex lines), grid = (W/32, batch).
#include "nis_device.cuh"
#include "nis_internal.h"
#include "nis_sizes.h"
#include "nis_warp.cuh"

namespace nis {

template <int N, int R0, int R1, int R2, int T, class Pro>
__global__ void __launch_bounds__(T) col_fwd_kernel(Pro pro, Twiddles twd, Dst<cpx> out, int W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* smem = reinterpret_cast<cpx*>(smem_raw);
  const int b = blockIdx.y, c0 = blockIdx.x * kColTile, tid = threadIdx.x;
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
  const int b = blockIdx.y, c0 = blockIdx.x * kColTile, tid = threadIdx.x;
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

// fused inverse column pass -> kernel function -> forward column pass (correlation_flow.cc:212-215 / :222-225 between
// the IFFT and the FFT): the real kernel image lives only in shared memory.
template <int N, int I0, int I1, int I2, int F0, int F1, int F2, int T>
__global__ void __launch_bounds__(T) colcol_kernel(Src<cpx> in, Dst<cpx> out, Twiddles twi, Twiddles twf, KernelFn kfn, int W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* smem = reinterpret_cast<cpx*>(smem_raw);
  const int b = blockIdx.y, c0 = blockIdx.x * kColTile, tid = threadIdx.x;
  col_inv_phase0<N, I0, I1, I2, T>(tid, smem, in.at(b), W, c0);
  __syncthreads();
  {
    CarryRegs<I1, ColGeom<N, I0, I1, I2, T>::ROUNDS1> st;
    col_stage1_read<N, I0, I1, I2, T, true>(tid, smem, twi, st);
    __syncthreads();
    col_stage1_write<N, I0, I1, I2, T, true>(tid, smem, st);
  }
  __syncthreads();
  auto fn = kfn.bind(b);
  col_inv_phase2_inplace<N, I0, I1, I2, T>(tid, smem, twi, fn);
  __syncthreads();
  {
    CarryRegs<F0, ColGeom<N, F0, F1, F2, T>::ROUNDS0> st;
    col_fwd_phase0s_read<N, F0, F1, F2, T>(tid, smem, st);
    __syncthreads();
    col_fwd_phase0s_write<N, F0, F1, F2, T>(tid, smem, st);
  }
  __syncthreads();
  {
    CarryRegs<F1, ColGeom<N, F0, F1, F2, T>::ROUNDS1> st;
    col_stage1_read<N, F0, F1, F2, T, false>(tid, smem, twf, st);
    __syncthreads();
    col_stage1_write<N, F0, F1, F2, T, false>(tid, smem, st);
  }
  __syncthreads();
  col_fwd_phase2<N, F0, F1, F2, T>(tid, smem, twf, out.at(b), W, c0);
  DeviceSync sync;
  fn.finish(tid, sync);
}

template <int N, int I0, int I1, int I2, int F0, int F1, int F2, int T>
static int run_colcol(Twiddles twi, Twiddles twf, Src<cpx> in, Dst<cpx> out, KernelFn fn, int W, int B, cudaStream_t s) {
  auto k = colcol_kernel<N, I0, I1, I2, F0, F1, F2, T>;
  const size_t smem = ColGeom<N, F0, F1, F2, T>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  k<<<dim3(W / kColTile, B), T, smem, s>>>(in, out, twi, twf, fn, W);
  return (int)cudaGetLastError();
}

template <int N, int R0, int R1, int R2, int T, class Pro>
static int run_col_fwd(Twiddles tw, Pro pro, Dst<cpx> out, int W, int B, cudaStream_t s) {
  auto k = col_fwd_kernel<N, R0, R1, R2, T, Pro>;
  const size_t smem = ColGeom<N, R0, R1, R2, T>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  k<<<dim3(W / kColTile, B), T, smem, s>>>(pro, tw, out, W);
  return (int)cudaGetLastError();
}
template <int N, int R0, int R1, int R2, int T, class Epi>
static int run_col_inv(Twiddles tw, Src<cpx> in, Epi epi, int W, int B, cudaStream_t s) {
  auto k = col_inv_kernel<N, R0, R1, R2, T, Epi>;
  const size_t smem = ColGeom<N, R0, R1, R2, T>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  k<<<dim3(W / kColTile, B), T, smem, s>>>(in, tw, epi, W);
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
int launch_col_fwd_polar(int N, Twiddles tw, PolarArgs pa, Dst<cpx> out, int W, int B, cudaStream_t s) {
  ProPolar pro{pa.power2, pa.H, pa.W, pa.Cp, pa.cs, pa.rho, pa.table};
  FWD_DISPATCH(ProPolar)
}
int launch_col_fwd_rotate(int N, Twiddles tw, RotateArgs ra, Dst<cpx> out, int W, int B, cudaStream_t s) {
  if (ra.is_u8) {
    ProRotate<true> pro{ra.f32, ra.u8, ra.lut, ra.H, ra.W, ra.mats, ra.sel};
    FWD_DISPATCH(ProRotate<true>)
  } else {
    ProRotate<false> pro{ra.f32, ra.u8, ra.lut, ra.H, ra.W, ra.mats, ra.sel};
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
#define INV_CASE(n, f0, f1, f2, i0, i1, i2, t) case n: return run_col_inv<n, i0, i1, i2, t>(tw, in, epi, W, B, s);
int launch_col_inv_store(int N, Twiddles tw, Src<cpx> in, EpiStore epi, int W, int B, cudaStream_t s) { INV_DISPATCH }
int launch_col_inv_store_pairs(int N, Twiddles tw, Src<cpx> in, EpiStorePairs epi, int W, int B, cudaStream_t s) { INV_DISPATCH }
int launch_col_inv_peak(int N, Twiddles tw, Src<cpx> in, EpiPeak epi, int W, int B, cudaStream_t s) { INV_DISPATCH }
#undef INV_CASE

int launch_colcol(int N, Twiddles twi, Twiddles twf, Src<cpx> in, Dst<cpx> out, KernelFn fn, int W, int B, cudaStream_t s) {
  if (B <= 0) return 0;
  switch (N) {
#define X(n, f0, f1, f2, i0, i1, i2, t) case n: return run_colcol<n, i0, i1, i2, f0, f1, f2, t>(twi, twf, in, out, fn, W, B, s);
    NIS_COL_PLANS(X)
#undef X
    default: return -1;
  }
}

}  // namespace nis
