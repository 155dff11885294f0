// nis_row.cu -- row-pass kernels (complex transform along contiguous image columns), sm_100a.
// One CTA = L consecutive spectrum lines out of the flattened (batch x (R/2+1)) line space.
#include "nis_device.cuh"
#include "nis_internal.h"
#include "nis_sizes.h"

#include <algorithm>

#ifndef NIS_ROW_THREADS_PER_SM
#define NIS_ROW_THREADS_PER_SM 1024     // resident row-pass threads per SM the register budget is sized for (64 registers / thread)
#endif

namespace nis {

template <int N, int R1, int R2, int L, int T, bool INV, class Pro, class Epi>
__global__ void __launch_bounds__(T, (T <= 256 ? NIS_ROW_THREADS_PER_SM / T : 1)) row_kernel(Pro pro, Epi epi, Twiddles twd, int nrows, int total_lines) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* smem = reinterpret_cast<cpx*>(smem_raw);
  const int tid = threadIdx.x, line0 = blockIdx.x * L;
  const int nl = min(L, total_lines - line0);
  const LineMap m{line0, nrows, N};
  const auto bp = pro.bind(m);
  auto be = epi.bind(m);
  row_phase0<N, R1, R2, L, T, INV>(tid, smem, bp, nl);
  __syncthreads();
  CarryRegs<R1, RowGeom<N, R1, R2, L, T>::ROUNDS1> st;
  row_stage1_read<N, R1, R2, L, T, INV>(tid, smem, twd, nl, st);
  __syncthreads();
  row_stage1_write<N, R1, R2, L, T, INV>(tid, smem, nl, st);
  __syncthreads();
  row_phase2<N, R1, R2, L, T, INV>(tid, smem, twd, nl, be);
}

// fused forward row pass -> element-wise -> inverse row pass: K^xz, X = FFT(rotated image) and the filtered spectrum G
// exist only in registers / shared memory.  Two padded line buffers (the inverse stage 0 cannot run in place).
template <int N, int R1, int R2, int L, int T, class Mid>
__global__ void __launch_bounds__(T, (T <= 256 ? NIS_ROW_THREADS_PER_SM / T : 1)) rowrow_kernel(Src<cpx> in, Dst<cpx> out, Mid mid, Twiddles twd, int nrows, int total_lines) {
  typedef RowGeom<N, R1, R2, L, T> Gm;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* bufA = reinterpret_cast<cpx*>(smem_raw);
  cpx* bufB = bufA + L * Gm::PITCH;
  __shared__ float line_acc[L];
  const int tid = threadIdx.x, line0 = blockIdx.x * L;
  const int nl = min(L, total_lines - line0);
  const LineMap m{line0, nrows, N};
  if (tid < L) line_acc[tid] = 0.f;
  {
    const auto bp = ProSpec{in}.bind(m);
    row_phase0<N, R1, R2, L, T, false>(tid, bufA, bp, nl);
  }
  __syncthreads();
  {
    CarryRegs<R1, Gm::ROUNDS1> st;
    row_stage1_read<N, R1, R2, L, T, false>(tid, bufA, twd, nl, st);
    __syncthreads();
    row_stage1_write<N, R1, R2, L, T, false>(tid, bufA, nl, st);
  }
  __syncthreads();
  auto bm = mid.bind(m, line_acc);
  row_phase2_mid<N, R1, R2, L, T>(tid, bufA, twd, nl, bm);
  __syncthreads();
  {
    const SmemLinePro<Gm::PITCH> sp{bufA};
    row_phase0<N, R1, R2, L, T, true>(tid, bufB, sp, nl);
  }
  __syncthreads();
  {
    CarryRegs<R1, Gm::ROUNDS1> st;
    row_stage1_read<N, R1, R2, L, T, true>(tid, bufB, twd, nl, st);
    __syncthreads();
    row_stage1_write<N, R1, R2, L, T, true>(tid, bufB, nl, st);
  }
  __syncthreads();
  auto be = EpiSpecStore{out}.bind(m);
  row_phase2<N, R1, R2, L, T, true>(tid, bufB, twd, nl, be);
  if (tid < nl) bm.finish_line(tid);
}

template <int N, int R1, int R2, int L, int T, class Mid>
static int run_rowrow(Twiddles tw, Src<cpx> in, Dst<cpx> out, Mid mid, int nrows, int B, cudaStream_t s) {
  auto k = rowrow_kernel<N, R1, R2, L, T, Mid>;
  const size_t smem = 2 * RowGeom<N, R1, R2, L, T>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  const int total = nrows * B;
  k<<<(total + L - 1) / L, T, smem, s>>>(in, out, mid, tw, nrows, total);
  return (int)cudaGetLastError();
}

// =========================================================================================================
// Pipelined variants (sm_100a bulk async copy + mbarrier): persistent CTAs walk the line tiles of the batch; while tile k runs its
// butterflies, the L input lines of tile k+1 -- one contiguous block of global memory -- stream into the other half of a double
// staging buffer by cp.async.bulk (the TMA engine, no registers, no LSU instructions), completion counted on an mbarrier.  Stage 0
// then reads its operands from shared memory instead of opening the pass with 16 dependent-latency global loads per thread.
// Shared memory (staging + work buffers) bounds these kernels at 7 (5 for the fused forward->inverse kernel) CTAs per SM, so the
// register budget is 73 / 102 per thread and nothing spills.
// MEASURED AND REJECTED (profiles/ab_r02.md): 60.5k solves/s with these kernels against 68.6k without -- the fused forward->inverse
// kernel goes from 2.16 to 2.51 ms per 1000 frames: 8 ordinary CTAs per SM already hide the load latency, and the staging buffers
// cost two of them.  Off by default; -DNIS_ROW_PIPE=1 builds the variant for re-measurement.
// =========================================================================================================
#ifndef NIS_ROW_PIPE
#define NIS_ROW_PIPE 0
#endif
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
// prologue reading a natural-order DENSE line out of the staging buffer
template <int N> struct SmemDensePro {
  const cpx* base;
  struct Line { const cpx* p; __device__ __forceinline__ cpx load(int c) const { return p[c]; } };
  __device__ __forceinline__ Line line(int ln) const { return Line{base + ln * N}; }
};

template <int N, int R1, int R2, int L, int T, bool INV, class Epi>
__global__ void __launch_bounds__(T, (T <= 128 ? 7 : 1)) row_pipe_kernel(const cpx* __restrict__ in, Epi epi, Twiddles twd, int nrows, int total_lines) {
  typedef RowGeom<N, R1, R2, L, T> Gm;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cpx* work = reinterpret_cast<cpx*>(smem_raw);
  cpx* stage = work + L * Gm::PITCH;                       // [2][L * N], dense
  __shared__ __align__(8) unsigned long long mbar[2];
  const int tid = threadIdx.x, ntiles = (total_lines + L - 1) / L;
  const uint32_t bar0 = smem_u32(&mbar[0]);
  if (tid == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int tile, int buf) {
    const int nl = min(L, total_lines - tile * L);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    bulk_load(smem_u32(stage + buf * L * N), in + (size_t)tile * L * N, (uint32_t)(nl * N * sizeof(cpx)), bar0 + 8 * buf);
  };
  if (tid == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1, next = tile + gridDim.x;
    if (tid == 0 && next < ntiles) issue(next, buf ^ 1);     // that half was consumed before the barriers of the previous tile
    mbar_wait(bar0 + 8 * buf, (it >> 1) & 1);
    const int line0 = tile * L, nl = min(L, total_lines - line0);
    const LineMap m{line0, nrows, N};
    auto be = epi.bind(m);
    row_phase0<N, R1, R2, L, T, INV>(tid, work, SmemDensePro<N>{stage + buf * L * N}, nl);
    __syncthreads();
    CarryRegs<R1, Gm::ROUNDS1> st;
    row_stage1_read<N, R1, R2, L, T, INV>(tid, work, twd, nl, st);
    __syncthreads();
    row_stage1_write<N, R1, R2, L, T, INV>(tid, work, nl, st);
    __syncthreads();
    row_phase2<N, R1, R2, L, T, INV>(tid, work, twd, nl, be);
    __syncthreads();                                         // the work buffer is free for the next tile
  }
}

template <int N, int R1, int R2, int L, int T, class Mid>
__global__ void __launch_bounds__(T, (T <= 128 ? 5 : 1)) rowrow_pipe_kernel(const cpx* __restrict__ in, Dst<cpx> out, Mid mid, Twiddles twd, int nrows, int total_lines) {
  typedef RowGeom<N, R1, R2, L, T> Gm;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cpx* bufA = reinterpret_cast<cpx*>(smem_raw);
  cpx* bufB = bufA + L * Gm::PITCH;
  cpx* stage = bufB + L * Gm::PITCH;                       // [2][L * N], dense
  __shared__ float line_acc[L];
  __shared__ __align__(8) unsigned long long mbar[2];
  const int tid = threadIdx.x, ntiles = (total_lines + L - 1) / L;
  const uint32_t bar0 = smem_u32(&mbar[0]);
  if (tid == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int tile, int buf) {
    const int nl = min(L, total_lines - tile * L);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    bulk_load(smem_u32(stage + buf * L * N), in + (size_t)tile * L * N, (uint32_t)(nl * N * sizeof(cpx)), bar0 + 8 * buf);
  };
  if (tid == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1, next = tile + gridDim.x;
    if (tid == 0 && next < ntiles) issue(next, buf ^ 1);
    if (tid < L) line_acc[tid] = 0.f;
    mbar_wait(bar0 + 8 * buf, (it >> 1) & 1);
    const int line0 = tile * L, nl = min(L, total_lines - line0);
    const LineMap m{line0, nrows, N};
    row_phase0<N, R1, R2, L, T, false>(tid, bufA, SmemDensePro<N>{stage + buf * L * N}, nl);
    __syncthreads();
    {
      CarryRegs<R1, Gm::ROUNDS1> st;
      row_stage1_read<N, R1, R2, L, T, false>(tid, bufA, twd, nl, st);
      __syncthreads();
      row_stage1_write<N, R1, R2, L, T, false>(tid, bufA, nl, st);
    }
    __syncthreads();
    auto bm = mid.bind(m, line_acc);
    row_phase2_mid<N, R1, R2, L, T>(tid, bufA, twd, nl, bm);
    __syncthreads();
    {
      const SmemLinePro<Gm::PITCH> sp{bufA};
      row_phase0<N, R1, R2, L, T, true>(tid, bufB, sp, nl);
    }
    __syncthreads();
    {
      CarryRegs<R1, Gm::ROUNDS1> st;
      row_stage1_read<N, R1, R2, L, T, true>(tid, bufB, twd, nl, st);
      __syncthreads();
      row_stage1_write<N, R1, R2, L, T, true>(tid, bufB, nl, st);
    }
    __syncthreads();
    auto be = EpiSpecStore{out}.bind(m);
    row_phase2<N, R1, R2, L, T, true>(tid, bufB, twd, nl, be);
    if (tid < nl) bm.finish_line(tid);
    __syncthreads();
  }
}

// persistent grid: as many CTAs as are resident at once (occupancy query per instantiation), capped by the tile count
template <class K> static int persistent_grid(K kernel, int threads, size_t smem, int ntiles) {
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  return std::max(1, std::min(ntiles, sms * per_sm));
}
// the pipelined kernels need one contiguous run of lines: a slab whose stride equals the spectrum size (or a single element)
static bool contiguous_lines(const Src<cpx>& x, int nrows, int N, int B) {
  return NIS_ROW_PIPE && x.base && !x.ptrs && !x.idx && x.shift == 0 && (B == 1 || x.stride == (long long)nrows * N);
}

template <int N, int R1, int R2, int L, int T, bool INV, class Epi>
static int run_row_pipe(Twiddles tw, const cpx* in, Epi epi, int nrows, int B, cudaStream_t s) {
  auto k = row_pipe_kernel<N, R1, R2, L, T, INV, Epi>;
  const size_t smem = RowGeom<N, R1, R2, L, T>::kSmemBytes + 2 * (size_t)L * N * sizeof(cpx);
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  const int total = nrows * B, ntiles = (total + L - 1) / L;
  static int cap = persistent_grid(k, T, smem, 1 << 30);
  k<<<std::min(ntiles, cap), T, smem, s>>>(in, epi, tw, nrows, total);
  return (int)cudaGetLastError();
}
template <int N, int R1, int R2, int L, int T, class Mid>
static int run_rowrow_pipe(Twiddles tw, const cpx* in, Dst<cpx> out, Mid mid, int nrows, int B, cudaStream_t s) {
  auto k = rowrow_pipe_kernel<N, R1, R2, L, T, Mid>;
  const size_t smem = 2 * RowGeom<N, R1, R2, L, T>::kSmemBytes + 2 * (size_t)L * N * sizeof(cpx);
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  const int total = nrows * B, ntiles = (total + L - 1) / L;
  static int cap = persistent_grid(k, T, smem, 1 << 30);
  k<<<std::min(ntiles, cap), T, smem, s>>>(in, out, mid, tw, nrows, total);
  return (int)cudaGetLastError();
}

template <int N, int R1, int R2, int L, int T, bool INV, class Pro, class Epi>
static int run_row(Twiddles tw, Pro pro, Epi epi, int nrows, int B, cudaStream_t s) {
  auto k = row_kernel<N, R1, R2, L, T, INV, Pro, Epi>;
  const size_t smem = RowGeom<N, R1, R2, L, T>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  const int total = nrows * B;
  k<<<(total + L - 1) / L, T, smem, s>>>(pro, epi, tw, nrows, total);
  return (int)cudaGetLastError();
}

bool row_size_supported(int N) {
#define X(n, r1, r2, l, t, lr) if (N == n) return true;
  NIS_ROW_PLANS(X)
#undef X
  return false;
}
void plan_radices_row(int N, int r[3]) {
#define X(n, r1, r2, l, t, lr) if (N == n) { r[0] = 16; r[1] = r1; r[2] = r2; return; }
  NIS_ROW_PLANS(X)
#undef X
  r[0] = r[1] = r[2] = 0;
}

#define ROW_DISPATCH(INV)                \
  if (B <= 0) return 0;                  \
  switch (N) {                           \
    NIS_ROW_PLANS(ROW_CASE_##INV)        \
    default: return -1;                  \
  }
#define ROW_CASE_false(n, r1, r2, l, t, lr) case n: return run_row<n, r1, r2, l, t, false>(tw, pro, epi, nrows, B, s);
#define ROW_CASE_pipe(n, r1, r2, l, t, lr) case n: return run_row_pipe<n, r1, r2, lr, t, false>(tw, pro.x.base, epi, nrows, B, s);
#define ROW_DISPATCH_PIPE                                        \
  if (B > 0 && contiguous_lines(pro.x, nrows, N, B)) {          \
    switch (N) {                                                 \
      NIS_ROW_PLANS(ROW_CASE_pipe)                               \
      default: return -1;                                        \
    }                                                            \
  }
#define ROW_CASE_true(n, r1, r2, l, t, lr) case n: return run_row<n, r1, r2, l, t, true>(tw, pro, epi, nrows, B, s);
int launch_row_fwd(int N, Twiddles tw, ProSpec pro, EpiSpecStore epi, int nrows, int B, cudaStream_t s) { ROW_DISPATCH_PIPE ROW_DISPATCH(false) }
int launch_row_fwd_h(int N, Twiddles tw, ProSpec pro, EpiHStore epi, int nrows, int B, cudaStream_t s) { ROW_DISPATCH_PIPE ROW_DISPATCH(false) }
int launch_row_inv_mulconj(int N, Twiddles tw, ProMulConj pro, EpiSpecStore epi, int nrows, int B, cudaStream_t s) { ROW_DISPATCH(true) }

#define RR_DISPATCH                       \
  if (B <= 0) return 0;                   \
  switch (N) {                            \
    NIS_ROW_PLANS(RR_CASE)                \
    default: return -1;                   \
  }
#define RR_CASE(n, r1, r2, l, t, lr) \
  case n: return contiguous_lines(in, nrows, N, B) ? run_rowrow_pipe<n, r1, r2, lr, t>(tw, in.base, out, mid, nrows, B, s) \
                                                  : run_rowrow<n, r1, r2, lr, t>(tw, in, out, mid, nrows, B, s);
int launch_rowrow_mulconj(int N, Twiddles tw, Src<cpx> in, Dst<cpx> out, MidMulConjZ mid, int nrows, int B, cudaStream_t s) { RR_DISPATCH }
int launch_rowrow_filter(int N, Twiddles tw, Src<cpx> in, Dst<cpx> out, MidFilterH mid, int nrows, int B, cudaStream_t s) { RR_DISPATCH }
int launch_rowrow_storeabs(int N, Twiddles tw, Src<cpx> in, Dst<cpx> out, MidStoreAbs mid, int nrows, int B, cudaStream_t s) { RR_DISPATCH }

}  // namespace nis
