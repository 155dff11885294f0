// nis_row.cu -- row-pass kernels (complex transform along contiguous image columns), sm_100a.
// One CTA = L consecutive spectrum lines out of the flattened (batch x (R/2+1)) line space.
#include "nis_device.cuh"
#include "nis_internal.h"
#include "nis_sizes.h"

#ifndef NIS_ROW_THREADS_PER_SM
#define NIS_ROW_THREADS_PER_SM 1024     // resident row-pass threads per SM the register budget is sized for (64 registers / thread)
#endif

namespace nis {

// resident CTAs per SM the register budget is sized for: 64 registers per thread for the three-stage plans; the two-stage plans
// (R2 == 1) hold a radix-32 butterfly per thread and get 102
template <int R0, int R1, int R2, int T> constexpr int row_min_blocks() {
  return T > 256 ? 1 : ((R2 == 1 && (R0 >= 32 || R1 >= 32)) ? 640 / T : NIS_ROW_THREADS_PER_SM / T);
}

template <int N, int R0, int R1, int R2, int L, int T, bool INV, class Pro, class Epi>
__global__ void __launch_bounds__(T, (row_min_blocks<R0, R1, R2, T>())) row_kernel(Pro pro, Epi epi, Twiddles twd, int nrows, int total_lines) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* smem = reinterpret_cast<cpx*>(smem_raw);
  const int tid = threadIdx.x, line0 = blockIdx.x * L;
  const int nl = min(L, total_lines - line0);
  const LineMap m{line0, nrows, N};
  const auto bp = pro.bind(m);
  auto be = epi.bind(m);
  row_phase0<N, R0, R1, R2, L, T, INV>(tid, smem, bp, nl);
  __syncthreads();
  if constexpr (R2 == 1) {
    row_stage1_out<N, R0, R1, L, T, INV>(tid, smem, twd, nl, be);
  } else {
    CarryRegs<R1, RowGeom<N, R0, R1, R2, L, T>::ROUNDS1> st;
    row_stage1_read<N, R0, R1, R2, L, T, INV>(tid, smem, twd, nl, st);
    __syncthreads();
    row_stage1_write<N, R0, R1, R2, L, T, INV>(tid, smem, nl, st);
    __syncthreads();
    row_phase2<N, R0, R1, R2, L, T, INV>(tid, smem, twd, nl, be);
  }
}

// fused forward row pass -> element-wise -> inverse row pass: K^xz, X = FFT(rotated image) and the filtered spectrum G
// exist only in registers / shared memory.  Two padded line buffers (the inverse stage 0 cannot run in place).
template <int N, int R0, int R1, int R2, int L, int T, class Mid>
__global__ void __launch_bounds__(T, (row_min_blocks<R0, R1, R2, T>())) rowrow_kernel(Src<cpx> in, Dst<cpx> out, Mid mid, Twiddles twd, int nrows, int total_lines) {
  typedef RowGeom<N, R0, R1, R2, L, T> Gm;
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
    row_phase0<N, R0, R1, R2, L, T, false>(tid, bufA, bp, nl);
  }
  __syncthreads();
  auto bm = mid.bind(m, line_acc);
  if constexpr (R2 == 1) {
    row_stage1_mid<N, R0, R1, L, T>(tid, bufA, twd, nl, bm);
  } else {
    {
      CarryRegs<R1, Gm::ROUNDS1> st;
      row_stage1_read<N, R0, R1, R2, L, T, false>(tid, bufA, twd, nl, st);
      __syncthreads();
      row_stage1_write<N, R0, R1, R2, L, T, false>(tid, bufA, nl, st);
    }
    __syncthreads();
    row_phase2_mid<N, R0, R1, R2, L, T>(tid, bufA, twd, nl, bm);
  }
  __syncthreads();
  {
    const SmemLinePro<Gm::PITCH, R0> sp{bufA};
    row_phase0<N, R0, R1, R2, L, T, true>(tid, bufB, sp, nl);
  }
  __syncthreads();
  auto be = EpiSpecStore{out}.bind(m);
  if constexpr (R2 == 1) {
    row_stage1_out<N, R0, R1, L, T, true>(tid, bufB, twd, nl, be);
  } else {
    {
      CarryRegs<R1, Gm::ROUNDS1> st;
      row_stage1_read<N, R0, R1, R2, L, T, true>(tid, bufB, twd, nl, st);
      __syncthreads();
      row_stage1_write<N, R0, R1, R2, L, T, true>(tid, bufB, nl, st);
    }
    __syncthreads();
    row_phase2<N, R0, R1, R2, L, T, true>(tid, bufB, twd, nl, be);
  }
  if (tid < nl) bm.finish_line(tid);
}

template <int N, int R0, int R1, int R2, int L, int T, class Mid>
static int run_rowrow(Twiddles tw, Src<cpx> in, Dst<cpx> out, Mid mid, int nrows, int B, cudaStream_t s) {
  auto k = rowrow_kernel<N, R0, R1, R2, L, T, Mid>;
  const size_t smem = 2 * RowGeom<N, R0, R1, R2, L, T>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  const int total = nrows * B;
  k<<<(total + L - 1) / L, T, smem, s>>>(in, out, mid, tw, nrows, total);
  return (int)cudaGetLastError();
}

// A bulk-async variant of these kernels (persistent CTAs, cp.async.bulk + mbarrier double staging buffer for the next tile's lines)
// was built and measured in round 2: 60.5k solves/s against 68.6k with the kernels above (profiles/ab_r02.md) -- 8 resident CTAs per
// SM already hide the load latency and the staging buffers cost two of them.  It lives in the history (commit 53c9120), not here.

template <int N, int R0, int R1, int R2, int L, int T, bool INV, class Pro, class Epi>
static int run_row(Twiddles tw, Pro pro, Epi epi, int nrows, int B, cudaStream_t s) {
  auto k = row_kernel<N, R0, R1, R2, L, T, INV, Pro, Epi>;
  const size_t smem = RowGeom<N, R0, R1, R2, L, T>::kSmemBytes;
  static int attr = set_smem(k, smem);
  if (attr) return attr;
  const int total = nrows * B;
  k<<<(total + L - 1) / L, T, smem, s>>>(pro, epi, tw, nrows, total);
  return (int)cudaGetLastError();
}

bool row_size_supported(int N) {
#define X(n, r0, r1, r2, l, t, lr) if (N == n) return true;
  NIS_ROW_PLANS(X)
#undef X
  return false;
}
void plan_radices_row(int N, int r[3]) {
#define X(n, r0, r1, r2, l, t, lr) if (N == n) { r[0] = r0; r[1] = r1; r[2] = r2; return; }
  NIS_ROW_PLANS(X)
#undef X
  r[0] = r[1] = r[2] = 0;
}
bool plan_radices_row_b(int N, int r[3]) {
#define X(n, r0, r1, r2, l, t, lr) if (N == n) { r[0] = r0; r[1] = r1; r[2] = r2; return true; }
  NIS_ROW_PLANS_B(X)
#undef X
  return false;
}

// plan B first where the launcher's family prefers it and the size has one, plan A otherwise
#define ROW_DISPATCH(INV, USE_B)                           \
  if (B <= 0) return 0;                                    \
  if (USE_B) switch (N) {                                  \
    NIS_ROW_PLANS_B(ROWB_CASE_##INV)                       \
    default: break;                                        \
  }                                                        \
  switch (N) {                                             \
    NIS_ROW_PLANS(ROW_CASE_##INV)                          \
    default: return -1;                                    \
  }
#define ROW_CASE_false(n, r0, r1, r2, l, t, lr) case n: return run_row<n, r0, r1, r2, l, t, false>(tw.a, pro, epi, nrows, B, s);
#define ROW_CASE_true(n, r0, r1, r2, l, t, lr) case n: return run_row<n, r0, r1, r2, l, t, true>(tw.a, pro, epi, nrows, B, s);
#define ROWB_CASE_false(n, r0, r1, r2, l, t, lr) case n: return run_row<n, r0, r1, r2, l, t, false>(tw.b, pro, epi, nrows, B, s);
#define ROWB_CASE_true(n, r0, r1, r2, l, t, lr) case n: return run_row<n, r0, r1, r2, l, t, true>(tw.b, pro, epi, nrows, B, s);
// match_fused: take the plan of the fused fwd->mid->inv kernels, so that the scan's cached-rotation path (row_fwd into the cache, then
// row_inv_mulconj) rounds exactly like its per-candidate path (rowrow_mulconj)
int launch_row_fwd(int N, RowTwiddles tw, ProSpec pro, EpiSpecStore epi, int nrows, int B, cudaStream_t s, bool match_fused) {
  ROW_DISPATCH(false, match_fused ? NIS_ROWB_RR : NIS_ROWB_FWD)
}
int launch_row_fwd_h(int N, RowTwiddles tw, ProSpec pro, EpiHStore epi, int nrows, int B, cudaStream_t s) { ROW_DISPATCH(false, NIS_ROWB_FWDH) }
int launch_row_inv_mulconj(int N, RowTwiddles tw, ProMulConj pro, EpiSpecStore epi, int nrows, int B, cudaStream_t s, bool match_fused) {
  const bool auto_form = pro.x.base == pro.z.base && pro.x.ptrs == pro.z.ptrs && pro.x.idx == pro.z.idx && pro.x.offset == pro.z.offset;
  ROW_DISPATCH(true, match_fused ? NIS_ROWB_RR : (auto_form ? NIS_ROWB_INVMC_AUTO : NIS_ROWB_INVMC))
}

#define RR_DISPATCH                                        \
  if (B <= 0) return 0;                                    \
  if (NIS_ROWB_RR) switch (N) {                            \
    NIS_ROW_PLANS_B(RRB_CASE)                              \
    default: break;                                        \
  }                                                        \
  switch (N) {                                             \
    NIS_ROW_PLANS(RR_CASE)                                 \
    default: return -1;                                    \
  }
#define RR_CASE(n, r0, r1, r2, l, t, lr) case n: return run_rowrow<n, r0, r1, r2, lr, t>(tw.a, in, out, mid, nrows, B, s);
#define RRB_CASE(n, r0, r1, r2, l, t, lr) case n: return run_rowrow<n, r0, r1, r2, lr, t>(tw.b, in, out, mid, nrows, B, s);
int launch_rowrow_mulconj(int N, RowTwiddles tw, Src<cpx> in, Dst<cpx> out, MidMulConjZ mid, int nrows, int B, cudaStream_t s) { RR_DISPATCH }
int launch_rowrow_filter(int N, RowTwiddles tw, Src<cpx> in, Dst<cpx> out, MidFilterH mid, int nrows, int B, cudaStream_t s) { RR_DISPATCH }
int launch_rowrow_storeabs(int N, RowTwiddles tw, Src<cpx> in, Dst<cpx> out, MidStoreAbs mid, int nrows, int B, cudaStream_t s) { RR_DISPATCH }
int launch_rowrow_storesq(int N, RowTwiddles tw, Src<cpx> in, Dst<cpx> out, MidStoreSq mid, int nrows, int B, cudaStream_t s) { RR_DISPATCH }

}  // namespace nis
