// Host emulation of the CUDA FFT passes: each CTA is replayed thread by thread, phase by phase (a phase is the
// code between two __syncthreads()).  Lets tests/test_emulation.py validate the kernels' index logic against
// numpy without a GPU.  Build: g++ -O2 -shared -fPIC -I ni_slam_b200/csrc tests/cpp/emu_fft.cc
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nis_ops.cuh"
#include "nis_sizes.h"

using namespace nis;

struct HostSync {
  void block_max_to(unsigned int* p, float mx, int) {
    union { float f; unsigned int u; } c; c.f = mx;
    if (c.u > *p) *p = c.u;
  }
  void block_peak_to(PeakStats* st, unsigned long long key, double s, double q, int) {
    if (key > st->key) st->key = key;
    st->sum += s; st->sumsq += q;
  }
};

static void make_tw(int R0, int R1, int R2, std::vector<cpx>& t1, std::vector<cpx>& t2) {
  const int N = R0 * R1 * R2;
  t1.assign((size_t)(R1 > 1 ? (R1 - 1) : 1) * R0, cpx{1, 0});
  t2.assign((size_t)(R2 > 1 ? (R2 - 1) : 1) * R0 * R1, cpx{1, 0});
  for (int r = 1; r < R1; ++r)
    for (int k = 0; k < R0; ++k) {
      double a = -2.0 * M_PI * r * k / (double)(R0 * R1);
      t1[(size_t)(r - 1) * R0 + k] = cpx{(float)cos(a), (float)sin(a)};
    }
  for (int r = 1; r < R2; ++r)
    for (int k = 0; k < R0 * R1; ++k) {
      double a = -2.0 * M_PI * r * k / (double)N;
      t2[(size_t)(r - 1) * R0 * R1 + k] = cpx{(float)cos(a), (float)sin(a)};
    }
}

// column-pass twiddles of the in-place engine (nis_fft.cuh): tw1[(b-1)*C + d2], tw2[(a-1)*BC + j]
static void make_tw_col(int A, int B, int C, std::vector<cpx>& t1, std::vector<cpx>& t2) {
  const int N = A * B * C;
  t1.assign((size_t)(B > 1 ? (B - 1) : 1) * C, cpx{1, 0});
  t2.assign((size_t)(A > 1 ? (A - 1) : 1) * B * C, cpx{1, 0});
  for (int r = 1; r < B; ++r)
    for (int k = 0; k < C; ++k) {
      double a = -2.0 * M_PI * r * k / (double)(B * C);
      t1[(size_t)(r - 1) * C + k] = cpx{(float)cos(a), (float)sin(a)};
    }
  for (int r = 1; r < A; ++r)
    for (int k = 0; k < B * C; ++k) {
      double a = -2.0 * M_PI * r * k / (double)N;
      t2[(size_t)(r - 1) * B * C + k] = cpx{(float)cos(a), (float)sin(a)};
    }
}

// ---- column forward: real [B][N][W] -> half-transformed spectrum [B][N/2+1][W]
template <int N, int A, int B, int C, int T, int LN = kColLanes>
static void emu_col_fwd(const float* x, int Bn, int W, cpx* out) {
  typedef ColGeom<N, A, B, C, T, LN> Gm;
  std::vector<cpx> t1, t2; make_tw_col(A, B, C, t1, t2);
  Twiddles twd{t1.data(), t2.data()};
  std::vector<cpx> smem((size_t)Gm::SLOTS * LN);
  ProRealF32 pro{Src<float>{x, (long long)N * W, nullptr, 0, nullptr, 0}, W};
  for (int b = 0; b < Bn; ++b)
    for (int c0 = 0; c0 < W; c0 += 2 * LN) {
      auto bp = pro.bind(b, c0);
      for (int t = 0; t < T; ++t) col_fwd_stage_a<N, A, B, C, T, decltype(bp), LN>(t, smem.data(), twd, bp);
      if constexpr (B > 1) for (int t = 0; t < T; ++t) col_stage_b<N, A, B, C, T, false, false, LN>(t, smem.data(), twd);
      for (int t = 0; t < T; ++t) col_fwd_stage_c<N, A, B, C, T, LN>(t, smem.data(), out + (size_t)b * (N / 2 + 1) * W, W, c0);
    }
}

// ---- column inverse: spectrum [B][N/2+1][W] -> real [B][N][W] / n
template <int N, int A, int B, int C, int T, class Epi>
static void emu_col_inv(const cpx* in, int Bn, int W, Epi& epi) {
  typedef ColGeom<N, A, B, C, T> Gm;
  std::vector<cpx> t1, t2; make_tw_col(A, B, C, t1, t2);
  Twiddles twd{t1.data(), t2.data()};
  std::vector<cpx> smem((size_t)Gm::SLOTS * kColLanes);
  HostSync sync;
  for (int b = 0; b < Bn; ++b)
    for (int c0 = 0; c0 < W; c0 += kColTile) {
      std::vector<typename Epi::Bound> eb(T, epi.bind(b, c0));
      for (int t = 0; t < T; ++t) col_inv_stage_a<N, A, B, C, T>(t, smem.data(), twd, in + (size_t)b * (N / 2 + 1) * W, W, c0);
      if constexpr (B > 1) for (int t = 0; t < T; ++t) col_stage_b<N, A, B, C, T, true, false>(t, smem.data(), twd);
      for (int t = 0; t < T; ++t) col_inv_stage_c<N, A, B, C, T>(t, smem.data(), eb[t]);
      for (int t = 0; t < T; ++t) eb[t].finish(t, sync);
    }
}

// ---- row pass: [nlines][N] complex, forward or inverse, plain load/store
template <int N, int R0, int R1, int R2, int L, int T, bool INV>
static void emu_row(const cpx* in, int total_lines, cpx* out) {
  typedef RowGeom<N, R0, R1, R2, L, T> Gm;
  std::vector<cpx> t1, t2; make_tw(R0, R1, R2, t1, t2);
  Twiddles twd{t1.data(), t2.data()};
  std::vector<cpx> smem((size_t)L * Gm::PITCH);
  std::vector<CarryRegs<R1, Gm::ROUNDS1>> st(T);
  ProSpec pro{Src<cpx>{in, (long long)total_lines * N, nullptr, 0, nullptr, 0}};
  EpiSpecStore epi{Dst<cpx>{out, (long long)total_lines * N}};
  for (int line0 = 0; line0 < total_lines; line0 += L) {
    const int nl = total_lines - line0 < L ? total_lines - line0 : L;
    LineMap m{line0, total_lines, N};
    auto bp = pro.bind(m);
    auto be = epi.bind(m);
    for (int t = 0; t < T; ++t) row_phase0<N, R0, R1, R2, L, T, INV>(t, smem.data(), bp, nl);
    if constexpr (R2 == 1) {
      for (int t = 0; t < T; ++t) row_stage1_out<N, R0, R1, L, T, INV>(t, smem.data(), twd, nl, be);
      continue;
    }
    for (int t = 0; t < T; ++t) row_stage1_read<N, R0, R1, R2, L, T, INV>(t, smem.data(), twd, nl, st[t]);
    for (int t = 0; t < T; ++t) row_stage1_write<N, R0, R1, R2, L, T, INV>(t, smem.data(), nl, st[t]);
    for (int t = 0; t < T; ++t) row_phase2<N, R0, R1, R2, L, T, INV>(t, smem.data(), twd, nl, be);
  }
}

// ---- fused: inverse column pass -> kernel function -> forward column pass (colcol kernel)
template <int N, int A, int B, int C, int T>
static void emu_colcol(const cpx* in, int Bn, int W, cpx* out, KernelFn kfn) {
  typedef ColGeom<N, A, B, C, T> Gm;
  std::vector<cpx> t1, t2; make_tw_col(A, B, C, t1, t2);
  Twiddles twd{t1.data(), t2.data()};
  std::vector<cpx> smem((size_t)Gm::SLOTS * kColLanes);
  HostSync sync;
  for (int b = 0; b < Bn; ++b)
    for (int c0 = 0; c0 < W; c0 += kColTile) {
      const cpx* src = in + (size_t)b * (N / 2 + 1) * W;
      cpx* dst = out + (size_t)b * (N / 2 + 1) * W;
      std::vector<KernelFn::Bound> fn(T, kfn.bind(b));
      for (int t = 0; t < T; ++t) col_inv_stage_a<N, A, B, C, T>(t, smem.data(), twd, src, W, c0);
      if constexpr (B > 1) for (int t = 0; t < T; ++t) col_stage_b<N, A, B, C, T, true, false>(t, smem.data(), twd);
      for (int t = 0; t < T; ++t) col_inv_fn_fwd_stage_c<N, A, B, C, T>(t, smem.data(), fn[t]);
      if constexpr (B > 1) for (int t = 0; t < T; ++t) col_stage_b<N, A, B, C, T, false, true>(t, smem.data(), twd);
      for (int t = 0; t < T; ++t) col_fwd_dit_stage_a<N, A, B, C, T>(t, smem.data(), twd, dst, W, c0);
      for (int t = 0; t < T; ++t) fn[t].finish(t, sync);
    }
}

// ---- fused: forward row pass -> element-wise -> inverse row pass (rowrow kernel)
template <int N, int R0, int R1, int R2, int L, int T, class Mid>
static void emu_rowrow(const cpx* in, int nrows, int B, cpx* out, Mid mid) {
  typedef RowGeom<N, R0, R1, R2, L, T> Gm;
  std::vector<cpx> t1, t2; make_tw(R0, R1, R2, t1, t2);
  Twiddles twd{t1.data(), t2.data()};
  const int total = nrows * B;
  std::vector<cpx> A((size_t)L * Gm::PITCH), Bb((size_t)L * Gm::PITCH);
  std::vector<CarryRegs<R1, Gm::ROUNDS1>> st(T);
  ProSpec pro{Src<cpx>{in, (long long)nrows * N, nullptr, 0, nullptr, 0}};
  EpiSpecStore epi{Dst<cpx>{out, (long long)nrows * N}};
  for (int line0 = 0; line0 < total; line0 += L) {
    const int nl = total - line0 < L ? total - line0 : L;
    LineMap m{line0, nrows, N};
    float acc[L] = {0};
    auto bp = pro.bind(m);
    auto be = epi.bind(m);
    auto bm = mid.bind(m, acc);
    for (int t = 0; t < T; ++t) row_phase0<N, R0, R1, R2, L, T, false>(t, A.data(), bp, nl);
    if constexpr (R2 == 1) {
      for (int t = 0; t < T; ++t) row_stage1_mid<N, R0, R1, L, T>(t, A.data(), twd, nl, bm);
    } else {
      for (int t = 0; t < T; ++t) row_stage1_read<N, R0, R1, R2, L, T, false>(t, A.data(), twd, nl, st[t]);
      for (int t = 0; t < T; ++t) row_stage1_write<N, R0, R1, R2, L, T, false>(t, A.data(), nl, st[t]);
      for (int t = 0; t < T; ++t) row_phase2_mid<N, R0, R1, R2, L, T>(t, A.data(), twd, nl, bm);
    }
    SmemLinePro<Gm::PITCH, R0> sp{A.data()};
    for (int t = 0; t < T; ++t) row_phase0<N, R0, R1, R2, L, T, true>(t, Bb.data(), sp, nl);
    if constexpr (R2 == 1) {
      for (int t = 0; t < T; ++t) row_stage1_out<N, R0, R1, L, T, true>(t, Bb.data(), twd, nl, be);
    } else {
      for (int t = 0; t < T; ++t) row_stage1_read<N, R0, R1, R2, L, T, true>(t, Bb.data(), twd, nl, st[t]);
      for (int t = 0; t < T; ++t) row_stage1_write<N, R0, R1, R2, L, T, true>(t, Bb.data(), nl, st[t]);
      for (int t = 0; t < T; ++t) row_phase2<N, R0, R1, R2, L, T, true>(t, Bb.data(), twd, nl, be);
    }
    for (int t = 0; t < nl; ++t) bm.finish_line(t);
  }
}

extern "C" {

// in/out: [B][N/2+1][W] complex; maxbuf[B]; polynomial kernel (x/(N*W) + offset)^power
int emu_colcol_poly(const float* in, int B, int N, int W, float* out, unsigned int* maxbuf, float offset, int power) {
  memset(maxbuf, 0, sizeof(unsigned int) * B);
  KernelFn kfn{(float)((long long)N * W), 0, offset, power, 0.f, nullptr, nullptr, 0, maxbuf, nullptr};
#define X(n, a, b, c, t) \
  if (N == n) { emu_colcol<n, a, b, c, t>((const cpx*)in, B, W, (cpx*)out, kfn); return 0; }
  NIS_COL_PLANS(X)
#undef X
  return -1;
}

// in/out: [B][nrows][N]; z: [B][nrows][N]; out = IFFT_rows(FFT_rows(in) * conj(z)) (unnormalised); xx_sum[B] += sum |FFT_rows(in)|^2
int emu_rowrow_mulconj(const float* in, const float* z, int B, int nrows, int N, float* out, double* xx_sum) {
  MidMulConjZ mid{Src<cpx>{(const cpx*)z, (long long)nrows * N, nullptr, 0, nullptr, 0}, xx_sum};
#define X(n, r0, r1, r2, l, t, lr) \
  if (N == n) { emu_rowrow<n, r0, r1, r2, lr, t>((const cpx*)in, nrows, B, (cpx*)out, mid); return 0; }
  NIS_ROW_PLANS(X)
#undef X
  return -1;
}

// out = IFFT_rows(H * FFT_rows(in) / max[b])
int emu_rowrow_filter(const float* in, const float* h, const unsigned int* maxbuf, int B, int nrows, int N, float* out) {
  MidFilterH mid{Src<cpx>{(const cpx*)h, (long long)nrows * N, nullptr, 0, nullptr, 0}, maxbuf};
#define X(n, r0, r1, r2, l, t, lr) \
  if (N == n) { emu_rowrow<n, r0, r1, r2, lr, t>((const cpx*)in, nrows, B, (cpx*)out, mid); return 0; }
  NIS_ROW_PLANS(X)
#undef X
  return -1;
}

// returns 0 ok, -1 unsupported size
int emu_col_fwd_f32(const float* x, int B, int N, int W, float* out) {
#define X(n, a, b, c, t) \
  if (N == n) { emu_col_fwd<n, a, b, c, t>(x, B, W, (cpx*)out); return 0; }
  NIS_COL_PLANS(X)
#undef X
  return -1;
}

// the 16-lane forward pass (32 real columns per CTA) the rotation prologue runs with
int emu_col_fwd_f32_l16(const float* x, int B, int N, int W, float* out) {
#define X(n, a, b, c, t) \
  if (N == n) { emu_col_fwd<n, a, b, c, t, 16>(x, B, W, (cpx*)out); return 0; }
  NIS_COL_PLANS(X)
#undef X
  return -1;
}

int emu_col_inv_store(const float* in, int B, int N, int W, float* out) {
  EpiStore epi{Dst<float>{out, (long long)N * W}, W, (float)((long long)N * W)};
#define X(n, a, b, c, t) \
  if (N == n) { emu_col_inv<n, a, b, c, t>((const cpx*)in, B, W, epi); return 0; }
  NIS_COL_PLANS(X)
#undef X
  return -1;
}

// c2r column pass with the peak epilogue; stats: B x {key(u64), sum, sumsq}
int emu_col_inv_peak(const float* in, int B, int N, int W, void* stats, float* g_out) {
  memset(stats, 0, sizeof(PeakStats) * B);
  EpiPeak epi{(PeakStats*)stats, N, (float)((long long)N * W), g_out, (long long)N * W, W};
#define X(n, a, b, c, t) \
  if (N == n) { emu_col_inv<n, a, b, c, t>((const cpx*)in, B, W, epi); return 0; }
  NIS_COL_PLANS(X)
#undef X
  return -1;
}

int emu_row(const float* in, int total_lines, int N, int inverse, float* out) {
#define X(n, r0, r1, r2, l, t, lr)                                                      \
  if (N == n) {                                                                         \
    if (inverse) emu_row<n, r0, r1, r2, l, t, true>((const cpx*)in, total_lines, (cpx*)out); \
    else emu_row<n, r0, r1, r2, l, t, false>((const cpx*)in, total_lines, (cpx*)out);        \
    return 0;                                                                           \
  }
  NIS_ROW_PLANS(X)
#undef X
  return -1;
}

// f = FFT_rows(in) stored, out = IFFT_rows(|f|^2) (unnormalised); plan_b selects the two-stage plans
int emu_rowrow_storesq(const float* in, int B, int nrows, int N, float* f, float* out, int plan_b) {
  MidStoreSq mid{Dst<cpx>{(cpx*)f, (long long)nrows * N}};
#define X(n, r0, r1, r2, l, t, lr) \
  if (N == n) { emu_rowrow<n, r0, r1, r2, lr, t>((const cpx*)in, nrows, B, (cpx*)out, mid); return 0; }
  if (plan_b) { NIS_ROW_PLANS_B(X) }
  NIS_ROW_PLANS(X)
#undef X
  return -1;
}

// the two-stage plan B of the production row lengths (nis_sizes.h NIS_ROW_PLANS_B)
int emu_row_b(const float* in, int total_lines, int N, int inverse, float* out) {
#define X(n, r0, r1, r2, l, t, lr)                                                      \
  if (N == n) {                                                                         \
    if (inverse) emu_row<n, r0, r1, r2, l, t, true>((const cpx*)in, total_lines, (cpx*)out); \
    else emu_row<n, r0, r1, r2, l, t, false>((const cpx*)in, total_lines, (cpx*)out);        \
    return 0;                                                                           \
  }
  NIS_ROW_PLANS_B(X)
#undef X
  return -1;
}
int emu_rowrow_mulconj_b(const float* in, const float* z, int B, int nrows, int N, float* out, double* xx_sum) {
  MidMulConjZ mid{Src<cpx>{(const cpx*)z, (long long)nrows * N, nullptr, 0, nullptr, 0}, xx_sum};
#define X(n, r0, r1, r2, l, t, lr) \
  if (N == n) { emu_rowrow<n, r0, r1, r2, lr, t>((const cpx*)in, nrows, B, (cpx*)out, mid); return 0; }
  NIS_ROW_PLANS_B(X)
#undef X
  return -1;
}
int emu_rowrow_filter_b(const float* in, const float* h, const unsigned int* maxbuf, int B, int nrows, int N, float* out) {
  MidFilterH mid{Src<cpx>{(const cpx*)h, (long long)nrows * N, nullptr, 0, nullptr, 0}, maxbuf};
#define X(n, r0, r1, r2, l, t, lr) \
  if (N == n) { emu_rowrow<n, r0, r1, r2, lr, t>((const cpx*)in, nrows, B, (cpx*)out, mid); return 0; }
  NIS_ROW_PLANS_B(X)
#undef X
  return -1;
}

// in-register DFT check: v (R complex) -> DFT
int emu_dft(float* v, int R, int inverse) {
#define D(r) if (R == r) { if (inverse) Dft<r, true>::run((cpx*)v); else Dft<r, false>::run((cpx*)v); return 0; }
  D(1) D(2) D(3) D(4) D(5) D(6) D(8) D(9) D(10) D(12) D(15) D(16) D(20) D(32)
#undef D
  return -1;
}
}
