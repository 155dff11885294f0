// nis_fft.cuh -- shared-memory mixed-radix FFT building blocks for the KCC hot path (sm_100a).
//
// Replaces the reference's FFTW3f calls (src/correlation_flow.cc:53-77: fftwf_plan_dft_r2c_2d / c2r_2d,
// planned and destroyed per call) with two batched kernel families:
//   * column pass  (transform along image rows r, the halved dimension): one CTA owns a tile of 32 adjacent real
//     columns = 16 complex lines (two real columns ride as re/im of one complex line), lanes <-> lines, so every
//     shared-memory access is conflict-free and every global access is a 128..256 B contiguous segment.
//   * row pass     (transform along image columns c, contiguous): lanes <-> butterfly index, first radix 16,
//     one pad slot per 16 complex keeps the Stockham exchanges conflict-free.
// Both families do their first stage straight from global memory and the last stage straight to global memory; only two
// exchanges go through shared memory.  The column family runs in place on a digit-addressed tile (see COLUMN PASS).  Element-wise work of the KCC (spectrum products, |F|, the kernel function,
// the H*Kxz filter, arg-max / sum / sum-of-squares) rides in the prologue / epilogue functors of these passes.
//
// The bodies are written as barrier-free "phases" so the same code runs on the host under tests/cpp (thread-by-
// thread emulation of a CTA) -- there is no GPU in the build container.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NIS_HD __host__ __device__ __forceinline__
#define NIS_HDC __host__ __device__ constexpr
#define NIS_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define NIS_HD_NOINLINE static
#define NIS_HD inline
#define NIS_HDC constexpr
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#endif

namespace nis {

typedef float2 cpx;

// ---------------------------------------------------------------------------------------------------------
// complex primitives.  On sm_100a every one of them is ONE or TWO packed f32x2 instructions (FADD2 / FMUL2 / FFMA2 work on an
// aligned register pair = one complex number; operand modifiers swap the halves, negate one half or broadcast a scalar,
// so multiplying by +-i and conjugating are free).  Measured on B200 (tools/ubench/f32x2.cu): FFMA2 has the flop rate of FFMA
// at half the issue slots, and the FFT passes are issue bound.  The host build (tests/cpp emulation) uses the scalar forms.
// ---------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__) && !defined(NIS_NO_F32X2)
typedef unsigned long long pk64;
NIS_HD pk64 pk2(float x, float y) { pk64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
NIS_HD cpx upk2(pk64 v) { cpx r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
NIS_HD pk64 add2(pk64 a, pk64 b) { pk64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
NIS_HD pk64 sub2(pk64 a, pk64 b) { pk64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
NIS_HD pk64 mul2(pk64 a, pk64 b) { pk64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
NIS_HD pk64 fma2(pk64 a, pk64 b, pk64 c) { pk64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
NIS_HD cpx cadd(cpx a, cpx b) { return upk2(add2(pk2(a.x, a.y), pk2(b.x, b.y))); }
NIS_HD cpx csub(cpx a, cpx b) { return upk2(sub2(pk2(a.x, a.y), pk2(b.x, b.y))); }
NIS_HD cpx cscale(cpx a, float s) { return upk2(mul2(pk2(s, s), pk2(a.x, a.y))); }
// s*a + b, s real
NIS_HD cpx cfma(float s, cpx a, cpx b) { return upk2(fma2(pk2(s, s), pk2(a.x, a.y), pk2(b.x, b.y))); }
NIS_HD cpx cmul(cpx a, cpx b) {                       // a*b: (a.x b.x - a.y b.y, a.y b.x + a.x b.y)
  const cpx t = upk2(mul2(pk2(b.y, b.y), pk2(a.y, a.x)));
  return upk2(fma2(pk2(b.x, b.x), pk2(a.x, a.y), pk2(-t.x, t.y)));
}
NIS_HD cpx cmulc(cpx a, cpx b) {                      // a*conj(b): (a.x b.x + a.y b.y, a.y b.x - a.x b.y)
  const cpx t = upk2(mul2(pk2(b.y, b.y), pk2(a.y, a.x)));
  return upk2(fma2(pk2(b.x, b.x), pk2(a.x, a.y), pk2(t.x, -t.y)));
}
#else
NIS_HD cpx cadd(cpx a, cpx b) { return make_float2(a.x + b.x, a.y + b.y); }
NIS_HD cpx csub(cpx a, cpx b) { return make_float2(a.x - b.x, a.y - b.y); }
NIS_HD cpx cscale(cpx a, float s) { return make_float2(a.x * s, a.y * s); }
NIS_HD cpx cfma(float s, cpx a, cpx b) { return make_float2(s * a.x + b.x, s * a.y + b.y); }
NIS_HD cpx cmul(cpx a, cpx b) { return make_float2(a.x * b.x - a.y * b.y, a.y * b.x + a.x * b.y); }
NIS_HD cpx cmulc(cpx a, cpx b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a*conj(b)
#endif
NIS_HD cpx cconj(cpx a) { return make_float2(a.x, -a.y); }
// multiply by -i (forward) / +i (inverse)
template <bool INV> NIS_HD cpx mul_mi(cpx a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
// x times the stored (forward, exp(-i..)) twiddle, or its conjugate for the inverse transform
template <bool INV> NIS_HD cpx ctw(cpx x, cpx w) { return INV ? cmulc(x, w) : cmul(x, w); }

// ---------------------------------------------------------------------------------------------------------
// in-register DFTs.  dft<R,INV>(v): v[k] <- sum_n v[n] exp(-/+ 2 pi i n k / R); written with the complex primitives only
// ---------------------------------------------------------------------------------------------------------
template <int R, bool INV> struct Dft;

template <bool INV> struct Dft<1, INV> { static NIS_HD void run(cpx*) {} };

template <bool INV> struct Dft<2, INV> {
  static NIS_HD void run(cpx* v) {
    cpx a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
  }
};

template <bool INV> struct Dft<3, INV> {
  static NIS_HD void run(cpx* v) {
    const float s3 = 0.86602540378443864676f;
    const cpx a = v[0], s = cadd(v[1], v[2]), d = csub(v[1], v[2]);
    const cpx m = cfma(-0.5f, s, a);
    const cpx t = mul_mi<INV>(cscale(d, s3));
    v[0] = cadd(a, s); v[1] = cadd(m, t); v[2] = csub(m, t);
  }
};

template <bool INV> struct Dft<4, INV> {
  static NIS_HD void run(cpx* v) {
    const cpx s0 = cadd(v[0], v[2]), s1 = csub(v[0], v[2]), s2 = cadd(v[1], v[3]), s3 = mul_mi<INV>(csub(v[1], v[3]));
    v[0] = cadd(s0, s2); v[2] = csub(s0, s2); v[1] = cadd(s1, s3); v[3] = csub(s1, s3);
  }
};

template <bool INV> struct Dft<5, INV> {
  static NIS_HD void run(cpx* v) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const cpx a = v[0];
    const cpx p1 = cadd(v[1], v[4]), m1 = csub(v[1], v[4]), p2 = cadd(v[2], v[3]), m2 = csub(v[2], v[3]);
    const cpx t1 = cfma(c2, p2, cfma(c1, p1, a));
    const cpx t2 = cfma(c1, p2, cfma(c2, p1, a));
    const cpx u = mul_mi<INV>(cfma(s2, m2, cscale(m1, s1)));
    const cpx w = mul_mi<INV>(cfma(-s1, m2, cscale(m1, s2)));
    v[0] = cadd(cadd(a, p1), p2);
    v[1] = cadd(t1, u); v[4] = csub(t1, u); v[2] = cadd(t2, w); v[3] = csub(t2, w);
  }
};

// radix 8 = 2 x 4 with W8 twiddles:  w8^1 x = h (x + (-/+ i) x),  w8^3 x = -h (x + (+/- i) x)
template <bool INV> struct Dft<8, INV> {
  static NIS_HD void run(cpx* v) {
    const float h = 0.70710678118654752440f;
    cpx e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    Dft<4, INV>::run(e); Dft<4, INV>::run(o);
    const cpx q1 = cadd(o[1], mul_mi<INV>(o[1]));
    const cpx o2 = mul_mi<INV>(o[2]);
    const cpx q3 = cadd(o[3], mul_mi<!INV>(o[3]));
    v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
    v[1] = cfma(h, q1, e[1]); v[5] = cfma(-h, q1, e[1]);
    v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
    v[3] = cfma(-h, q3, e[3]); v[7] = cfma(h, q3, e[3]);
  }
};

// radix 16 = 2 x 8 with W16 twiddles
template <bool INV> struct Dft<16, INV> {
  static NIS_HD void run(cpx* v) {
    const float h = 0.70710678118654752440f, c = 0.92387953251128675613f, s = 0.38268343236508977173f;
    cpx e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    Dft<8, INV>::run(e); Dft<8, INV>::run(o);
    // forward twiddles w16^k = (wr, wi): k=1:(c,-s) 2:(h,-h) 3:(s,-c) 4:(0,-1) 5:(-s,-c) 6:(-h,-h) 7:(-c,-s)
    const float wr[8] = {1.f, c, h, s, 0.f, -s, -h, -c};
    const float wi[8] = {0.f, -s, -h, -c, -1.f, -c, -h, -s};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      cpx t;
      if (k == 0) t = o[0];
      else if (k == 4) t = mul_mi<INV>(o[4]);
      else t = ctw<INV>(o[k], make_float2(wr[k], wi[k]));
      v[k] = cadd(e[k], t); v[k + 8] = csub(e[k], t);
    }
  }
};

// radix 32 = 2 x 16 with W32 twiddles (first stage of the two-stage row plans: 32 independent loads per thread in flight)
template <bool INV> struct Dft<32, INV> {
  static NIS_HD void run(cpx* v) {
    const float h = 0.70710678118654752440f;
    const float c1 = 0.98078528040323044913f, s1 = 0.19509032201612826785f;   // cos/sin(pi/16)
    const float c2 = 0.92387953251128675613f, s2 = 0.38268343236508977173f;   // cos/sin(pi/8)
    const float c3 = 0.83146961230254523708f, s3 = 0.55557023301960222474f;   // cos/sin(3pi/16)
    cpx e[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    Dft<16, INV>::run(e); Dft<16, INV>::run(o);
    // forward twiddles w32^k = (cos, -sin)(2 pi k / 32), k = 0..15
    const float wr[16] = {1.f, c1, c2, c3, h, s3, s2, s1, 0.f, -s1, -s2, -s3, -h, -c3, -c2, -c1};
    const float wi[16] = {0.f, -s1, -s2, -s3, -h, -c3, -c2, -c1, -1.f, -c1, -c2, -c3, -h, -s3, -s2, -s1};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      cpx t;
      if (k == 0) t = o[0];
      else if (k == 8) t = mul_mi<INV>(o[8]);
      else t = ctw<INV>(o[k], make_float2(wr[k], wi[k]));
      v[k] = cadd(e[k], t); v[k + 16] = csub(e[k], t);
    }
  }
};

// radix 9 = 3 x 3 with W9 twiddles
template <bool INV> struct Dft<9, INV> {
  static NIS_HD void run(cpx* v) {
    const float c1 = 0.76604444311897803520f, s1 = 0.64278760968653932632f;   // cos/sin(2pi/9)
    const float c2 = 0.17364817766693034885f, s2 = 0.98480775301220805937f;   // cos/sin(4pi/9)
    const float c4 = -0.93969262078590838405f, s4 = 0.34202014332566873304f;  // cos/sin(8pi/9)
    cpx y[3][3];
#pragma unroll
    for (int n2 = 0; n2 < 3; ++n2) {
      cpx t[3] = {v[n2], v[3 + n2], v[6 + n2]};
      Dft<3, INV>::run(t);
      y[n2][0] = t[0]; y[n2][1] = t[1]; y[n2][2] = t[2];
    }
    y[1][1] = ctw<INV>(y[1][1], make_float2(c1, -s1));
    y[1][2] = ctw<INV>(y[1][2], make_float2(c2, -s2));
    y[2][1] = ctw<INV>(y[2][1], make_float2(c2, -s2));
    y[2][2] = ctw<INV>(y[2][2], make_float2(c4, -s4));
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1) {
      cpx t[3] = {y[0][k1], y[1][k1], y[2][k1]};
      Dft<3, INV>::run(t);
      v[k1] = t[0]; v[k1 + 3] = t[1]; v[k1 + 6] = t[2];
    }
  }
};

// Good-Thomas prime-factor DFT for coprime N1*N2 (no twiddles, only index maps resolved at compile time)
NIS_HDC int cx_modinv(int a, int m) {
  for (int x = 1; x < m; ++x)
    if ((a * x) % m == 1) return x;
  return 1;
}
template <int N1, int N2, bool INV> struct DftPfa {
  static NIS_HD void run(cpx* v) {
    constexpr int N = N1 * N2;
    constexpr int A = N2 * cx_modinv(N2 % N1, N1), B = N1 * cx_modinv(N1 % N2, N2);
    cpx t[N];
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) {
#pragma unroll
      for (int n1 = 0; n1 < N1; ++n1) t[n2 * N1 + n1] = v[(N2 * n1 + N1 * n2) % N];
      Dft<N1, INV>::run(t + n2 * N1);
    }
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) {
      cpx u[N2];
#pragma unroll
      for (int n2 = 0; n2 < N2; ++n2) u[n2] = t[n2 * N1 + k1];
      Dft<N2, INV>::run(u);
#pragma unroll
      for (int k2 = 0; k2 < N2; ++k2) v[(k1 * A + k2 * B) % N] = u[k2];
    }
  }
};
template <bool INV> struct Dft<6, INV> { static NIS_HD void run(cpx* v) { DftPfa<2, 3, INV>::run(v); } };
template <bool INV> struct Dft<10, INV> { static NIS_HD void run(cpx* v) { DftPfa<2, 5, INV>::run(v); } };
template <bool INV> struct Dft<12, INV> { static NIS_HD void run(cpx* v) { DftPfa<4, 3, INV>::run(v); } };
template <bool INV> struct Dft<15, INV> { static NIS_HD void run(cpx* v) { DftPfa<3, 5, INV>::run(v); } };
template <bool INV> struct Dft<20, INV> { static NIS_HD void run(cpx* v) { DftPfa<4, 5, INV>::run(v); } };

// ---------------------------------------------------------------------------------------------------------
// twiddle tables (built on the host in double, stored f32): for a 3-stage plan (R0,R1,R2), N = R0*R1*R2
//   tw1[(r-1)*R0 + k]       = exp(-2 pi i r k / (R0*R1)),  r in [1,R1), k in [0,R0)
//   tw2[(r-1)*R0*R1 + k]    = exp(-2 pi i r k / N),        r in [1,R2), k in [0,R0*R1)
// ---------------------------------------------------------------------------------------------------------
struct Twiddles {
  const cpx* tw1;
  const cpx* tw2;
};

#if defined(__CUDA_ARCH__)
#define NIS_LDG(p) __ldg(p)
#else
#define NIS_LDG(p) (*(p))
#endif

// =========================================================================================================
// COLUMN PASS: lanes <-> kColLanes complex lines (2*kColLanes real columns), T threads, G = T/kColLanes butterfly groups.
//
// In-place three-stage transform of length N = A*B*C on a digit-addressed shared-memory tile: the element with digits
// (d0 < A, d1 < B, d2 < C) of line l sits at smem[pos(d0,d1,d2) * kColLanes + l], pos = (d0*B + d1)*PADC + d2 with one pad slot
// per block of C when C is even (so that the radix-C stage, whose groups walk whole blocks, alternates bank halves).
//   decimation in frequency (natural index in, digit-reversed out), used by every pass that starts from global memory:
//     n = d0*BC + d1*C + d2  ->  stage A over d0 (then * W_N^(k0 (d1 C + d2)))  ->  stage B over d1 (then * W_BC^(k1 d2))
//     ->  stage C over d2  ->  k = k0 + A k1 + A B k2 at pos(k0,k1,k2)
//   decimation in time (digit-reversed in, natural out), the second half of the fused inverse->forward kernel:
//     n = d0 + A d1 + A B d2 at pos(d0,d1,d2)  ->  stage C over d2  ->  (* W_BC^(d1 e2)) stage B over d1
//     ->  (* W_N^(d0 (e1 C + e2))) stage A over d0  ->  k = e0*BC + e1*C + e2
// Every butterfly reads and writes the SAME R slots, so a stage needs no carry registers and no barrier between its reads and
// its writes: one __syncthreads per stage boundary, registers = one butterfly.  Any row permutation is free at the global ends
// (a row is an independent contiguous segment), which is what makes the digit-reversed order harmless.
// Twiddles (forward sign, conjugated for the inverse):  tw1[(b-1)*C + d2] = exp(-2 pi i b d2 / (B C)),  b in [1,B)
//                                                      tw2[(a-1)*BC + j] = exp(-2 pi i a j / N),         a in [1,A), j in [0,BC)
// =========================================================================================================
#ifndef NIS_COL_LANES
#define NIS_COL_LANES 8
#endif
constexpr int kColLanes = NIS_COL_LANES;          // complex lines per CTA
constexpr int kColTile = 2 * kColLanes;           // real image columns per CTA

// LN = complex lines per CTA (2 LN real columns); the forward pass is templated on it for the lane-count experiments recorded in
// nis_col.cu (NIS_ROT_LANES); production uses kColLanes everywhere
template <int N, int A, int B, int C, int T, int LN = kColLanes> struct ColGeom {
  static_assert(A * B * C == N, "bad factorisation");
  static_assert(T % LN == 0, "T must be a multiple of the lane count");
  static constexpr int G = T / LN;
  static constexpr int BC = B * C, AB = A * B, AC = A * C;
  static constexpr int PADC = C + ((C % 2 == 0) ? 1 : 0);
  static constexpr int SLOTS = AB * PADC;
  static constexpr size_t kSmemBytes = sizeof(cpx) * (size_t)SLOTS * LN;
  static NIS_HD int pos(int d0, int d1, int d2) { return (d0 * B + d1) * PADC + d2; }
  static NIS_HD int pos_j(int d0, int j) { return (d0 * B + j / C) * PADC + j % C; }      // j = d1*C + d2
};

// ---- stage B (middle), both directions.  DIT = false: decimation in frequency (twiddle after the butterfly); true: before.
template <int N, int A, int B, int C, int T, bool INV, bool DIT, int LN = kColLanes>
NIS_HD void col_stage_b(int tid, cpx* smem, const Twiddles& twd) {
  typedef ColGeom<N, A, B, C, T, LN> Gm;
  const int l = tid % LN, gi = tid / LN;
  // butterfly index bi = d2 + C*d0: when the group count is a multiple of C the twiddle index d2 is the same in every round
  constexpr bool kHoist = (Gm::G % C == 0);
  cpx twv[B > 1 ? B - 1 : 1];
  if (kHoist) {
#pragma unroll
    for (int r = 1; r < B; ++r) twv[r - 1] = NIS_LDG(&twd.tw1[(r - 1) * C + gi % C]);
  }
  for (int bi = gi; bi < Gm::AC; bi += Gm::G) {
    const int d0 = bi / C, d2 = bi % C;
    cpx* s = smem + (size_t)Gm::pos(d0, 0, d2) * LN + l;
    cpx v[B];
#pragma unroll
    for (int r = 0; r < B; ++r) {
      cpx x = s[r * Gm::PADC * LN];
      if (DIT && r > 0) x = ctw<INV>(x, kHoist ? twv[r - 1] : NIS_LDG(&twd.tw1[(r - 1) * C + d2]));
      v[r] = x;
    }
    Dft<B, INV>::run(v);
#pragma unroll
    for (int r = 0; r < B; ++r) {
      cpx x = v[r];
      if (!DIT && r > 0) x = ctw<INV>(x, kHoist ? twv[r - 1] : NIS_LDG(&twd.tw1[(r - 1) * C + d2]));
      s[r * Gm::PADC * LN] = x;
    }
  }
}

// ---- forward r2c column pass ------------------------------------------------------------------------------
// Pro::lane(l).load_all<R>(row0, stride, v): v[r] = (real column c0+2l, real column c0+2l+1) at image row row0 + r*stride;
// the per-lane context lets a prologue hoist everything that depends only on the column pair, and handing it all R rows at
// once lets gather-type prologues issue every independent load before the first use
// stage A from global (decimation in frequency)
template <int N, int A, int B, int C, int T, class Pro, int LN = kColLanes>
NIS_HD void col_fwd_stage_a(int tid, cpx* smem, const Twiddles& twd, const Pro& pro) {
  typedef ColGeom<N, A, B, C, T, LN> Gm;
  const int l = tid % LN, gi = tid / LN;
  const auto ln = pro.lane(l);
  for (int j = gi; j < Gm::BC; j += Gm::G) {
    cpx v[A];
    ln.template load_all<A>(j, Gm::BC, v);        // v[r] = sample at row j + r*BC
    Dft<A, false>::run(v);
    cpx* s = smem + (size_t)Gm::pos_j(0, j) * LN + l;
#pragma unroll
    for (int r = 0; r < A; ++r) {
      cpx x = v[r];
      if (r > 0) x = cmul(x, NIS_LDG(&twd.tw2[(r - 1) * Gm::BC + j]));
      s[r * B * Gm::PADC * LN] = x;
    }
  }
}

// separation of the two real lines packed in Z:  A[k] = (Z[k]+conj Z[N-k])/2,  B[k] = (Z[k]-conj Z[N-k])/(2i)
NIS_HD float4 r2c_split(cpx zk, cpx zn) {
  const cpx a = cscale(cadd(zk, cconj(zn)), 0.5f), b = cscale(mul_mi<false>(csub(zk, cconj(zn))), 0.5f);
  return make_float4(a.x, a.y, b.x, b.y);
}

// last stage of a forward pass done by butterfly pairs (p, NS2-p): butterfly p yields Z[p + r*NS2] in v[r], butterfly NS2-p
// yields Z[N - (p + r*NS2)] in u[R2-1-r], so Z[k] and Z[N-k] separate in registers.  Writes spectrum rows k in [0,N/2].
template <int N, int NS2, int R2>
NIS_HD void r2c_emit(int p, bool single, const cpx* v, const cpx* u, float4* out4, int pitch, int c0, int l) {
#pragma unroll
  for (int r = 0; r < R2; ++r) {
    const int k = p + r * NS2;
    cpx zk, zn;
    int kk;
    if (single) {
      if (2 * k > N) continue;
      kk = k;
      zk = v[r];
      zn = (p == 0) ? v[(R2 - r) % R2] : v[R2 - 1 - r];
    } else if (2 * k <= N) {
      kk = k; zk = v[r]; zn = u[R2 - 1 - r];
    } else {
      kk = N - k; zk = u[R2 - 1 - r]; zn = v[r];
    }
    out4[((size_t)kk * pitch + c0) / 2 + l] = r2c_split(zk, zn);
  }
}

// stage C of the decimation-in-frequency forward pass: butterfly q = k0 + A*k1 yields k = q + A*B*k2
// out: spectrum rows k in [0,N/2], row pitch `pitch` complex; this CTA's columns start at c0 (even).
template <int N, int A, int B, int C, int T, int LN = kColLanes>
NIS_HD void col_fwd_stage_c(int tid, const cpx* smem, cpx* out, int pitch, int c0) {
  typedef ColGeom<N, A, B, C, T, LN> Gm;
  constexpr int AB = Gm::AB;
  const int l = tid % LN, gi = tid / LN;
  float4* out4 = reinterpret_cast<float4*>(out);   // (A.x,A.y,B.x,B.y) = two adjacent complex columns
  for (int p = gi; 2 * p <= AB; p += Gm::G) {
    const bool single = (p == 0) || (2 * p == AB);
    cpx v[C], u[C];
    const cpx* s = smem + (size_t)Gm::pos(p % A, p / A, 0) * LN + l;
#pragma unroll
    for (int r = 0; r < C; ++r) v[r] = s[r * LN];
    Dft<C, false>::run(v);
    if (!single) {
      const int q = AB - p;
      const cpx* t = smem + (size_t)Gm::pos(q % A, q / A, 0) * LN + l;
#pragma unroll
      for (int r = 0; r < C; ++r) u[r] = t[r * LN];
      Dft<C, false>::run(u);
    }
    r2c_emit<N, AB, C>(p, single, v, u, out4, pitch, c0, l);
  }
}

// stage A of the decimation-in-time forward pass (second half of the fused kernel): butterfly j = e1*C + e2 yields k = j + BC*e0
template <int N, int A, int B, int C, int T>
NIS_HD void col_fwd_dit_stage_a(int tid, const cpx* smem, const Twiddles& twd, cpx* out, int pitch, int c0) {
  typedef ColGeom<N, A, B, C, T> Gm;
  constexpr int BC = Gm::BC;
  const int l = tid % kColLanes, gi = tid / kColLanes;
  float4* out4 = reinterpret_cast<float4*>(out);
  for (int p = gi; 2 * p <= BC; p += Gm::G) {
    const bool single = (p == 0) || (2 * p == BC);
    cpx v[A], u[A];
    const cpx* s = smem + (size_t)Gm::pos_j(0, p) * kColLanes + l;
#pragma unroll
    for (int r = 0; r < A; ++r) {
      cpx x = s[r * B * Gm::PADC * kColLanes];
      if (r > 0) x = cmul(x, NIS_LDG(&twd.tw2[(r - 1) * BC + p]));
      v[r] = x;
    }
    Dft<A, false>::run(v);
    if (!single) {
      const int q = BC - p;
      const cpx* t = smem + (size_t)Gm::pos_j(0, q) * kColLanes + l;
#pragma unroll
      for (int r = 0; r < A; ++r) {
        cpx x = t[r * B * Gm::PADC * kColLanes];
        if (r > 0) x = cmul(x, NIS_LDG(&twd.tw2[(r - 1) * BC + q]));
        u[r] = x;
      }
      Dft<A, false>::run(u);
    }
    r2c_emit<N, BC, A>(p, single, v, u, out4, pitch, c0, l);
  }
}

// ---- inverse c2r column pass ------------------------------------------------------------------------------
// in: half spectrum rows k in [0,N/2] (pitch complex), this CTA's columns start at c0.
// stage A (inverse, decimation in frequency) with butterflies j and BC-j paired so one (A,B) load feeds Z[k] and Z[N-k].
template <int N, int A, int B, int C, int T>
NIS_HD void col_inv_stage_a(int tid, cpx* smem, const Twiddles& twd, const cpx* in, int pitch, int c0) {
  typedef ColGeom<N, A, B, C, T> Gm;
  constexpr int M0 = Gm::BC;
  const int l = tid % kColLanes, gi = tid / kColLanes;
  const float4* in4 = reinterpret_cast<const float4*>(in);
  for (int p = gi; 2 * p <= M0; p += Gm::G) {
    const bool single = (p == 0) || (2 * p == M0);
    cpx v[A], u[A];
#pragma unroll
    for (int r = 0; r < A; ++r) {
      const int k = p + r * M0;
      const bool lo = (2 * k <= N);
      const int kk = lo ? k : N - k;
      float4 ab = NIS_LDG(&in4[((size_t)kk * pitch + c0) / 2 + l]);
      if (kk == 0 || 2 * kk == N) { ab.y = 0.f; ab.w = 0.f; }     // c2r ignores these imaginary parts
      const cpx ca = make_float2(ab.x, ab.y), ib = mul_mi<true>(make_float2(ab.z, ab.w));
      const cpx zlo = cadd(ca, ib);                                // A + iB        = Z[kk]
      const cpx zhi = cconj(csub(ca, ib));                         // conjA + i conjB = Z[N-kk]
      v[r] = lo ? zlo : zhi;
      if (!single) u[A - 1 - r] = lo ? zhi : zlo;
    }
    Dft<A, true>::run(v);
    cpx* s = smem + (size_t)Gm::pos_j(0, p) * kColLanes + l;
#pragma unroll
    for (int r = 0; r < A; ++r) {
      cpx x = v[r];
      if (r > 0) x = cmulc(x, NIS_LDG(&twd.tw2[(r - 1) * M0 + p]));
      s[r * B * Gm::PADC * kColLanes] = x;
    }
    if (!single) {
      const int q = M0 - p;
      Dft<A, true>::run(u);
      cpx* t = smem + (size_t)Gm::pos_j(0, q) * kColLanes + l;
#pragma unroll
      for (int r = 0; r < A; ++r) {
        cpx x = u[r];
        if (r > 0) x = cmulc(x, NIS_LDG(&twd.tw2[(r - 1) * M0 + q]));
        t[r * B * Gm::PADC * kColLanes] = x;
      }
    }
  }
}

// stage C (inverse) -> epilogue.  Epi::put(row, l, re, im): re -> real column c0+2l, im -> c0+2l+1 (unnormalised).
// Block m = d0*B + d1 (position order, so adjacent groups alternate bank halves) yields rows k = d0 + A*d1 + A*B*k2.
template <int N, int A, int B, int C, int T, class Epi>
NIS_HD void col_inv_stage_c(int tid, const cpx* smem, Epi& epi) {
  typedef ColGeom<N, A, B, C, T> Gm;
  const int l = tid % kColLanes, gi = tid / kColLanes;
  for (int m = gi; m < Gm::AB; m += Gm::G) {
    const cpx* s = smem + (size_t)m * Gm::PADC * kColLanes + l;
    cpx v[C];
#pragma unroll
    for (int r = 0; r < C; ++r) v[r] = s[r * kColLanes];
    Dft<C, true>::run(v);
    const int row0 = m / B + A * (m % B);
#pragma unroll
    for (int r = 0; r < C; ++r) epi.put(row0 + r * Gm::AB, l, v[r].x, v[r].y);
  }
}

// fused middle of the inverse->forward kernel: inverse stage C, fn on both packed reals, forward (decimation-in-time) stage C,
// all on the same block of C slots in registers -- the real kernel image never exists anywhere else.
template <int N, int A, int B, int C, int T, class Fn>
NIS_HD void col_inv_fn_fwd_stage_c(int tid, cpx* smem, Fn& fn) {
  typedef ColGeom<N, A, B, C, T> Gm;
  const int l = tid % kColLanes, gi = tid / kColLanes;
  for (int m = gi; m < Gm::AB; m += Gm::G) {
    cpx* s = smem + (size_t)m * Gm::PADC * kColLanes + l;
    cpx v[C];
#pragma unroll
    for (int r = 0; r < C; ++r) v[r] = s[r * kColLanes];
    Dft<C, true>::run(v);
#pragma unroll
    for (int r = 0; r < C; ++r) v[r] = fn.apply(v[r]);
    Dft<C, false>::run(v);
#pragma unroll
    for (int r = 0; r < C; ++r) s[r * kColLanes] = v[r];
  }
}

// =========================================================================================================
// ROW PASS: contiguous complex lines of length N = R0*R1*R2 (Stockham, decimation in time), L lines per CTA, lanes <-> butterfly index.
// smem: cpx[L][N + N/R0] (one pad slot per R0, so that the R0 contiguous outputs of neighbouring stage-0 butterflies start R0+1
// slots apart; R0+1 must be odd or the plan pays bank conflicts).
// Global-memory alignment: stage 0 reads runs of M0 = N/R0 consecutive elements and the last stage writes runs of N/R2; a warp's
// request covers whole 128-byte lines exactly when both are multiples of 16 elements.  The production plans (R0 = 16: runs of
// 40 / 30; R0 = 32: runs of 20 / 15) do not, on purpose -- the passes are latency bound and what pays is the number of independent
// loads a thread has in flight in stage 0 (see nis_sizes.h for the measurements).  Plans with R2 == 1 take the two-stage path
// (row_stage1_out / row_stage1_mid): one exchange, one barrier.
// =========================================================================================================
template <int R1, int ROUNDS> struct CarryRegs { cpx v[ROUNDS][R1]; };

template <int N, int R0, int R1, int R2, int L, int T> struct RowGeom {
  static_assert(R0 * R1 * R2 == N, "bad factorisation");
  static constexpr int M0 = N / R0, M1 = N / R1, M2 = N / R2;
  static constexpr int NS1 = R0, NS2 = R0 * R1;
  static constexpr int PITCH = N + N / R0;
  static constexpr int ROUNDS1 = (L * M1 + T - 1) / T;
  static constexpr size_t kSmemBytes = sizeof(cpx) * (size_t)L * PITCH;
  static_assert(M1 % R0 == 0 && NS2 % R0 == 0, "stage strides must be multiples of the pad period");
  static NIS_HD int pad(int i) { return i + i / R0; }
};

// Pro::line(ln).load(c) -> cpx for local line index ln in [0,L) (caller guards valid lines); c in [0,N)
template <int N, int R0, int R1, int R2, int L, int T, bool INV, class Pro>
NIS_HD void row_phase0(int tid, cpx* smem, const Pro& pro, int nlines) {
  typedef RowGeom<N, R0, R1, R2, L, T> Gm;
  for (int w = tid; w < L * Gm::M0; w += T) {
    const int ln = w / Gm::M0, j = w % Gm::M0;
    if (ln >= nlines) break;
    cpx v[R0];
    const auto lc = pro.line(ln);
#pragma unroll
    for (int r = 0; r < R0; ++r) v[r] = lc.load(j + r * Gm::M0);
    Dft<R0, INV>::run(v);
    cpx* s = smem + ln * Gm::PITCH + j * (R0 + 1);     // pad(R0 j + r) = (R0 + 1) j + r
#pragma unroll
    for (int r = 0; r < R0; ++r) s[r] = v[r];
  }
}

template <int N, int R0, int R1, int R2, int L, int T, bool INV>
NIS_HD void row_stage1_read(int tid, const cpx* smem, const Twiddles& twd, int nlines, CarryRegs<R1, RowGeom<N, R0, R1, R2, L, T>::ROUNDS1>& st) {
  typedef RowGeom<N, R0, R1, R2, L, T> Gm;
  // k = j % R0 with j = (tid + it*T) % M1; M1 is a multiple of R0, so when T is one too k = tid % R0 in every round: the R1-1
  // twiddles are loaded once per thread instead of once per round (they were two thirds of this kernel's global loads)
  constexpr bool kHoist = (T % R0 == 0);
  cpx twv[R1 > 1 ? R1 - 1 : 1];
  if (kHoist) {
#pragma unroll
    for (int r = 1; r < R1; ++r) twv[r - 1] = NIS_LDG(&twd.tw1[(r - 1) * R0 + tid % R0]);
  }
#pragma unroll
  for (int it = 0; it < Gm::ROUNDS1; ++it) {
    const int w = tid + it * T;
    const int ln = w / Gm::M1, j = w % Gm::M1;
    if (w < L * Gm::M1 && ln < nlines) {
      // M1 is a multiple of R0, so pad(j + r*M1) = pad(j) + r*(M1 + M1/R0): one padded base, constant offsets
      const cpx* s = smem + ln * Gm::PITCH + Gm::pad(j);
#pragma unroll
      for (int r = 0; r < R1; ++r) {
        cpx x = s[r * (Gm::M1 + Gm::M1 / R0)];
        if (r > 0) x = ctw<INV>(x, kHoist ? twv[r - 1] : NIS_LDG(&twd.tw1[(r - 1) * R0 + j % R0]));
        st.v[it][r] = x;
      }
      Dft<R1, INV>::run(st.v[it]);
    }
  }
}
template <int N, int R0, int R1, int R2, int L, int T, bool INV>
NIS_HD void row_stage1_write(int tid, cpx* smem, int nlines, const CarryRegs<R1, RowGeom<N, R0, R1, R2, L, T>::ROUNDS1>& st) {
  typedef RowGeom<N, R0, R1, R2, L, T> Gm;
#pragma unroll
  for (int it = 0; it < Gm::ROUNDS1; ++it) {
    const int w = tid + it * T;
    const int ln = w / Gm::M1, j = w % Gm::M1;
    if (w < L * Gm::M1 && ln < nlines) {
      const int k = j % R0, j0 = (j / R0) * R0 * R1 + k;
      cpx* s = smem + ln * Gm::PITCH + Gm::pad(j0);      // pad(j0 + R0 r) = pad(j0) + (R0 + 1) r
#pragma unroll
      for (int r = 0; r < R1; ++r) s[r * (R0 + 1)] = st.v[it][r];
    }
  }
}

// Epi::line(ln).put(c, value)
template <int N, int R0, int R1, int R2, int L, int T, bool INV, class Epi>
NIS_HD void row_phase2(int tid, const cpx* smem, const Twiddles& twd, int nlines, Epi& epi) {
  typedef RowGeom<N, R0, R1, R2, L, T> Gm;
  constexpr int NS2 = Gm::NS2;
  // when T is a multiple of NS2 the butterfly index j = w % NS2 is the same in every round: hoist the R2-1 twiddles
  constexpr bool kHoist = (T % NS2 == 0);
  cpx twv[R2 > 1 ? R2 - 1 : 1];
  if (kHoist) {
#pragma unroll
    for (int r = 1; r < R2; ++r) twv[r - 1] = NIS_LDG(&twd.tw2[(r - 1) * NS2 + tid % NS2]);
  }
  for (int w = tid; w < L * NS2; w += T) {
    const int ln = w / NS2, j = w % NS2;
    if (ln >= nlines) break;
    const cpx* s = smem + ln * Gm::PITCH + Gm::pad(j);    // NS2 is a multiple of R0: pad(j + r*NS2) = pad(j) + r*(NS2 + NS2/R0)
    cpx v[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) {
      cpx x = s[r * (NS2 + NS2 / R0)];
      if (r > 0) x = ctw<INV>(x, kHoist ? twv[r - 1] : NIS_LDG(&twd.tw2[(r - 1) * NS2 + j]));
      v[r] = x;
    }
    Dft<R2, INV>::run(v);
    const auto lc = epi.line(ln);
#pragma unroll
    for (int r = 0; r < R2; ++r) lc.put(j + r * NS2, v[r]);
  }
}

// forward phase 2 kept in shared memory with an element-wise step (fused forward -> element-wise -> inverse):
// Mid::line(ln).apply(c, value) -> value; every thread overwrites exactly the padded slots it read.
template <int N, int R0, int R1, int R2, int L, int T, class Mid>
NIS_HD void row_phase2_mid(int tid, cpx* smem, const Twiddles& twd, int nlines, Mid& mid) {
  typedef RowGeom<N, R0, R1, R2, L, T> Gm;
  constexpr int NS2 = Gm::NS2;
  constexpr bool kHoist = (T % NS2 == 0);                  // see row_phase2
  cpx twv[R2 > 1 ? R2 - 1 : 1];
  if (kHoist) {
#pragma unroll
    for (int r = 1; r < R2; ++r) twv[r - 1] = NIS_LDG(&twd.tw2[(r - 1) * NS2 + tid % NS2]);
  }
  for (int w = tid; w < L * NS2; w += T) {
    const int ln = w / NS2, j = w % NS2;
    if (ln >= nlines) break;
    cpx* s = smem + ln * Gm::PITCH + Gm::pad(j);
    cpx v[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) {
      cpx x = s[r * (NS2 + NS2 / R0)];
      if (r > 0) x = cmul(x, kHoist ? twv[r - 1] : NIS_LDG(&twd.tw2[(r - 1) * NS2 + j]));
      v[r] = x;
    }
    Dft<R2, false>::run(v);
    auto lc = mid.line(ln);
#pragma unroll
    for (int r = 0; r < R2; ++r) s[r * (NS2 + NS2 / R0)] = lc.apply(j + r * NS2, v[r]);
    lc.flush();
  }
}

// ---- two-stage plans (R2 == 1, N = R0*R1): ONE shared-memory exchange and ONE barrier per pass.  Stage-1 butterfly j in [0,R0) of
// a line reads the padded slots pad(j) + r (R0+1) (= stage-0 outputs k0 = j of the sub-sequences r) and yields the natural indices
// j + R0 r, which go straight to the epilogue (or, in the fused kernel, back into the slots they were read from).
template <int N, int R0, int R1, int L, int T, bool INV, class Epi>
NIS_HD void row_stage1_out(int tid, const cpx* smem, const Twiddles& twd, int nlines, Epi& epi) {
  typedef RowGeom<N, R0, R1, 1, L, T> Gm;
  constexpr bool kHoist = (T % R0 == 0);
  cpx twv[R1 > 1 ? R1 - 1 : 1];
  if (kHoist) {
#pragma unroll
    for (int r = 1; r < R1; ++r) twv[r - 1] = NIS_LDG(&twd.tw1[(r - 1) * R0 + tid % R0]);
  }
  for (int w = tid; w < L * R0; w += T) {
    const int ln = w / R0, j = w % R0;
    if (ln >= nlines) break;
    const cpx* s = smem + ln * Gm::PITCH + j;              // pad(j) = j for j < R0
    cpx v[R1];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      cpx x = s[r * (R0 + 1)];
      if (r > 0) x = ctw<INV>(x, kHoist ? twv[r - 1] : NIS_LDG(&twd.tw1[(r - 1) * R0 + j]));
      v[r] = x;
    }
    Dft<R1, INV>::run(v);
    const auto lc = epi.line(ln);
#pragma unroll
    for (int r = 0; r < R1; ++r) lc.put(j + R0 * r, v[r]);
  }
}
template <int N, int R0, int R1, int L, int T, class Mid>
NIS_HD void row_stage1_mid(int tid, cpx* smem, const Twiddles& twd, int nlines, Mid& mid) {
  typedef RowGeom<N, R0, R1, 1, L, T> Gm;
  constexpr bool kHoist = (T % R0 == 0);
  cpx twv[R1 > 1 ? R1 - 1 : 1];
  if (kHoist) {
#pragma unroll
    for (int r = 1; r < R1; ++r) twv[r - 1] = NIS_LDG(&twd.tw1[(r - 1) * R0 + tid % R0]);
  }
  for (int w = tid; w < L * R0; w += T) {
    const int ln = w / R0, j = w % R0;
    if (ln >= nlines) break;
    cpx* s = smem + ln * Gm::PITCH + j;
    cpx v[R1];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      cpx x = s[r * (R0 + 1)];
      if (r > 0) x = cmul(x, kHoist ? twv[r - 1] : NIS_LDG(&twd.tw1[(r - 1) * R0 + j]));
      v[r] = x;
    }
    Dft<R1, false>::run(v);
    auto lc = mid.line(ln);
#pragma unroll
    for (int r = 0; r < R1; ++r) s[r * (R0 + 1)] = lc.apply(j + R0 * r, v[r]);      // natural index j + R0 r sits at pad(j + R0 r) = j + (R0+1) r
    lc.flush();
  }
}

// prologue reading a natural-order padded line out of shared memory (input of the fused inverse pass)
template <int PITCH, int R0> struct SmemLinePro {
  const cpx* base;
  struct Line { const cpx* p; NIS_HD cpx load(int c) const { return p[c + c / R0]; } };
  NIS_HD Line line(int ln) const { return Line{base + ln * PITCH}; }
};

}  // namespace nis
