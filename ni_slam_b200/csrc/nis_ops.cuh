// nis_ops.cuh -- prologue / epilogue functors fused into the FFT passes (the element-wise arithmetic of
// src/correlation_flow.cc:89-95, :145-243) plus the batch operand descriptors.
#pragma once
#include <math.h>
#include "nis_fft.cuh"

namespace nis {

// Batch operand: element b lives at base + b*stride, or (DB keyframes) at ptrs[idx[b]] + offset.
// `shift`: batch entry e reads element e >> shift (two rotation hypotheses of one pair share their operands).
template <class Tp> struct Src {
  const Tp* base;
  long long stride;
  const Tp* const* ptrs;
  long long offset;
  const int* idx;
  int shift;
  NIS_HD const Tp* at(int b) const {
    const int e = b >> shift;
    const int i = idx ? idx[e] : e;
    return ptrs ? ptrs[i] + offset : base + (long long)i * stride;
  }
};
template <class Tp> struct Dst {
  Tp* base;
  long long stride;
  NIS_HD Tp* at(int b) const { return base + (long long)b * stride; }
};

// order-preserving map float -> uint32 (for atomicMax based arg-max)
NIS_HD uint32_t f2ord(float f) {
#if defined(__CUDA_ARCH__)
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
NIS_HD float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
NIS_HD void atomic_add_f(float* p, float v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
NIS_HD void atomic_add_d(double* p, double v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
NIS_HD float bits2f(unsigned int u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

// key = (ordered value << 32) | (0xffffffff - column_major_index): max key = largest value, then the FIRST
// element in column-major order (Eigen maxCoeff on a column-major array, correlation_flow.cc:175).
NIS_HD unsigned long long peak_key(float v, int row, int col, int R) {
  return ((unsigned long long)f2ord(v) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(col * R + row));
}

struct PeakStats {            // per batch element, zero-initialised before the pass
  unsigned long long key;
  double sum, sumsq;
};

// ---------------------------------------------------------------------------------------------------------
// column-pass prologues: lane(l).load(row) -> (col c0+2l, col c0+2l+1) of a row-major real image
// ---------------------------------------------------------------------------------------------------------
struct ProRealF32 {
  Src<float> src; int W;
  struct Lane {
    const float2* p; int W;
    NIS_HD cpx load(int row) const { return NIS_LDG(p + (size_t)row * (W / 2)); }
    template <int R> NIS_HD void load_all(int row0, int stride, cpx* v) const {
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = load(row0 + r * stride);
    }
  };
  struct Bound {
    const float* img; int W, c0;
    NIS_HD Lane lane(int l) const { return Lane{reinterpret_cast<const float2*>(img + c0) + l, W}; }
  };
  NIS_HD Bound bind(int b, int c0) const { return Bound{src.at(b), W, c0}; }
};

// u8 -> f32/255 (utils.cc:110-118: (float)((double)u/255.0)).  The correctly rounded f32 quotient u/255.f equals the double
// quotient rounded to f32 for all 256 inputs, and so does one FMA-corrected step from u * fl(1/255) (both checked exhaustively
// in exact rational arithmetic by the CPU test suite), so the conversion is three FP instructions: a table lookup per tap costs
// the L1 data pipe -- the bottleneck of these kernels -- up to 8 wavefronts per warp.
NIS_HD float u8_to_unit(unsigned int u) {
#if defined(__CUDA_ARCH__)
  const float x = (float)u, r = 1.0f / 255.0f;
  const float q = __fmul_rn(x, r);
  return __fmaf_rn(__fmaf_rn(-q, 255.0f, x), r, q);
#else
  return (float)u / 255.0f;
#endif
}

// u8 image -> f32/255
struct ProRealU8 {
  Src<uint8_t> src; int W; const float* lut;
  struct Lane {
    const uint8_t* p; int W; const float* lut;
    NIS_HD cpx load(int row) const {
      const unsigned int w = NIS_LDG(reinterpret_cast<const unsigned short*>(p + (size_t)row * W));   // columns c0+2l (low byte), c0+2l+1
      return make_float2(u8_to_unit(w & 0xffu), u8_to_unit(w >> 8));
    }
    template <int R> NIS_HD void load_all(int row0, int stride, cpx* v) const {
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = load(row0 + r * stride);
    }
  };
  struct Bound {
    const uint8_t* img; int W, c0; const float* lut;
    NIS_HD Lane lane(int l) const { return Lane{img + c0 + 2 * l, W, lut}; }
  };
  NIS_HD Bound bind(int b, int c0) const { return Bound{src.at(b), W, c0, lut}; }
};

// ---------------------------------------------------------------------------------------------------------
// column-pass (c2r) epilogues: put(row, l, re, im), values unnormalised (divide by n = R*C like IFFT :76)
// ---------------------------------------------------------------------------------------------------------
// The reference divides by n = R*C (IFFT, :76); the kernels multiply by the f32 reciprocal instead (one FMUL instead of
// an IEEE division sequence per element; the results differ by at most 1 ulp, far below the f32 FFT noise).
struct EpiStore {
  Dst<float> dst; int W; float n;
  struct Bound {
    float* img; int W, c0; float rn;
    NIS_HD void put(int row, int l, float re, float im) {
      reinterpret_cast<float2*>(img + (size_t)row * W + c0)[l] = cscale(make_float2(re, im), rn);
    }
    template <class Sync> NIS_HD void finish(int, Sync&) {}
  };
  NIS_HD Bound bind(int b, int c0) const { return Bound{dst.at(b), W, c0, 1.0f / n}; }
};

// power = IFFT(|F|) stored fftshift-ed (circ_shift.h:238-244: out(r, c) = in((r - R/2) mod R, (c - C/2) mod C)), the layout the TMA-staged
// polar gather reads; the column pair (c0+2l, c0+2l+1) stays adjacent because c0 and C/2 are multiples of 16
struct EpiStoreShift {
  Dst<float> dst; int H, W; float n;
  struct Bound {
    float* img; int H, W, cs; float rn;
    NIS_HD void put(int row, int l, float re, float im) {
      const int y = row + H / 2 - (row + H / 2 >= H ? H : 0);
      reinterpret_cast<float2*>(img + (size_t)y * W + cs)[l] = cscale(make_float2(re, im), rn);
    }
    template <class Sync> NIS_HD void finish(int, Sync&) {}
  };
  NIS_HD Bound bind(int b, int c0) const { const int x = c0 + W / 2 - (c0 + W / 2 >= W ? W : 0); return Bound{dst.at(b), H, W, x, 1.0f / n}; }
};

// rare kernel-function paths (general integer power, gaussian): kept out of line so the 20 inlined call sites of the fused
// column kernel stay small (the cubic of every shipped config is the inline fast path)
NIS_HD_NOINLINE float kernel_fn_slow(float xz, int kernel, float offset, int power, float gcoef, float xxzz, float n);

NIS_HD float powi_double(float x, int p) {
  // Eigen 3.3 ArrayBase::pow(int) -> std::pow(float,int) -> double pow, rounded to float (correlation_flow.cc:213,223)
  double b = (double)x, r = 1.0;
  int e = p < 0 ? -p : p;
  for (int i = 0; i < e; ++i) r *= b;
  if (p < 0) r = 1.0 / r;
  return (float)r;
}

NIS_HD_NOINLINE float kernel_fn_slow(float xz, int kernel, float offset, int power, float gcoef, float xxzz, float n) {
  if (kernel == 0) return powi_double(xz + offset, power);
  return expf(gcoef * ((xxzz - 2.f * xz) / n));
}

// kernel function applied between the fused inverse and forward column passes (colcol kernel); the real pairs stay in
// shared memory.  polynomial (:208-226): k = (xz/n + offset)^power; gaussian (:181-206): k = exp(-1/sigma^2 (xx + zz - 2 xz/n)/n);
// max|k| is reduced into maxbuf[b] and the division by it is deferred to the consumer.  Gaussian sums arrive as raw double sums over the half spectrum.
struct KernelFn {
  float n; int kernel; float offset; int power; float gcoef;
  const double* xx_sum; const double* zz_sum; int zz_shift;
  unsigned int* maxbuf;
  const int* xx_idx;            // optional: xx_sum[xx_idx[b]] (rotated-query cache of the scan)
  struct Bound {
    float n, rn; int kernel; float offset; int power; float gcoef, xxzz; unsigned int* maxp; float mx;
    NIS_HD float kfun(float v) const {
      const float xz = v * rn;
      if (kernel == 0 && power == 3) { const double b3 = (double)(xz + offset); return (float)(b3 * b3 * b3); }
      return kernel_fn_slow(xz, kernel, offset, power, gcoef, xxzz, n);
    }
    NIS_HD cpx apply(cpx v) {
      const float a = kfun(v.x), b = kfun(v.y);
      mx = fmaxf(mx, fmaxf(fabsf(a), fabsf(b)));
      return make_float2(a, b);
    }
    template <class Sync> NIS_HD void finish(int tid, Sync& sync) { sync.block_max_to(maxp, mx, tid); }
  };
  NIS_HD Bound bind(int b) const {
    float s = 0.f;
    if (kernel == 1) s = (float)xx_sum[xx_idx ? xx_idx[b] : b] / n + (float)zz_sum[b >> zz_shift] / n;
    return Bound{n, 1.0f / n, kernel, offset, power, gcoef, s, maxbuf + b, 0.f};
  }
};

// final response (:173-178, :238-243): arg-max (column-major first), sum and sum of squares of g = v/n; g never stored
struct EpiPeak {
  PeakStats* stats; int R; float n;
  float* g_out; long long g_stride; int W;            // optional debug store (nullptr in production)
  struct Bound {
    PeakStats* st; int R, c0; float rn; float* g; int W;
    unsigned long long key; float s, q;
    NIS_HD void one(int row, int col, float v) {
      const float x = v * rn;
      const unsigned long long k = peak_key(x, row, col, R);
      key = k > key ? k : key;
      s += x; q += x * x;
      if (g) g[(size_t)row * W + col] = x;
    }
    NIS_HD void put(int row, int l, float re, float im) {
      one(row, c0 + 2 * l, re);
      one(row, c0 + 2 * l + 1, im);
    }
    template <class Sync> NIS_HD void finish(int tid, Sync& sync) { sync.block_peak_to(st, key, (double)s, (double)q, tid); }
  };
  NIS_HD Bound bind(int b, int c0) const {
    return Bound{stats + b, R, c0, 1.0f / n, g_out ? g_out + (long long)b * g_stride : nullptr, W, 0ull, 0.f, 0.f};
  }
};

// ---------------------------------------------------------------------------------------------------------
// row-pass prologues / epilogue; local line ln of this CTA -> global line g = line0+ln -> (batch b, spectrum row k1)
// ---------------------------------------------------------------------------------------------------------
struct LineMap {
  int line0, nrows, W;     // nrows = R/2+1 spectrum rows per batch element
  NIS_HD void map(int ln, int& b, size_t& off, int& k1) const {
    const int g = line0 + ln;
    b = g / nrows; k1 = g - b * nrows;
    off = (size_t)k1 * W;
  }
};

// Each prologue exposes line(ln) -> per-line context (batch element, row offset, per-image scalars resolved once)
// whose load(c) produces the input sample at column c.
struct ProSpec {      // plain spectrum load
  Src<cpx> x;
  struct Line { const cpx* p; NIS_HD cpx load(int c) const { return NIS_LDG(p + c); } };
  struct Bound {
    Src<cpx> x; LineMap m;
    NIS_HD Line line(int ln) const { int b, k1; size_t off; m.map(ln, b, off, k1); return Line{x.at(b) + off}; }
  };
  NIS_HD Bound bind(const LineMap& m) const { return Bound{x, m}; }
};

// |z|^2 with one fixed rounding sequence: the auto form of the product prologue and the fused store-and-square step of the row
// kernel must agree bit for bit (a keyframe's H factor may be computed through either)
NIS_HD float sqmag(cpx z) { return fmaf(z.x, z.x, z.y * z.y); }

struct ProMulConj {   // x * conj(z)  (:210-211); auto form when x aliases z (:220-221)
  Src<cpx> x, z;
  struct Line {
    const cpx* px; const cpx* pz;
    NIS_HD cpx load(int c) const {
      const cpx x = NIS_LDG(px + c);
      if (px == pz) return make_float2(sqmag(x), 0.f);     // auto form z * conj(z): one load, imaginary part exactly 0
      return cmulc(x, NIS_LDG(pz + c));
    }
  };
  struct Bound {
    Src<cpx> x, z; LineMap m;
    NIS_HD Line line(int ln) const { int b, k1; size_t off; m.map(ln, b, off, k1); return Line{x.at(b) + off, z.at(b) + off}; }
  };
  NIS_HD Bound bind(const LineMap& m) const { return Bound{x, z, m}; }
};


// ---- element-wise steps of the fused row kernel (forward FFT -> mid -> inverse FFT); Line::apply(c, x) -> y
// y = x * conj(Z): X = FFT(rotated image) is consumed in registers and never stored (correlation_flow.cc:210-211).
// With the gaussian kernel the half-spectrum sum of |x^2| (:184) is accumulated on the side.
struct MidMulConjZ {
  Src<cpx> z; double* xx_sum;       // xx_sum != nullptr only with the gaussian kernel
  struct Line {
    const cpx* pz; float* acc; float part;
    NIS_HD cpx apply(int c, cpx x) {
      if (acc) part += x.x * x.x + x.y * x.y;          // |x^2| = |x|^2
      return cmulc(x, NIS_LDG(pz + c));
    }
    NIS_HD void flush() { if (acc) atomic_add_f(acc, part); }
  };
  struct Bound {
    Src<cpx> z; double* xx_sum; LineMap m; float* scratch;       // scratch: one f32 accumulator per local line (shared memory)
    NIS_HD Line line(int ln) const { int b, k1; size_t off; m.map(ln, b, off, k1); return Line{z.at(b) + off, xx_sum ? scratch + ln : nullptr, 0.f}; }
    NIS_HD void finish_line(int ln) const {
      if (!xx_sum) return;
      int b, k1; size_t off; m.map(ln, b, off, k1);
      atomic_add_d(xx_sum + b, (double)scratch[ln]);
    }
  };
  NIS_HD Bound bind(const LineMap& m, float* scratch) const { return Bound{z, xx_sum, m, scratch}; }
};

// y = H * x / max_xz with H = T/(Kzz/max_zz + lambda) cached per keyframe (G = H * Kxz, correlation_flow.cc:171-172)
struct MidFilterH {
  Src<cpx> h; const unsigned int* max_xz;
  struct Line {
    const cpx* ph; float ixz;
    NIS_HD cpx apply(int c, cpx x) const { return cmul(cscale(x, ixz), NIS_LDG(ph + c)); }
    NIS_HD void flush() const {}
  };
  struct Bound {
    Src<cpx> h; const unsigned int* max_xz; LineMap m;
    NIS_HD Line line(int ln) const { int b, k1; size_t off; m.map(ln, b, off, k1); return Line{h.at(b) + off, 1.0f / bits2f(max_xz[b])}; }
    NIS_HD void finish_line(int) const {}
  };
  NIS_HD Bound bind(const LineMap& m, float*) const { return Bound{h, max_xz, m}; }
};

// store P = x, continue with x * conj(x) = |x|^2: fft_polar leaves the forward row pass and the same kernel runs the row half of
// IFFT(P conj P), the first step of the keyframe factor Kzz of the polar stage (correlation_flow.cc:94 followed by :220-221)
struct MidStoreSq {
  Dst<cpx> f;
  struct Line {
    cpx* pf;
    NIS_HD cpx apply(int c, cpx x) const { pf[c] = x; return make_float2(sqmag(x), 0.f); }
    NIS_HD void flush() const {}
  };
  struct Bound {
    Dst<cpx> f; LineMap m;
    NIS_HD Line line(int ln) const { int b, k1; size_t off; m.map(ln, b, off, k1); return Line{f.at(b) + off}; }
    NIS_HD void finish_line(int) const {}
  };
  NIS_HD Bound bind(const LineMap& m, float*) const { return Bound{f, m}; }
};

// store F = x, continue with |x| (ComputeIntermedium: fft_result = FFT(image); IFFT(fft_result.abs()), :91-92)
struct MidStoreAbs {
  Dst<cpx> f;
  struct Line {
    cpx* pf;
    NIS_HD cpx apply(int c, cpx x) const { pf[c] = x; return make_float2(sqrtf(x.x * x.x + x.y * x.y), 0.f); }
    NIS_HD void flush() const {}
  };
  struct Bound {
    Dst<cpx> f; LineMap m;
    NIS_HD Line line(int ln) const { int b, k1; size_t off; m.map(ln, b, off, k1); return Line{f.at(b) + off}; }
    NIS_HD void finish_line(int) const {}
  };
  NIS_HD Bound bind(const LineMap& m, float*) const { return Bound{f, m}; }
};

// row-pass epilogue: H = T / (Kzz/max + lambda), T = (-1)^(k1+c)  (the keyframe-only half of :171, cached per frame)
struct EpiHStore {
  Dst<cpx> dst; const unsigned int* max_zz; float lambda;
  struct Line {
    cpx* p; float izz, lambda; int k1;
    NIS_HD void put(int c, cpx v) const {
      const float dr = v.x * izz + lambda, di = v.y * izz;
      const float t = ((k1 + c) & 1) ? -1.f : 1.f;
      const float s = t / (dr * dr + di * di);
      p[c] = make_float2(s * dr, -s * di);
    }
  };
  struct Bound {
    Dst<cpx> dst; const unsigned int* max_zz; float lambda; LineMap m;
    NIS_HD Line line(int ln) const { int b, k1; size_t off; m.map(ln, b, off, k1); return Line{dst.at(b) + off, 1.0f / bits2f(max_zz[b]), lambda, k1}; }
  };
  NIS_HD Bound bind(const LineMap& m) const { return Bound{dst, max_zz, lambda, m}; }
};

// row-pass epilogue: store the spectrum line
struct EpiSpecStore {
  Dst<cpx> dst;
  struct Line { cpx* p; NIS_HD void put(int c, cpx v) const { p[c] = v; } };
  struct Bound {
    Dst<cpx> dst; LineMap m;
    NIS_HD Line line(int ln) const { int b, k1; size_t off; m.map(ln, b, off, k1); return Line{dst.at(b) + off}; }
  };
  NIS_HD Bound bind(const LineMap& m) const { return Bound{dst, m}; }
};

}  // namespace nis
