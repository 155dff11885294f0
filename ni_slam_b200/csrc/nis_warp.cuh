// nis_warp.cuh -- device-only, OpenCV-exact resampling of one output pixel, shared by the standalone warp kernels
// (nis_misc.cu) and by the column-pass prologues that fuse the warp into the first FFT stage (nis_col.cu).
//   polar_pixel  : RemoveZeroComponent + fftshift + cv::warpPolar   (correlation_flow.cc:79-87, :94, :228-236)
//   rotate_pixel : RotateArray = getRotationMatrix2D + warpAffine   (utils.cc:154-161)
// Explicit round-to-nearest intrinsics everywhere: an FMA-contracted 4-tap sum differs from OpenCV's by 1 ulp.
#pragma once
#include <cuda_runtime.h>

#include "nis_ops.cuh"

namespace nis {

__device__ __forceinline__ int sat_short(int v) { return max(-32768, min(32767, v)); }

// OpenCV BilinearTab_f weights: {(1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx}, all exact in f32
__device__ __forceinline__ float bilinear4(float v0, float v1, float v2, float v3, int fx, int fy) {
  const float x = (float)fx * (1.f / 32.f), y = (float)fy * (1.f / 32.f);
  const float x0 = 1.f - x, y0 = 1.f - y;
  float acc = __fmul_rn(v0, y0 * x0);
  acc = __fadd_rn(acc, __fmul_rn(v1, y0 * x));
  acc = __fadd_rn(acc, __fmul_rn(v2, y * x0));
  acc = __fadd_rn(acc, __fmul_rn(v3, y * x));
  return acc;
}

// element i of the power image: plain f32 [H][W], or the pair-duplicated layout of EpiStorePairs (.x of entry i)
__device__ __forceinline__ float power_at(const float* __restrict__ p, size_t i) { return __ldg(p + i); }
__device__ __forceinline__ float power_at(const float2* __restrict__ p, size_t i) { return __ldg(&p[i].x); }

// tap of fftshift(RemoveZeroComponent(power)) at shifted coordinates (y, x); outside -> 0 (WARP_FILL_OUTLIERS)
template <class P>
__device__ __forceinline__ float shifted_tap(const P* __restrict__ p, int y, int x, int H, int W) {
  if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) return 0.f;
  int r = y - H / 2; r += (r < 0) ? H : 0;                                     // circ_shift.h:238-244
  int c = x - W / 2; c += (c < 0) ? W : 0;
  if (c == 0) return __fadd_rn(power_at(p, (size_t)r * W + 1), power_at(p, (size_t)r * W + W - 1)) * 0.5f;   // column rule (incl. (0,0))
  if (r == 0) return __fadd_rn(power_at(p, (size_t)W + c), power_at(p, (size_t)(H - 1) * W + c)) * 0.5f;      // row rule
  return power_at(p, (size_t)r * W + c);
}

// cs = (cos, sin) of the output row's angle (double, host libm); rf = (float)(rho * maxRadius / Cp).
// General (border / RemoveZeroComponent-aware) path: out of line, the table-driven fast path covers almost every pixel.
template <class P>
static __device__ __noinline__ float polar_pixel(const P* __restrict__ power, int H, int W, double cp, double sp, float rf) {
  const float cx = (float)W / 2, cy = (float)H / 2;
  const float mx = (float)__dadd_rn(__dmul_rn((double)rf, cp), (double)cx);
  const float my = (float)__dadd_rn(__dmul_rn((double)rf, sp), (double)cy);
  const int sx = __float2int_rn(mx * 32.f), sy = __float2int_rn(my * 32.f);   // cvRound: half to even
  const int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5);
  const float v0 = shifted_tap(power, iy, ix, H, W), v1 = shifted_tap(power, iy, ix + 1, H, W);
  const float v2 = shifted_tap(power, iy + 1, ix, H, W), v3 = shifted_tap(power, iy + 1, ix + 1, H, W);
  return bilinear4(v0, v1, v2, v3, sx & 31, sy & 31);
}

// BORDER_WRAP (cv::borderInterpolate): p mod len.  Written as add/subtract loops -- the coordinates of a rotation about the
// image centre stay within a few image sizes, and an integer division sequence inlined at every tap bloats the kernel.
__device__ __forceinline__ int wrap_idx(int p, int len) {
  while (p < 0) p += len;
  while (p >= len) p -= len;
  return p;
}

__device__ __forceinline__ int wrap1(int p, int len) {      // exact modulo for p in (-len, 2 len)
  p += (p < 0) ? len : 0;
  p -= (p >= len) ? len : 0;
  return p;
}

// per output row: X0, Y0 of warpAffine's fixed-point walk (AB_BITS = 10, round_delta = 16)
__device__ __forceinline__ void rotate_row_setup(const double* __restrict__ M, int y, int& X0, int& Y0) {
  X0 = __double2int_rn(__dadd_rn(__dmul_rn(M[1], (double)y), M[2]) * 1024.0) + 16;
  Y0 = __double2int_rn(__dadd_rn(__dmul_rn(M[4], (double)y), M[5]) * 1024.0) + 16;
}

template <bool U8>
__device__ __forceinline__ float rotate_pixel(const float* __restrict__ f32, const uint8_t* __restrict__ u8, const float* __restrict__ lut,
                                              int H, int W, const double* __restrict__ M, int X0, int Y0, int x) {
  const int adelta = __double2int_rn(__dmul_rn(M[0], (double)x) * 1024.0);
  const int bdelta = __double2int_rn(__dmul_rn(M[3], (double)x) * 1024.0);
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int ix = sat_short(X >> 5), iy = sat_short(Y >> 5);
  const int x0 = wrap_idx(ix, W), x1 = wrap_idx(ix + 1, W), y0 = wrap_idx(iy, H), y1 = wrap_idx(iy + 1, H);   // BORDER_WRAP
  float v0, v1, v2, v3;
  if (U8) {
    v0 = u8_to_unit(__ldg(u8 + (size_t)y0 * W + x0)); v1 = u8_to_unit(__ldg(u8 + (size_t)y0 * W + x1));
    v2 = u8_to_unit(__ldg(u8 + (size_t)y1 * W + x0)); v3 = u8_to_unit(__ldg(u8 + (size_t)y1 * W + x1));
  } else {
    v0 = __ldg(f32 + (size_t)y0 * W + x0); v1 = __ldg(f32 + (size_t)y0 * W + x1);
    v2 = __ldg(f32 + (size_t)y1 * W + x0); v3 = __ldg(f32 + (size_t)y1 * W + x1);
  }
  return bilinear4(v0, v1, v2, v3, X & 31, Y & 31);
}

// ---- polar gather table: the polar sampling grid does not depend on the image, so the fixed-point source position of
// every output pixel is computed once per context.  entry = offset(r0*W + c0) | fx << 21 | fy << 26 when the 2x2 footprint
// is an interior, unwrapped block of `power` (no RemoveZeroComponent row/column, no border); bit 31 = take the exact
// general path.
constexpr uint32_t kPolarSlow = 0x80000000u;
__device__ __forceinline__ uint32_t polar_table_entry(int H, int W, double cp, double sp, float rf) {
  const float cx = (float)W / 2, cy = (float)H / 2;
  const float mx = (float)__dadd_rn(__dmul_rn((double)rf, cp), (double)cx);
  const float my = (float)__dadd_rn(__dmul_rn((double)rf, sp), (double)cy);
  const int sx = __float2int_rn(mx * 32.f), sy = __float2int_rn(my * 32.f);
  const int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5);
  if ((size_t)H * W > (1u << 21)) return kPolarSlow;
  if (ix < 0 || iy < 0 || ix + 1 >= W || iy + 1 >= H) return kPolarSlow;
  int r0 = iy - H / 2; r0 += (r0 < 0) ? H : 0;
  int c0 = ix - W / 2; c0 += (c0 < 0) ? W : 0;
  if (r0 < 1 || c0 < 1 || r0 + 1 >= H || c0 + 1 >= W) return kPolarSlow;
  return (uint32_t)(r0 * W + c0) | ((uint32_t)(sx & 31) << 21) | ((uint32_t)(sy & 31) << 26);
}
__device__ __forceinline__ float polar_pixel_tab(const float* __restrict__ power, int H, int W, uint32_t e, double cp, double sp, float rf) {
  if (e & kPolarSlow) return polar_pixel(power, H, W, cp, sp, rf);
  const float* q = power + (e & 0x1fffffu);
  return bilinear4(__ldg(q), __ldg(q + 1), __ldg(q + W), __ldg(q + W + 1), (e >> 21) & 31, (e >> 26) & 31);
}

// ---- column-pass prologues: the warp feeds the first FFT stage directly, the warped image is never stored ------
struct ProPolar {
  Src<float2> power; int H, W, Cp; const double* cs; const float* rho; const uint32_t* table;   // table [D][Cp]
  struct Lane {
    const float2* p; int H, W, Cp, q; const double* cs; float rf0, rf1; const uint32_t* tab;
    // branch-free table path for all R rows first (every gather is in flight before the first use), then the rare
    // general-path pixels are patched
    template <int R> __device__ __forceinline__ void load_all(int phi0, int stride, cpx* v) const {
      uint2 e[R];
#pragma unroll
      for (int r = 0; r < R; ++r) e[r] = __ldg(reinterpret_cast<const uint2*>(tab + (phi0 + r * stride) * Cp + q));
      uint32_t any = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        any |= e[r].x | e[r].y;
        const float2* q0 = p + ((e[r].x & kPolarSlow) ? 0u : (e[r].x & 0x1fffffu));
        const float2* q1 = p + ((e[r].y & kPolarSlow) ? 0u : (e[r].y & 0x1fffffu));
        const float2 a0 = __ldg(q0), b0 = __ldg(q0 + W), a1 = __ldg(q1), b1 = __ldg(q1 + W);     // (tap, tap+1) of both footprint rows
        v[r] = make_float2(bilinear4(a0.x, a0.y, b0.x, b0.y, (e[r].x >> 21) & 31, (e[r].x >> 26) & 31),
                           bilinear4(a1.x, a1.y, b1.x, b1.y, (e[r].y >> 21) & 31, (e[r].y >> 26) & 31));
      }
      if (any & kPolarSlow) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int phi = phi0 + r * stride;
          if ((e[r].x | e[r].y) & kPolarSlow) {
            const double cp = __ldg(cs + 2 * phi), sp = __ldg(cs + 2 * phi + 1);
            if (e[r].x & kPolarSlow) v[r].x = polar_pixel(p, H, W, cp, sp, rf0);
            if (e[r].y & kPolarSlow) v[r].y = polar_pixel(p, H, W, cp, sp, rf1);
          }
        }
      }
    }
  };
  struct Bound {
    const float2* p; int H, W, Cp, c0; const double* cs; const float* rho; const uint32_t* tab;
    __device__ __forceinline__ Lane lane(int l) const {
      const int q = c0 + 2 * l;
      return Lane{p, H, W, Cp, q, cs, __ldg(rho + q), __ldg(rho + q + 1), tab};
    }
  };
  __device__ __forceinline__ Bound bind(int b, int c0) const { return Bound{power.at(b), H, W, Cp, c0, cs, rho, table}; }
};

template <bool U8> struct ProRotate {
  Src<float> f32; Src<uint8_t> u8; const float* lut; int H, W; const double* mats; const int* sel;
  struct Lane {
    const float* f; const uint8_t* u; const float* lut; int H, W; const double* M; int a0, b0, a1, b1;   // adelta / bdelta of both columns
    __device__ __forceinline__ float pixel(int X0, int Y0, int ad, int bd) const {
      const int X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
      const int ix = sat_short(X >> 5), iy = sat_short(Y >> 5);
      // BORDER_WRAP, branch-free: a rotation about the centre keeps every source coordinate inside (-len, 2 len) for the aspect
      // ratios nis_create accepts (max(W,H) <= 2.8 min(W,H)), so one conditional add / subtract is the exact modulo
      int x0 = wrap1(ix, W), x1 = wrap1(ix + 1, W), y0 = wrap1(iy, H), y1 = wrap1(iy + 1, H);
      const int r0 = y0 * W, r1 = y1 * W;          // 32-bit indexing: an image has < 2^31 pixels
      float v0, v1, v2, v3;
      if (U8) {
        v0 = u8_to_unit(__ldg(u + r0 + x0)); v1 = u8_to_unit(__ldg(u + r0 + x1));
        v2 = u8_to_unit(__ldg(u + r1 + x0)); v3 = u8_to_unit(__ldg(u + r1 + x1));
      } else {
        v0 = __ldg(f + r0 + x0); v1 = __ldg(f + r0 + x1);
        v2 = __ldg(f + r1 + x0); v3 = __ldg(f + r1 + x1);
      }
      return bilinear4(v0, v1, v2, v3, X & 31, Y & 31);
    }
    __device__ __forceinline__ cpx load(int y) const {
      int X0, Y0;
      rotate_row_setup(M, y, X0, Y0);
      return make_float2(pixel(X0, Y0, a0, b0), pixel(X0, Y0, a1, b1));
    }
    template <int R> __device__ __forceinline__ void load_all(int row0, int stride, cpx* v) const {
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = load(row0 + r * stride);
    }
  };
  struct Bound {
    const float* f; const uint8_t* u; const float* lut; int H, W, c0; const double* M;
    __device__ __forceinline__ Lane lane(int l) const {
      const int x = c0 + 2 * l;
      return Lane{f, u, lut, H, W, M,
                  __double2int_rn(__dmul_rn(M[0], (double)x) * 1024.0), __double2int_rn(__dmul_rn(M[3], (double)x) * 1024.0),
                  __double2int_rn(__dmul_rn(M[0], (double)(x + 1)) * 1024.0), __double2int_rn(__dmul_rn(M[3], (double)(x + 1)) * 1024.0)};
    }
  };
  __device__ __forceinline__ Bound bind(int e, int c0) const {
    return Bound{U8 ? nullptr : f32.at(e), U8 ? u8.at(e) : nullptr, lut, H, W, c0, mats + 6 * (size_t)sel[e]};
  }
};

}  // namespace nis
