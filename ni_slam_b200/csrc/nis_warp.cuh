// nis_warp.cuh -- device-only, OpenCV-exact resampling of one output pixel, shared by the standalone warp kernels
// (nis_misc.cu) and by the column-pass prologues that fuse the warp into the first FFT stage (nis_col.cu).
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

// ---- column-pass prologue: the rotation feeds the first FFT stage directly, the rotated image is never stored ------
template <bool U8> struct ProRotate {
  static constexpr bool kSmemLut = U8;        // the kernel stages the 256-entry u8 -> f32/255 table in shared memory and hands it to bind()
  Src<float> f32; Src<uint8_t> u8; const float* lut; int H, W; const double* mats; const int* sel; const int2* rowtab;
  const PeakStats* polar; int D, loop;          // polar != nullptr: the slot comes from the polar-stage peak (see RotateArgs)
  struct Lane {
    const float* f; const uint8_t* u; const float* lut; int H, W; const int2* rt; int a0, b0, a1, b1;   // adelta / bdelta of both columns
    __device__ __forceinline__ float pixel(int X0, int Y0, int ad, int bd) const {
      const int X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
      // cv::warpAffine saturates X >> 5, Y >> 5 to int16; for the sizes nis_create accepts (<= 2048 px, rotation about the centre) the
      // coordinates stay within a few image sizes, far inside that range, so the saturation is the identity and is not emitted
      const int ix = X >> 5, iy = Y >> 5;
      // BORDER_WRAP, branch-free: a rotation about the centre keeps every source coordinate inside (-len, 2 len) for the aspect
      // ratios nis_create accepts (max(W,H) <= 2.8 min(W,H)), so one conditional add / subtract is the exact modulo
      const int x0 = wrap1(ix, W), y0 = wrap1(iy, H);
      const int x1 = (x0 + 1 == W) ? 0 : x0 + 1, y1 = (y0 + 1 == H) ? 0 : y0 + 1;
      const unsigned r0 = (unsigned)(y0 * W), r1 = (unsigned)(y1 * W);          // 32-bit indexing: an image has < 2^31 pixels
      float v0, v1, v2, v3;
      if (U8) {
        // exact (float)(u / 255.0) (utils.cc:117) from the staged table: one shared-memory load per tap instead of a conversion and a
        // three-instruction correctly rounded division
        v0 = lut[u[r0 + (unsigned)x0]]; v1 = lut[u[r0 + (unsigned)x1]];
        v2 = lut[u[r1 + (unsigned)x0]]; v3 = lut[u[r1 + (unsigned)x1]];
      } else {
        v0 = __ldg(f + r0 + (unsigned)x0); v1 = __ldg(f + r0 + (unsigned)x1);
        v2 = __ldg(f + r1 + (unsigned)x0); v3 = __ldg(f + r1 + (unsigned)x1);
      }
      return bilinear4(v0, v1, v2, v3, X & 31, Y & 31);
    }
    __device__ __forceinline__ cpx load(int y) const {
      const int2 xy = __ldg(rt + y);              // (X0, Y0) of warpAffine's fixed-point walk for this output row, per rotation matrix
      return make_float2(pixel(xy.x, xy.y, a0, b0), pixel(xy.x, xy.y, a1, b1));
    }
    template <int R> __device__ __forceinline__ void load_all(int row0, int stride, cpx* v) const {
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = load(row0 + r * stride);
    }
  };
  struct Bound {
    const float* f; const uint8_t* u; const float* lut; int H, W, c0; const double* M; const int2* rt;
    __device__ __forceinline__ Lane lane(int l) const {
      const int x = c0 + 2 * l;
      return Lane{f, u, lut, H, W, rt,
                  __double2int_rn(__dmul_rn(M[0], (double)x) * 1024.0), __double2int_rn(__dmul_rn(M[3], (double)x) * 1024.0),
                  __double2int_rn(__dmul_rn(M[0], (double)(x + 1)) * 1024.0), __double2int_rn(__dmul_rn(M[3], (double)(x + 1)) * 1024.0)};
    }
  };
  __device__ __forceinline__ Bound bind(int e, int c0, const float* lut_s = nullptr) const {
    int slot;
    if (polar) {
      const uint32_t idx = 0xffffffffu - (uint32_t)(polar[e >> loop].key & 0xffffffffull);      // peak_key: column-major index col * D + row
      const int row = (int)(idx % (uint32_t)D);
      slot = loop ? ((e & 1) ? 2 * D + row : D + row) : row;
    } else {
      slot = sel[e];
    }
    return Bound{U8 ? nullptr : f32.at(e), U8 ? u8.at(e) : nullptr, lut_s ? lut_s : lut, H, W, c0, mats + 6 * (size_t)slot, rowtab + (size_t)slot * H};
  }
};

}  // namespace nis
